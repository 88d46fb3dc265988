#!/bin/bash
# final 1-GPU pass of round 2: whole GPU suite, smoke, default bench line, reference arm, launch
# list, and ONE rank of a two-process run under ncu for the pack / push kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16 | tee gpurun_out/r02i_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -8 | tee gpurun_out/r02i_smoke.txt
python bench.py 2>gpurun_out/bench_err.log | tail -1 | tee gpurun_out/r02i_bench_cfg2.json
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02i_bench_cfg2_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02i_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also --quick > gpurun_out/ncu_launches.log 2>&1
D=$(mktemp -d)
CFG='{"np":[2,1,1],"n":[16,16,16],"b":[2,8,8],"vars":40,"stencil":27,"stages":2,"seed":41}'
export MAMR_P2P_TIMEOUT_S=150
(timeout 240 ncu --set full --clock-control none -k regex:"facepack_kernel|p2p_push_kernel" -c 4 -f -o gpurun_out/r02i_pack_kernels \
    python tests/lb_worker.py uniform 0 2 $D "$CFG" > gpurun_out/ncu_pack.log 2>&1 &)
timeout 240 python tests/lb_worker.py uniform 1 2 $D "$CFG" 2>&1 | tail -2
sleep 5
tail -3 gpurun_out/ncu_pack.log
ls -la gpurun_out | tail -12
