#!/bin/bash
# 1-GPU box, start of round 2: parity tests (incl. the BASELINE-variable-count goldens), smoke,
# bench lines per workload, ncu launch list + full captures of the kernels below roofline.
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02a_pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -8 | tee gpurun_out/r02a_smoke.txt
python bench.py 2>gpurun_out/bench_err.log | tail -1 | tee gpurun_out/r02a_bench_cfg2.json
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02a_bench_cfg2_reference.json
for w in cfg1 cfg1u cfg5; do
  python bench.py --no-cpu-baseline --workload $w 2>&1 | tail -1 | tee gpurun_out/r02a_bench_$w.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02a_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:fused2 -s 3 -c 1 -o gpurun_out/r02a_fused2_cfg5 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg5.log 2>&1
$NCU -k regex:fused2 -s 3 -c 1 -o gpurun_out/r02a_fused2_cfg1 python bench.py --workload cfg1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cfg1.log 2>&1
$NCU -k regex:"split_kernel|consolidate_kernel|block_payload" -c 6 -o gpurun_out/r02a_refine_kernels python scripts/drive_refine_kernels.py > gpurun_out/ncu_refine.log 2>&1
ls -la gpurun_out | tail -20
