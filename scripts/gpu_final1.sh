#!/bin/bash
# final 1-GPU pass of the round: tests, smoke, bench lines, launch list, one full capture
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_cfg2.json
python bench.py --impl reference --steps 20 2>&1 | tail -1 | tee gpurun_out/bench_cfg2_reference.json
python bench.py --no-cpu-baseline --workload cfg3 2>&1 | tail -1 | tee gpurun_out/bench_cfg3.json
python bench.py --no-cpu-baseline --workload cfg1u 2>&1 | tail -1 | tee gpurun_out/bench_cfg1u.json
python bench.py --no-cpu-baseline --workload cfg5 2>&1 | tail -1 | tee gpurun_out/bench_cfg5.json
python scripts/s0_time.py 12 2>&1 | tail -1 | tee gpurun_out/s0_time3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:fused2 -s 3 -c 1 -o gpurun_out/fused2_cfg2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused2.log 2>&1
