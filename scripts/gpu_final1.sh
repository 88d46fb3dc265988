#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg2.json
python bench.py --no-cpu-baseline --workload cfg3 2>&1 | tail -1 | tee gpurun_out/bench_cfg3.json
python bench.py --no-cpu-baseline --workload cfg1u 2>&1 | tail -1 | tee gpurun_out/bench_cfg1u.json
