#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, bench lines of the three uniform workloads.
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_cfg2.json
python bench.py --no-cpu-baseline --workload cfg3 --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_cfg3.json
python bench.py --no-cpu-baseline --workload cfg1u --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_cfg1u.json
