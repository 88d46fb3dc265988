#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_final.log
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_cfg2_final.json
