#!/bin/bash
# 8-GPU box: partners in all three directions (kept short: 8x the box time)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -k "14-1 or 15-1" 2>&1 | tail -8 | tee gpurun_out/pytest_mgpu8.log
timeout 600 python -m pytest tests/test_multi_rank_dropin.py -q -k "uni27_staged_8 or two_objects_8" 2>&1 | tail -15 | tee gpurun_out/pytest_dropin8.log
timeout 300 bash scripts/bench_n.sh 8 --steps 10 --no-cpu-baseline --blocks 8 2>&1 | tail -1 | tee gpurun_out/bench_cfg2_n8_small.json
