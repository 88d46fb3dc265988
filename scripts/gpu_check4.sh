#!/bin/bash
# 4-GPU box: partners in two directions
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_multi_rank_dropin.py -q -k "_4 or code" 2>&1 | tail -15 | tee gpurun_out/pytest_dropin4.log
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "11-1 or 12-1 or 13-1 or 11-0" 2>&1 | tail -8 | tee gpurun_out/pytest_mgpu4.log
timeout 600 bash scripts/bench_n.sh 4 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg2_n4.json
timeout 600 bash scripts/bench_n.sh 4 --steps 10 --no-cpu-baseline --workload cfg3 2>&1 | tail -1 | tee gpurun_out/bench_cfg3_n4.json
