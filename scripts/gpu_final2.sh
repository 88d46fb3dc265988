#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_rank_dropin.py -q -k "uni0 or amr7_rcb or uni27_staged or morton" 2>&1 | tail -8 | tee gpurun_out/pytest_dropin2_final.log
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -k "6-1 or 8-1 or 1-1" 2>&1 | tail -5 | tee gpurun_out/pytest_mgpu2_final.log
timeout 600 bash scripts/bench_n.sh 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg2_n2_final.json
