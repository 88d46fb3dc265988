#!/bin/bash
# 2-GPU box: the N-rank drop-in against the N-rank reference, the 2-rank device tests, bench at N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_rank_dropin.py -q 2>&1 | tail -30 | tee gpurun_out/pytest_dropin2.log
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -k "1]" 2>&1 | tail -8 | tee gpurun_out/pytest_mgpu2.log
timeout 600 bash scripts/bench_n.sh 2 --steps 10 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_cfg2_n2.json
timeout 600 bash scripts/bench_n.sh 2 --steps 10 --no-cpu-baseline --workload cfg3 2>&1 | tail -1 | tee gpurun_out/bench_cfg3_n2.json
