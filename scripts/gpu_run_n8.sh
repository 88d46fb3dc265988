#!/bin/bash
# 8-GPU box (charged 8x): parity at 8 ranks over NVLink, bench at N=8 with both transports,
# BASELINE configs[3] and [4] at full size through the drop-in program
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -k "test_two_ranks_match_single_rank_oracle and 15-1" 2>&1 | tail -3 | tee gpurun_out/r02f_pytest_mgpu8.txt
timeout 300 python -m pytest tests/test_multi_rank_dropin.py -x -q -k "two_objects_8" 2>&1 | tail -3 | tee gpurun_out/r02f_pytest_dropin8.txt
timeout 300 bash scripts/bench_n.sh 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02f_bench_cfg2_n8_p2p.json
timeout 300 bash scripts/bench_n.sh 8 --no-cpu-baseline --quick --transport nccl 2>&1 | tail -1 | tee gpurun_out/r02f_bench_cfg2_n8_nccl.json
timeout 200 bash scripts/bench_n.sh 8 --no-cpu-baseline --quick --no-also --blocks 8 2>&1 | tail -1 | tee gpurun_out/r02f_bench_cfg2_n8_b8_p2p.json
timeout 200 bash scripts/bench_n.sh 8 --no-cpu-baseline --quick --no-also --blocks 8 --transport nccl 2>&1 | tail -1 | tee gpurun_out/r02f_bench_cfg2_n8_b8_nccl.json
OUT=gpurun_out/r02f_cfg4_b200_8.txt timeout 300 bash scripts/dropin_cfg.sh cfg4 b200 8 | tee gpurun_out/r02f_cfg4_b200_8_summary.txt
OUT=gpurun_out/r02f_cfg5_b200_8.txt timeout 300 bash scripts/dropin_cfg.sh cfg5 b200 8 | tee gpurun_out/r02f_cfg5_b200_8_summary.txt
