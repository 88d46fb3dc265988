#!/bin/bash
# 2-GPU box: real NVLink hop of the peer-memory transport (CUDA IPC windows), NCCL transport
# beside it, bench at N=2 with both
set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -8
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -6 | tee gpurun_out/r02c_pytest_mgpu2.txt
timeout 600 python -m pytest tests/test_multi_rank_dropin.py -x -q -k "rcb or two_objects or nccl or staged" 2>&1 | tail -6 | tee gpurun_out/r02c_pytest_dropin2.txt
timeout 300 bash scripts/bench_n.sh 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02c_bench_cfg2_n2_p2p.json
timeout 300 bash scripts/bench_n.sh 2 --no-cpu-baseline --transport nccl 2>&1 | tail -1 | tee gpurun_out/r02c_bench_cfg2_n2_nccl.json
