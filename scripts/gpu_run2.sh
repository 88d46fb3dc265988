#!/bin/bash
# 1-GPU box: the whole GPU suite (loopback of the off-rank path included), N-rank drop-in with
# all ranks on one GPU (CUDA IPC windows between processes on the same device)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=10 2>&1 | tail -30 | tee gpurun_out/r02b_pytest_gpu.txt
