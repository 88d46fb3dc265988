#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_stencil0.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_s0.log
python -m pytest tests/test_integration.py -x -q -k "uni0 or amr7_moving" 2>&1 | tail -15 | tee -a gpurun_out/pytest_s0.log
