#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lookahead.py tests/test_integration.py tests/test_gpu_stencil0.py -x -q 2>&1 | tail -4
OUT=gpurun_out/r02e_cfg1_b200.txt timeout 300 bash scripts/dropin_cfg.sh cfg1 b200 1
OUT=gpurun_out/r02e_cfg4s_b200_2.txt timeout 300 bash scripts/dropin_cfg.sh cfg4s b200 2
OUT=gpurun_out/r02e_cfg4s_ref_2.txt timeout 300 bash scripts/dropin_cfg.sh cfg4s ref 2
OUT=gpurun_out/r02e_cfg5s_b200_2.txt timeout 300 bash scripts/dropin_cfg.sh cfg5s b200 2
python bench.py --quick --no-cpu-baseline --workload cfg1 2>&1 | tail -1 | cut -c1-400
