#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lookahead.py tests/test_loopback.py -x -q -k "not test_ranks_on_one_gpu or 13 or 15 or 12" 2>&1 | tail -5
OUT=gpurun_out/r02d_cfg4s_b200_2.txt timeout 300 bash scripts/dropin_cfg.sh cfg4s b200 2
OUT=gpurun_out/r02d_cfg4s_ref_2.txt timeout 300 bash scripts/dropin_cfg.sh cfg4s ref 2
for w in cfg1 cfg5; do python bench.py --quick --no-cpu-baseline --workload $w 2>&1 | tail -1 | tee gpurun_out/r02d_bench_$w.json; done
python bench.py --quick --no-cpu-baseline --no-also 2>&1 | tail -1 | tee gpurun_out/r02d_bench_cfg2.json
