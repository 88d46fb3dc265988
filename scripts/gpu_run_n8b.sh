#!/bin/bash
set -x
mkdir -p gpurun_out
for w in cfg2 cfg3; do
  MAMR_TRACE=1 timeout 200 bash scripts/bench_n.sh 8 --no-cpu-baseline --quick --no-also --workload $w > gpurun_out/r02g_n8_$w.log 2>&1
  grep "trace rank" gpurun_out/r02g_n8_$w.log | sort | head -8
  tail -1 gpurun_out/r02g_n8_$w.log | cut -c1-330
done
