#!/bin/bash
set -x
mkdir -p gpurun_out
for w in cfg2 cfg3; do
  MAMR_TRACE=1 timeout 200 bash scripts/bench_n.sh 8 --no-cpu-baseline --quick --no-also --workload $w > gpurun_out/r02h_n8_$w.log 2>&1
  grep -o "trace rank [0-9]: .*" gpurun_out/r02h_n8_$w.log | cut -c1-200 | sort | head -8
  tail -1 gpurun_out/r02h_n8_$w.log | cut -c1-330
done
timeout 200 bash scripts/bench_n.sh 8 --no-cpu-baseline --quick --no-also --blocks 8 > gpurun_out/r02h_n8_cfg2_b8.log 2>&1
tail -1 gpurun_out/r02h_n8_cfg2_b8.log | cut -c1-330
