#!/bin/bash
# ncu --set full of one fused-kernel launch at the bench size; report lands in gpurun_out/$1.ncu-rep
name=${1:-fused}
ncu --set full --clock-control none --import-source on -k regex:"fused" -s 3 -c 1 -f -o gpurun_out/$name \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${@:2} > gpurun_out/ncu_$name.log 2>&1
tail -2 gpurun_out/ncu_$name.log
