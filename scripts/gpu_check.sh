#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, a bench line, the ncu launch list.
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --blocks 8 --steps 5 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_small.json
python bench.py 2>&1 | tail -3 | tee gpurun_out/bench_full.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
