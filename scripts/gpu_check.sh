#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, a bench line, the ncu launch list.
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_quick.json
MAMR_NO_ELIDE=1 python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_noelide.json
MAMR_NO_FUSED2=1 python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_nofused2.json
