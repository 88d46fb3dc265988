#!/bin/bash
# 1-GPU box: parity tests, bench lines, ncu launch list of the default bench command and
# ncu --set full captures of the hot kernels.  Results land in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_cfg2.json
python bench.py --impl reference --steps 20 2>&1 | tail -1 | tee gpurun_out/bench_cfg2_reference.json
python bench.py --no-cpu-baseline --workload cfg3 2>&1 | tail -1 | tee gpurun_out/bench_cfg3.json
python bench.py --no-cpu-baseline --workload cfg1u 2>&1 | tail -1 | tee gpurun_out/bench_cfg1u.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:fused2 -s 3 -c 1 -o gpurun_out/fused2_cfg2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused2.log 2>&1
$NCU -k regex:slab7 -s 3 -c 1 -o gpurun_out/slab7_cfg3 python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_slab7.log 2>&1
$NCU -k regex:fused2 -s 3 -c 1 -o gpurun_out/fused2_cfg1u python bench.py --workload cfg1u --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fused2_10.log 2>&1
$NCU -k regex:checksum_tile -s 0 -c 1 -o gpurun_out/checksum_cfg2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_checksum.log 2>&1
MAMR_NO_FUSED=1 $NCU -k regex:"ghost_phase|stencil_kernel" -s 8 -c 4 -o gpurun_out/split_cfg2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_split.log 2>&1
ls -la gpurun_out
