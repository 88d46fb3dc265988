#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_lookahead.py -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_la.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee -a gpurun_out/pytest_la.log
CFG1="--nx 10 --ny 10 --nz 10 --num_vars 40 --stencil 7 --num_refine 4 --max_blocks 4000 --num_objects 1 --object 2 0 0.3 0.3 0.3 0.01 0.01 0.01 0.25 0.25 0.25 0 0 0 --num_tsteps 20 --stages_per_ts 20"
MAMR_VERBOSE=1 integration/_bin/miniAMR_b200.x $CFG1 2>&1 | grep -i "summary\|miniamr_b200\|error" | tee gpurun_out/cfg1_dropin_lookahead.log
MAMR_NO_LOOKAHEAD=1 MAMR_VERBOSE=1 integration/_bin/miniAMR_b200.x $CFG1 2>&1 | grep -i "summary\|miniamr_b200\|error" | tee -a gpurun_out/cfg1_dropin_lookahead.log
