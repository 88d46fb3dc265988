#!/bin/bash
# BASELINE configs[3] and [4] at full size through the reference's own main()/driver() with the
# CUDA stage path underneath (integration/_bin/miniAMR_b200_mp.x, one rank per GPU), or through the
# unmodified reference (oracle/_ref/miniAMR_ref_mp.x) on the host cores.
#   usage: dropin_cfg.sh cfg4|cfg5|cfg4s|cfg5s  b200|ref  [ranks]
CFG=$1; IMPL=$2; N=${3:-8}
OBJ="--num_objects 2 --object 2 0 0.3 0.3 0.3 0.01 0.01 0.01 0.25 0.25 0.25 0 0 0 --object 0 0 0.5 0.5 0.9 0 0 -0.01 0.6 0.6 0.02 0 0 0"
case $N in 8) NP="--npx 2 --npy 2 --npz 2";; 4) NP="--npx 2 --npy 2";; 2) NP="--npx 2";; 1) NP="";; esac
case $CFG in
  cfg1)  ARGS="$NP --nx 10 --ny 10 --nz 10 --num_vars 40 --stencil 7 --num_refine 4 --max_blocks 4000 --num_objects 1 --object 2 0 0.3 0.3 0.3 0.01 0.01 0.01 0.25 0.25 0.25 0 0 0 --num_tsteps 20 --stages_per_ts 20";;
  cfg4)  ARGS="$NP --init_x 1 --init_y 1 --init_z 1 --nx 10 --ny 10 --nz 10 --num_vars 40 --num_refine 5 --max_blocks 6000 --refine_freq 5 --num_tsteps 20 --stages_per_ts 20 --lb_opt 1 $OBJ";;
  cfg4s) ARGS="$NP --init_x 1 --init_y 1 --init_z 1 --nx 10 --ny 10 --nz 10 --num_vars 40 --num_refine 3 --max_blocks 3000 --refine_freq 2 --num_tsteps 4 --stages_per_ts 5 --lb_opt 1 $OBJ";;
  cfg5)  ARGS="$NP --init_x 3 --init_y 3 --init_z 3 --nx 10 --ny 10 --nz 10 --num_vars 160 --comm_vars 40 --stencil 27 --uniform_refine 1 --num_refine 2 --max_blocks 1800 --num_tsteps 2 --stages_per_ts 10 --checksum_freq 1";;
  cfg5s) ARGS="$NP --init_x 1 --init_y 1 --init_z 1 --nx 10 --ny 10 --nz 10 --num_vars 160 --comm_vars 40 --stencil 27 --uniform_refine 1 --num_refine 1 --max_blocks 100 --num_tsteps 2 --stages_per_ts 4 --checksum_freq 1";;
esac
ROOT=$(cd "$(dirname "$0")/.." && pwd)
if [ "$IMPL" = ref ]; then EXE=$ROOT/oracle/_ref/miniAMR_ref_mp.x; else EXE=$ROOT/integration/_bin/miniAMR_b200_mp.x; fi
echo "== $CFG $IMPL ranks=$N: $ARGS"
OUT=${OUT:-/dev/null}
MAMR_VERBOSE=1 $ROOT/minimpi/_bin/minimpirun -n $N $EXE $ARGS --report_diffusion 2>&1 | tee $OUT | grep -i "summary\|total time\|miniamr_b200\|error\|Total number of blocks at timestep 0" | head -40
