#!/bin/bash
# experiments on one GPU: block visiting order, cfg5 bench, the drop-in on configs[0]
mkdir -p gpurun_out
run() { python bench.py --no-cpu-baseline --steps 10 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4f ms  %.4g upd/s  frac %.3f'%(d['ms_per_step'], d['value'], d['roofline']['frac']))" || echo failed; }
for o in zchain brick:4,4,16 brick:2,2,16 brick:8,8,16 brick:4,4,4 brick:2,8,16 brick:16,1,1; do
  echo "cfg2 ORDER=$o: $(MAMR_ORDER=$o run)"
done 2>&1 | tee gpurun_out/exp_order.log
for o in zchain brick:4,4,16 brick:8,8,16; do
  echo "cfg1u ORDER=$o: $(MAMR_ORDER=$o run --workload cfg1u)"
done 2>&1 | tee -a gpurun_out/exp_order.log
for v in 10 8 5; do
  echo "cfg2 VPC=$v brick:4,4,16: $(MAMR_VPC=$v MAMR_ORDER=brick:4,4,16 run)"
done 2>&1 | tee -a gpurun_out/exp_order.log
python bench.py --no-cpu-baseline --workload cfg5 --steps 10 2>&1 | tail -1 | tee gpurun_out/bench_cfg5.json
# configs[0]: the reference program vs the drop-in, same command line (fewer time steps)
CFG1="--nx 10 --ny 10 --nz 10 --num_vars 40 --stencil 7 --num_refine 4 --max_blocks 4000 --num_objects 1 --object 2 0 0.3 0.3 0.3 0.01 0.01 0.01 0.25 0.25 0.25 0 0 0 --num_tsteps 20 --stages_per_ts 20"
( time integration/_bin/miniAMR_b200.x $CFG1 ) 2>&1 | grep -i "summary\|real\|error" | tee gpurun_out/cfg1_dropin.log
CFG1S="--nx 10 --ny 10 --nz 10 --num_vars 40 --stencil 7 --num_refine 4 --max_blocks 4000 --num_objects 1 --object 2 0 0.3 0.3 0.3 0.01 0.01 0.01 0.25 0.25 0.25 0 0 0 --num_tsteps 4 --stages_per_ts 20"
( time integration/_bin/miniAMR_b200.x $CFG1S ) 2>&1 | grep -i "summary\|real\|error" | tee -a gpurun_out/cfg1_dropin.log
( time oracle/_ref/miniAMR_ref.x $CFG1S ) 2>&1 | grep -i "summary\|real\|error" | tee gpurun_out/cfg1_ref.log
