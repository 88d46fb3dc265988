#!/bin/bash
# A/B of an experimental library build against the product build (bench --quick lines)
mkdir -p gpurun_out
for w in cfg5 cfg1u cfg1; do
  for lib in "" miniamr_b200/exp_r80.so; do
    echo "== $w lib=${lib:-product}"
    MAMR_LIB_PATH=${lib:+$PWD/$lib} python bench.py --quick --no-cpu-baseline --workload $w 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step %.4f frac %.3f'%(d['ms_per_step'], d['roofline']['frac']))"
  done
done
