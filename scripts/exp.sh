#!/bin/bash
export MAMR_NO_FUSED3=1
for d in 0 1 2; do
  echo "SKIP=$d"; MAMR_DEBUG_SKIP=$d python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" || echo failed
done
