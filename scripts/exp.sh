#!/bin/bash
for d in ${SKIPS:-0 4 8 12}; do
  echo "SKIP=$d"; MAMR_DEBUG_SKIP=$d python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])" || echo failed
done
