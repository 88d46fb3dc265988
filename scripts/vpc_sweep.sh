#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 8 10 20 40; do
  echo "VPC=$v"; MAMR_VPC=$v python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
