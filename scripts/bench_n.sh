#!/bin/bash
# usage: bench_n.sh N [bench args]  -- launch bench.py exactly like the driver does
N=$1; shift
if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@"; else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N "$@"; fi
