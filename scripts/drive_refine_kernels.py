#!/usr/bin/env python
"""Launch the block-data kernels of refinement and migration a few times at BASELINE shape
(16^3 cells x 40 variables) so that ncu can capture them:
split_kernel / consolidate_kernel (block.c:161-173, 418-430), block_payload_kernel
(pack.c:66-70, 103-107).  usage: [ncu ...] python scripts/drive_refine_kernels.py [n] [vars]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miniamr_b200.capi import DeviceMesh  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
V = int(sys.argv[2]) if len(sys.argv) > 2 else 40
MB = 64
d = DeviceMesh(n, n, n, V, MB, stencil=7)
rs = np.random.RandomState(0)
for s in range(MB):
    d.upload_block(s, rs.random_sample((V, n + 2, n + 2, n + 2)))
for rep in range(3):
    kids = np.arange(8, dtype=np.int32) + 8*(rep + 1)
    d.split_block(rep, kids)
    d.consolidate_block(kids, 40 + rep)
    p = d.pack_block(50 + rep)
    d.unpack_block(60 + rep, p)
d.sync()
print("launches", d.counters()["kernel_launches"])
d.close()
