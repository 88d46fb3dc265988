#!/usr/bin/env python
"""--stencil 0 timing on one GPU: ms per stage by update kind (stage % 6) on a uniform
mesh of 10^3-cell blocks, 40 variables, beside the unmodified reference on the host cores.
    python scripts/s0_time.py [blocks_per_edge]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miniamr_b200.capi import DeviceMesh  # noqa: E402
from miniamr_b200.mesh import uniform_mesh  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
n, V = 10, 40
nb = B**3
top = uniform_mesh(B, B, B, 1, 1, 1, 0, n, n, n, comm_vars=V, stencil=0)
d = DeviceMesh(n, n, n, V, nb, stencil=0)
d.set_topology(top["slots"], top["level"], top["nei_level"], top["nei"])
rs = np.random.RandomState(1)
d.set_stencil0(V//4, rs.random_sample(), rs.random_sample(V//4))
tile = np.zeros((V, n + 2, n + 2, n + 2))
for s in range(nb):
    tile[:, 1:-1, 1:-1, 1:-1] = rs.random_sample((V, n, n, n)) if s < 64 else tile[:, 1:-1, 1:-1, 1:-1]
    d.upload_block(s, tile)
for st in range(6):
    d.stage(st)
d.sync()
names = ["pointwise", "sweep i", "sweep j", "sweep k", "7-pt weighted", "27-pt banded"]
out = {}
for kind in range(6):
    d.sync()
    d.kernel_timing(True)
    d.timer_begin()
    for rep in range(3):
        d.stage(6*(rep + 1) + kind)
    ms = d.timer_end()/3
    kt = d.kernel_times()
    d.kernel_timing(False)
    out[names[kind]] = dict(ms_per_stage=ms, upd_per_s=nb*n**3*V/(ms*1e-3),
                            update_kernels_ms=kt["stencil_ms"]/3, ghost_exchange_ms=kt["ghost_ms"]/3)
d.close()
res = dict(blocks=nb, cells=n**3, num_vars=V, device=out)
try:
    from oracle import refharness
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    v = "omp" if refharness.available("omp") else "ref"
    r = refharness.RefMiniAMR(f"--nx {n} --ny {n} --nz {n} --num_vars {V} --stencil 0 --uniform_refine 1 "
                              f"--num_refine 2 --max_blocks 80".split(), variant=v)
    r.init(); r.refine(0)
    r.stage(0)
    t0 = time.perf_counter()
    for st in range(6):
        r.stage(st)
    dt = (time.perf_counter() - t0)/6
    res["reference"] = dict(variant=v, cores=os.cpu_count(), blocks=r.p["num_active"], ms_per_stage=dt*1e3,
                            upd_per_s=r.p["num_active"]*n**3*V/dt)
except Exception as e:
    res["reference"] = str(e)
print(json.dumps(res))
