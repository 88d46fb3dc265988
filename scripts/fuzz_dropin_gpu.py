#!/usr/bin/env python
"""TEST TOOL (needs one GPU): random single-rank command lines run twice in one process -- by the
UNMODIFIED reference and by the drop-in build (reference host code + CUDA stage path) -- and compared
as tests/test_integration.py compares them (same mesh, bit-identical tiles on every cell the exchange
defines, checksums, face and flop counters).
    python scripts/fuzz_dropin_gpu.py [seed] [seconds]"""
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refharness  # noqa: E402
from mputil import defined_mask  # noqa: E402

random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
t_end = time.time() + (float(sys.argv[2]) if len(sys.argv) > 2 else 60)
ok = 0
while time.time() < t_end:
    kind = random.choice(["amr7", "amr7", "amr7", "uni27", "uni0"])
    nx, ny, nz = [random.choice([2, 4, 6, 8, 10]) for _ in range(3)]
    if random.random() < 0.4:
        nx = ny = nz = random.choice([8, 10, 12, 16])          # the fixed-size fused kernel
    V = random.choice([1, 2, 3, 5]) if kind != "uni0" else random.choice([8, 9, 13])
    cv = random.choice([0, 1, 2, V])
    common = (f"--nx {nx} --ny {ny} --nz {nz} --num_vars {V} --comm_vars {cv} --max_blocks 3000 "
              f"--num_tsteps {random.choice([2, 3, 4])} --stages_per_ts {random.choice([3, 5, 7])} "
              f"--checksum_freq {random.choice([1, 2, 3, 5])} {'--permute' if random.random() < 0.4 else ''} "
              f"--init_x {random.choice([1, 2])} --init_y {random.choice([1, 2])} --init_z {random.choice([1, 2])}")
    if kind == "amr7":
        objs = []
        n_obj = random.choice([1, 2])
        for _ in range(n_obj):
            c = [round(random.uniform(0.1, 0.9), 2) for _ in range(3)]
            mv = [round(random.uniform(-0.1, 0.1), 2) for _ in range(3)]
            sz = [round(random.uniform(0.1, 0.4), 2) for _ in range(3)]
            objs.append(f"--object {random.choice([0, 2, 2, 4, 6, 8])} 0 {c[0]} {c[1]} {c[2]} {mv[0]} {mv[1]} {mv[2]} "
                        f"{sz[0]} {sz[1]} {sz[2]} 0 0 0")
        args = (f"{common} --stencil 7 --num_refine {random.choice([1, 2, 3])} --block_change {random.choice([0, 1])} "
                f"--refine_freq {random.choice([1, 2])} --num_objects {n_obj} {' '.join(objs)}")
    else:
        args = f"{common} --stencil {27 if kind == 'uni27' else 0} --uniform_refine 1 --num_refine {random.choice([0, 1])}"
    ref = refharness.RefMiniAMR(args.split(), variant="ref", run_driver=True)
    dev = refharness.RefMiniAMR(args.split(), variant="int", run_driver=True)
    p = ref.p
    rs, rl, rnl, rne = ref.topology()
    ds, dl, dnl, dne = dev.topology()
    assert (rs == ds).all() and (rl == dl).all() and (rnl == dnl).all(), args
    cvv = p["comm_vars"] if 0 < p["comm_vars"] <= p["num_vars"] else p["num_vars"]
    for start in range(0, p["num_vars"], cvv):
        ref.comm(start, min(cvv, p["num_vars"] - start), 0)
        dev.comm(start, min(cvv, p["num_vars"] - start), 0)
    dev.sync_host()
    mask = defined_mask(p["nx"], p["ny"], p["nz"], p["stencil"])[None]
    for s in rs:
        bad = (ref.get_slot(int(s)).view(np.uint64) != dev.get_slot(int(s)).view(np.uint64)) & mask
        assert not bad.any(), f"{args}\nslot {s}: {int(bad.sum())} cells differ, first {np.argwhere(bad)[0]}"
    for v in range(p["num_vars"]):
        a, b = ref.lib.refh_get_grid_sum(v), dev.lib.refh_get_grid_sum(v)
        assert abs(a - b) <= 1e-13*abs(a), (args, v, a, b)
    assert ref.counters() == dev.counters(), args
    assert ref.flops() == dev.flops(), args
    ok += 1
print(f"drop-in fuzz: {ok} random command lines, all identical to the reference")
