"""--stencil 0: the per-cell update functions of
miniamr_b200/csrc/stencil0.cuh -- the ones the CUDA kernel executes -- compiled for
the host (tests/s0_host.cpp) and pinned bit for bit, flop counters included, against the
unmodified reference's stencil_driver() for every update kind (stage % 6) and every
variable class."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import refharness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "s0_host.cpp")
LIB = os.path.join(ROOT, "tests", "_build", "libs0host.so")

needs_ref = pytest.mark.skipif(not refharness.available("ref"), reason="oracle/_ref not built")


def host_lib():
    hdr = os.path.join(ROOT, "miniamr_b200", "csrc", "stencil0.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.s0_host_driver.argtypes = [C.c_void_p] + [C.c_int]*7 + [C.c_double, C.c_void_p, C.c_void_p]
    return L


@needs_ref
@pytest.mark.parametrize("args", [
    "--nx 4 --ny 6 --nz 8 --num_vars 9 --stencil 0 --uniform_refine 1 --num_refine 1 --max_blocks 40 "
    "--num_tsteps 1 --stages_per_ts 1",
    "--nx 6 --ny 4 --nz 4 --num_vars 14 --comm_vars 5 --stencil 0 --num_refine 2 --max_blocks 400 "
    "--num_objects 1 --object 2 0 0.3 0.3 0.3 0.01 0.01 0.01 0.25 0.25 0.25 0 0 0 --num_tsteps 1 --stages_per_ts 1",
])
def test_every_kind_and_class_matches_reference(args):
    L = host_lib()
    r = refharness.RefMiniAMR(args.split(), variant="ref")
    r.init()
    r.refine(0)
    p = r.p
    V = p["num_vars"]
    cv = p["comm_vars"] if 0 < p["comm_vars"] <= V else V
    mat, a1, a0 = r.stencil0()
    assert mat == V//4 and mat >= 2
    slots = [int(s) for s in r.sorted_slots()]
    flops = np.zeros(3)
    f0 = r.flops()
    for stage in range(13):                      # every kind twice, values evolve
        for start in range(0, V, cv):
            r.comm(start, min(cv, V - start), stage)
            for var in range(start, min(start + cv, V)):
                before = {s: r.get_slot(s) for s in slots}
                r.stencil_driver(var, stage)
                for s in slots:
                    mine = np.ascontiguousarray(before[s])
                    L.s0_host_driver(mine.ctypes.data, p["nx"], p["ny"], p["nz"], V, var, stage, mat, a1,
                                     a0.ctypes.data, flops.ctypes.data)
                    want = r.get_slot(s)
                    bad = mine.view(np.uint64) != want.view(np.uint64)
                    assert not bad.any(), (f"stage {stage} (kind {stage % 6}) var {var} slot {s}: "
                                           f"{int(bad.sum())} cells differ, first {np.argwhere(bad)[0]}")
    f1 = r.flops()
    assert flops[0] == f1["adds"] - f0["adds"]
    assert flops[1] == f1["muls"] - f0["muls"]
    assert flops[2] == f1["divs"] - f0["divs"]


@needs_ref
@pytest.mark.parametrize("args", [
    "--nx 4 --ny 6 --nz 8 --num_vars 9 --stencil 0 --uniform_refine 1 --num_refine 1 --max_blocks 40 "
    "--num_tsteps 1 --stages_per_ts 1",
    "--nx 6 --ny 4 --nz 4 --num_vars 14 --comm_vars 5 --stencil 0 --uniform_refine 1 --num_refine 1 --permute "
    "--init_x 2 --init_y 1 --init_z 2 --max_blocks 100 --num_tsteps 1 --stages_per_ts 1",
])
def test_oracle_restatement_matches_reference(args):
    """oracle/oracle.c: orc_stencil0_driver (an independent plain-C restatement) + the oracle's
    comm(), stage by stage against the unmodified reference: tiles bit for bit (ghosts
    included), flop counters exactly."""
    from oracle.oracle import OracleMesh
    r = refharness.RefMiniAMR(args.split(), variant="ref")
    r.init()
    r.refine(0)
    p = r.p
    V = p["num_vars"]
    mat, a1, a0 = r.stencil0()
    slots, level, nei_level, nei = r.topology()
    m = OracleMesh(p["nx"], p["ny"], p["nz"], V, p["max_blocks"], stencil=0, comm_vars=p["comm_vars"],
                   permute=p["permute"])
    m.set_topology(slots, level, nei_level, nei)
    m.set_stencil0(mat, a1, a0)
    for s in slots:
        m.data[s] = r.get_slot(int(s))
    f0 = r.flops()
    for stage in range(13):
        r.stage(stage)
        m.stage(stage)
        for s in slots:
            bad = m.data[s].view(np.uint64) != r.get_slot(int(s)).view(np.uint64)
            assert not bad.any(), f"stage {stage} (kind {stage % 6}) slot {s}: first {np.argwhere(bad)[0]}"
    f1 = r.flops()
    assert list(m.flops) == [f1["adds"] - f0["adds"], f1["muls"] - f0["muls"], f1["divs"] - f0["divs"]]
