"""--stencil 0 (the "variable work" mix, stencil.c:43-74,147-983) on
the device, through the C ABI, against the UNMODIFIED reference stepped call by call:
after every stage (every update kind stage % 6, every variable class, stencil_check
included) every tile must be bit-identical, ghost cells included, and the flop counters
the reference books (data-dependent for stencil_check) must agree exactly."""
import numpy as np
import pytest

from oracle import refharness

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not refharness.available("ref"), reason="oracle/_ref not built")

CASES = {
    # one block, reflective boundaries everywhere
    "one_block": "--nx 4 --ny 6 --nz 8 --num_vars 9 --stencil 0 --max_blocks 8 --num_tsteps 1 --stages_per_ts 1",
    # uniform mesh, anisotropic blocks, staged comm, mat = 3 with two leftover variables
    "uniform": "--nx 6 --ny 4 --nz 8 --num_vars 14 --comm_vars 5 --stencil 0 --uniform_refine 1 --num_refine 1 "
               "--init_x 2 --init_y 1 --init_z 2 --max_blocks 80 --num_tsteps 1 --stages_per_ts 1",
    # the default block size, --permute, 40 variables (configs[0] variable count)
    "cfg1_shape": "--nx 10 --ny 10 --nz 10 --num_vars 40 --stencil 0 --uniform_refine 1 --num_refine 1 "
                  "--max_blocks 16 --permute --num_tsteps 1 --stages_per_ts 1",
}


@needs_ref
@pytest.mark.parametrize("name", sorted(CASES))
def test_every_stage_matches_reference(name):
    from miniamr_b200.capi import DeviceMesh
    r = refharness.RefMiniAMR(CASES[name].split(), variant="ref")
    r.init()
    r.refine(0)
    p = r.p
    V = p["num_vars"]
    cv = p["comm_vars"] if 0 < p["comm_vars"] <= V else V
    mat, a1, a0 = r.stencil0()
    slots, level, nei_level, nei = r.topology()
    d = DeviceMesh(p["nx"], p["ny"], p["nz"], V, p["max_blocks"], stencil=0, comm_vars=p["comm_vars"],
                   permute=p["permute"])
    d.set_topology(slots, level, nei_level, nei)
    d.set_stencil0(mat, a1, a0)
    for s in slots:
        d.upload_block(int(s), r.get_slot(int(s)))
    f0 = r.flops()
    for stage in range(13):                       # every kind at least twice
        for start in range(0, V, cv):
            num = min(cv, V - start)
            r.comm(start, num, stage)
            d.comm(start, num, stage)
            for var in range(start, start + num):
                r.stencil_driver(var, stage)
                d.stencil_driver(var, stage)
        for s in slots:
            got, want = d.download_block(int(s)), r.get_slot(int(s))
            bad = got.view(np.uint64) != want.view(np.uint64)
            assert not bad.any(), (f"{name}: stage {stage} (kind {stage % 6}) slot {s}: {int(bad.sum())} cells "
                                   f"differ, first (var, i, j, k) = {np.argwhere(bad)[0]}")
        for v in (0, 1, V - 1):
            want = r.check_sum(v)
            assert abs(d.check_sum(v) - want) <= 1e-13*abs(want)
    f1, c = r.flops(), d.counters()
    assert c["total_fp_adds"] == f1["adds"] - f0["adds"]
    assert c["total_fp_muls"] == f1["muls"] - f0["muls"]
    assert c["total_fp_divs"] == f1["divs"] - f0["divs"]
    d.close()


def test_stencil0_needs_its_coefficients():
    from miniamr_b200.capi import DeviceMesh, MamrError
    d = DeviceMesh(4, 4, 4, 8, 4, stencil=0)
    d.set_topology(np.array([0]), np.array([0]), np.full((1, 6), -2), np.zeros((1, 6, 2, 2), int))
    d.stencil_driver(1, 0)
    with pytest.raises(MamrError, match="mamr_set_stencil0"):
        d.sync()
    d.close()
