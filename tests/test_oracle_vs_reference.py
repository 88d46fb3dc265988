"""Pins oracle/oracle.c (the CPU restatement) bit-for-bit against the UNMODIFIED
reference compiled into oracle/_ref/libminiamr_ref.so (the
reference ships no golden vectors, so the reference itself is the authority)."""
import numpy as np
import pytest

from oracle.refharness import RefMiniAMR
from refutil import (MOVING, SPHERE, assert_bits_equal, compare_all, needs_ref,
                     oracle_from_ref, sync_oracle)

pytestmark = needs_ref


def test_known_answer_checksums():
    # SURVEY.md §4 / BASELINE.md: glibc rand() seed 1, one initial block
    r = RefMiniAMR("--nx 10 --ny 10 --nz 10 --num_vars 40".split())
    r.init()
    assert f"{r.check_sum(0):.6f}" == "508.125470"
    assert f"{r.check_sum(39):.6f}" == "505.594636"
    m = oracle_from_ref(r)
    assert m.check_sum(0) == r.check_sum(0)
    assert m.check_sum(39) == r.check_sum(39)


@pytest.mark.parametrize("args,stages", [
    # anisotropic blocks catch any axis mix-up; AMR sphere -> level boundaries
    (f"--nx 4 --ny 6 --nz 8 --num_vars 3 --num_refine 2 --max_blocks 600 {SPHERE}", 3),
    (f"--nx 6 --ny 4 --nz 4 --num_vars 5 --comm_vars 2 --num_refine 3 --max_blocks 2000 {SPHERE}", 2),
    (f"--nx 4 --ny 4 --nz 6 --num_vars 2 --num_refine 2 --max_blocks 600 --permute {SPHERE}", 7),
    # uniform 27-point, several initial blocks (edges/corners must propagate X->Y->Z)
    ("--nx 4 --ny 6 --nz 4 --num_vars 3 --stencil 27 --uniform_refine 1 --num_refine 1 "
     "--init_x 2 --init_y 1 --init_z 2 --max_blocks 100", 3),
    ("--nx 4 --ny 4 --nz 6 --num_vars 2 --stencil 27 --uniform_refine 1 --num_refine 2 "
     "--max_blocks 100 --permute", 7),
    ("--nx 6 --ny 4 --nz 4 --num_vars 4 --comm_vars 3 --stencil 7 --uniform_refine 1 --num_refine 1 "
     "--init_x 1 --init_y 2 --init_z 3 --max_blocks 100", 2),
])
def test_stage_loop_bit_exact(args, stages):
    r = RefMiniAMR(args.split())
    r.init()
    r.refine(0)
    m = oracle_from_ref(r)
    p = r.p
    for stage in range(stages):
        for start in range(0, p["num_vars"], p["comm_vars"]):
            num = min(p["comm_vars"], p["num_vars"] - start)
            r.comm(start, num, stage)
            m.comm(start, num, stage)
            compare_all(r, m, f"after comm stage {stage} start {start}")
            for v in range(start, start + num):
                r.stencil_driver(v, stage)
                m.stencil_driver(v, stage)
        compare_all(r, m, f"after stage {stage}")
        for v in range(p["num_vars"]):
            assert m.check_sum(v) == r.check_sum(v)


def test_comm_counters_match():
    r = RefMiniAMR(f"--nx 4 --ny 4 --nz 4 --num_vars 2 --num_refine 2 --max_blocks 600 {SPHERE}".split())
    r.init(); r.refine(0)
    m = oracle_from_ref(r)
    r.comm(0, 2, 0)
    c = m.comm(0, 2, 0)
    rc = r.counters()
    assert list(c[:, 0]) == rc["same"] and list(c[:, 1]) == rc["diff"] and list(c[:, 2]) == rc["bc"]


def _by_geometry(ref):
    out = {}
    for s in ref.sorted_slots():
        b = ref.block(int(s))
        out[(b["level"],) + tuple(int(c) for c in b["cen"])] = int(s)
    return out


def test_split_and_consolidate_bit_exact():
    """refine() with a moving object: every block that appears must be either an
    octant of a split parent (block.c:161-173) or the 8-sum of consolidated
    children (block.c:418-430)."""
    num_refine = 2
    r = RefMiniAMR(f"--nx 4 --ny 6 --nz 8 --num_vars 3 --num_refine {num_refine} --block_change 1 "
                   f"--max_blocks 3000 --refine_freq 1 {MOVING}".split())
    r.init()
    r.refine(0)
    m = oracle_from_ref(r)
    n_split = n_cons = 0
    for ts in range(1, 9):
        r.stage(ts); m.stage(ts)
        compare_all(r, m, f"ts {ts}")
        before = _by_geometry(r)
        old = {k: r.get_slot(s) for k, s in before.items()}
        r.move(1.0)
        r.refine(ts)
        after = _by_geometry(r)
        scratch = oracle_from_ref(r)          # only used as a 10-slot workspace
        for key, slot in after.items():
            new = r.get_slot(slot)
            lev, cx, cy, cz = key
            if key in old:
                assert_bits_equal(new[:, 1:-1, 1:-1, 1:-1], old[key][:, 1:-1, 1:-1, 1:-1], "unchanged block")
                continue
            h = 2 ** (num_refine - lev)       # child half-size in mesh units (block.c:151-156)
            parent_key = None
            for o in range(8):
                pk = (lev - 1, cx - (2 * (o % 2) - 1) * h, cy - (2 * ((o // 2) % 2) - 1) * h,
                      cz - (2 * (o // 4) - 1) * h)
                if pk in old:
                    parent_key, octant = pk, o
            if parent_key is not None:        # I am a child of a split parent
                scratch.data[0] = old[parent_key]
                scratch.split_block(0, np.arange(1, 9))
                assert_bits_equal(new[:, 1:-1, 1:-1, 1:-1],
                                  scratch.data[1 + octant][:, 1:-1, 1:-1, 1:-1], "split child")
                n_split += 1
            else:                             # I am a consolidated parent
                hh = h // 2
                for o in range(8):
                    ck = (lev + 1, cx + (2 * (o % 2) - 1) * hh, cy + (2 * ((o // 2) % 2) - 1) * hh,
                          cz + (2 * (o // 4) - 1) * hh)
                    scratch.data[1 + o] = old[ck]
                scratch.consolidate_block(np.arange(1, 9), 0)
                assert_bits_equal(new[:, 1:-1, 1:-1, 1:-1], scratch.data[0][:, 1:-1, 1:-1, 1:-1],
                                  "consolidated parent")
                n_cons += 1
        sync_oracle(r, m)
        m_new = oracle_from_ref(r)
        m.data[:] = m_new.data
    assert n_split > 0 and n_cons > 0, (n_split, n_cons)


@pytest.mark.parametrize("stencil", [7, 27])
def test_pack_unpack_face_all_cases(stencil):
    r = RefMiniAMR(f"--nx 4 --ny 6 --nz 8 --num_vars 3 --stencil {stencil} --init_x 2 --max_blocks 10".split())
    r.init()
    rng = np.random.default_rng(5)
    for s in (0, 1):
        r.set_slot(s, rng.random((3, 6, 8, 10)))
    m = oracle_from_ref(r)
    for d in range(3):
        for case in list(range(10)) + list(range(10, 20)):
            for start, num in ((0, 3), (1, 2)):
                a = r.pack_face(0, case, d, start, num)
                b = m.pack_face(0, case, d, start, num)
                assert_bits_equal(a, b, f"pack dir {d} case {case}")
                r.unpack_face(a, 1, case, d, start, num)
                n = m.unpack_face(a, 1, case, d, start, num)
                assert n == len(a)
                assert_bits_equal(r.get_slot(1), m.data[1], f"unpack dir {d} case {case}")


def test_pack_unpack_block_payload():
    r = RefMiniAMR("--nx 4 --ny 6 --nz 8 --num_vars 3 --init_x 2 --max_blocks 10".split())
    r.init()
    m = oracle_from_ref(r)
    msg = r.pack_block(0)
    assert_bits_equal(msg[50:], m.pack_block(0), "pack_block payload (pack.c:66-70)")
    m.unpack_block(1, msg[50:])
    assert_bits_equal(m.data[1][:, 1:-1, 1:-1, 1:-1], m.data[0][:, 1:-1, 1:-1, 1:-1], "unpack_block")
