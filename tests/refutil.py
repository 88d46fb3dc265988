"""Shared helpers for tests that talk to the compiled reference (oracle/_ref)."""
import numpy as np
import pytest

from oracle import refharness
from oracle.oracle import OracleMesh

needs_ref = pytest.mark.skipif(not refharness.available("ref"),
                               reason="oracle/_ref/libminiamr_ref.so not built")

SPHERE = "--num_objects 1 --object 2 0 0.3 0.3 0.3 0.01 0.01 0.01 0.25 0.25 0.25 0 0 0"
# sphere that moves fast enough to force both refinement and coarsening
MOVING = "--num_objects 1 --object 2 0 0.2 0.2 0.2 0.09 0.07 0.05 0.2 0.2 0.2 0 0 0"


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bits_equal(a, b, what=""):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = bits(a) != bits(b)
    if bad.any():
        idx = np.argwhere(bad.reshape(a.shape))[0]
        raise AssertionError(f"{what}: {int(bad.sum())} of {a.size} values differ; "
                             f"first at {tuple(idx)}: {a[tuple(idx)]!r} vs {b[tuple(idx)]!r}")


def oracle_from_ref(ref, permute=0):
    """OracleMesh holding a bit-copy of the reference's active blocks
    (ghost cells included, whatever they contain) and its topology."""
    p = ref.p
    m = OracleMesh(p["nx"], p["ny"], p["nz"], p["num_vars"], p["max_blocks"],
                   stencil=p["stencil"], comm_vars=p["comm_vars"], permute=p["permute"])
    sync_oracle(ref, m)
    return m


def sync_oracle(ref, m):
    slots, lev, nl, ne = ref.topology()
    m.set_topology(slots, lev, nl, ne)
    for s in slots:
        m.data[s] = ref.get_slot(int(s))


def compare_all(ref, m, what=""):
    for s in ref.sorted_slots():
        assert_bits_equal(ref.get_slot(int(s)), m.data[s], f"{what} slot {s}")
