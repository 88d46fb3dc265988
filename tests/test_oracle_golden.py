"""The CPU oracle against the committed golden vectors (generated from the
unmodified reference by tests/golden/make_golden.py).  Runs without
/root/reference."""
import numpy as np
import pytest

from goldenutil import NAMES, Golden, digest
from oracle.oracle import OracleMesh


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    g = Golden(name)
    m = OracleMesh(g.nx, g.ny, g.nz, g.num_vars, g.max_blocks, stencil=g.stencil,
                   comm_vars=g.comm_vars, permute=g.permute)
    m.set_topology(g.slots, g.level, g.nei_level, g.nei)
    g.apply_stencil0(m)
    for s, tiles in g.seeded_blocks():
        m.data[s] = tiles
    for st in range(g.stages):
        m.stage(st)
        for v in range(g.num_vars):
            assert m.check_sum(v) == g.check_sums[st, v], (st, v)   # bit-exact
    assert digest(m.data[s] for s in g.slots) == g.sha256


def test_golden_set_is_complete():
    assert {"amr7_aniso", "amr7_moved_permute", "uni27_aniso", "uni27_permute", "cfg1_like",
            "cfg2_like", "cfg3_like_ring", "ring27", "uni0_variable_work",
            "cfg2_v40", "cfg3_v40", "cfg5_v160", "cfg1_v40"} <= set(NAMES)
