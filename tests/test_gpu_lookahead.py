"""Look-ahead over a comm group (api.cu: flush_pending): the reference's driver interleaves
stencil_driver(v) and check_sum(v) on checksum stages (driver.c:85-103); the device computes
the whole group at the first flush and commits variable by variable.  Whatever the host does
in between -- checksums of not-yet-committed variables, a second comm(), downloads, skipping
variables -- must give exactly what the reference gives."""
import numpy as np
import pytest

from goldenutil import Golden
from oracle.oracle import OracleMesh

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def pair(name):
    from miniamr_b200.capi import DeviceMesh
    g = Golden(name)
    d = DeviceMesh(g.nx, g.ny, g.nz, g.num_vars, g.max_blocks, stencil=g.stencil,
                   comm_vars=g.comm_vars, permute=g.permute)
    m = OracleMesh(g.nx, g.ny, g.nz, g.num_vars, g.max_blocks, stencil=g.stencil,
                   comm_vars=g.comm_vars, permute=g.permute)
    for x in (d, m):
        x.set_topology(g.slots, g.level, g.nei_level, g.nei)
    for s, tiles in g.seeded_blocks():
        d.upload_block(s, tiles)
        m.data[s] = tiles
    return g, d, m


def same(g, d, m, what):
    for s in g.slots:
        bad = bits(d.download_block(int(s))) != bits(m.data[s])
        assert not bad.any(), f"{what}: slot {s}: first {np.argwhere(bad)[0]}"


@pytest.mark.parametrize("name", ["cfg2_like", "cfg1_like", "amr7_moved_permute", "uni27_permute", "amr7_aniso",
                                  "cfg1_v40", "cfg2_v40", "cfg5_v160"])
def test_interleaved_stencil_and_checksum(name):
    g, d, m = pair(name)
    V = g.num_vars
    for st in range(3):
        for start in range(0, V, g.comm_vars):
            num = min(g.comm_vars, V - start)
            d.comm(start, num, st)
            m.comm(start, num, st)
            for v in range(start, start + num):
                d.stencil_driver(v, st)
                m.stencil_driver(v, st)
                want = m.check_sum(v)
                assert abs(d.check_sum(v) - want) <= 1e-13*abs(want), (st, v)
                if v + 1 < start + num:
                    # a variable that has been looked ahead but not committed still shows its OLD state
                    want = m.check_sum(v + 1)
                    assert abs(d.check_sum(v + 1) - want) <= 1e-13*abs(want), (st, v, "next")
    same(g, d, m, name)
    d.close()


@pytest.mark.parametrize("name", ["amr7_aniso", "uni27_aniso", "amr7_moved_permute"])
def test_lookahead_result_is_dropped_when_the_host_changes_course(name):
    g, d, m = pair(name)
    if g.comm_vars < g.num_vars:      # one comm group spanning every variable for this scenario
        pytest.skip("needs a single comm group")
    V = g.num_vars
    assert V >= 3
    # stage 0: stencil of variable 0 only, then a download (forces the other variables' comm real)
    d.comm(0, V, 0); m.comm(0, V, 0)
    d.stencil_driver(0, 0); m.stencil_driver(0, 0)
    assert abs(d.check_sum(0) - m.check_sum(0)) <= 1e-13*abs(m.check_sum(0))
    same(g, d, m, "after download")           # variables 1.. hold comm() ghosts, old interiors
    # their stencil now runs on the materialised tiles
    for v in range(1, V):
        d.stencil_driver(v, 0); m.stencil_driver(v, 0)
    same(g, d, m, "stage 0")
    # stage 1: variable 0, then a SECOND comm() of the whole group before the others are asked for
    d.comm(0, V, 1); m.comm(0, V, 1)
    d.stencil_driver(0, 1); m.stencil_driver(0, 1)
    d.check_sum(0)
    d.comm(0, V, 1); m.comm(0, V, 1)
    for v in range(V):
        d.stencil_driver(v, 1); m.stencil_driver(v, 1)
    same(g, d, m, "stage 1")
    # stage 2: skip a variable altogether, upload new data into it afterwards
    d.comm(0, V, 2); m.comm(0, V, 2)
    for v in range(V):
        if v != 1:
            d.stencil_driver(v, 2); m.stencil_driver(v, 2)
            d.check_sum(v)
    s0 = int(g.slots[0])
    t = m.data[s0].copy(); t[1] += 1.0
    d.upload_block(s0, t); m.data[s0] = t
    same(g, d, m, "stage 2")
    d.close()
