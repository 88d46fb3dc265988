"""The off-rank path on ONE GPU (runs under the driver's single-GPU `pytest -m gpu`): N ranks as
N contexts in this process, connected through the peer-memory transport (tests/loopback.py).
Same cases and same oracle comparison as tests/test_multi_gpu.py (which needs N GPUs): pack
kernels (boxop_kernel) reading resolved origins and earlier receive buffers, widened Y/Z faces
forwarded through up to three ranks, receive-buffer reads inside fused2 / slab7, staged comm
groups, --permute, and the all-reduce of check_sum (comm.c:254-401,1002-1150; check_sum.c:57)."""
import os

import pytest

from loopback import run_uniform_case
from test_multi_gpu import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [1, 0])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_ranks_on_one_gpu_match_single_rank_oracle(case, fused):
    old = os.environ.get("MAMR_NO_FUSED")
    os.environ["MAMR_NO_FUSED"] = "0" if fused else "1"
    try:
        out = run_uniform_case(CASES[case])
    finally:
        if old is None:
            os.environ.pop("MAMR_NO_FUSED", None)
        else:
            os.environ["MAMR_NO_FUSED"] = old
    assert len(out) == CASES[case]["np"][0]*CASES[case]["np"][1]*CASES[case]["np"][2]


BASELINE_SHAPES = [
    # BASELINE variable counts on 8 ranks: cfg2 (16^3, 40 vars, 27-pt), cfg3 (32^3, 40 vars, 7-pt),
    # cfg5 (10^3, 160 vars in four comm groups = four receive-buffer sets, 27-pt)
    dict(np=[2, 2, 2], n=[16, 16, 16], b=[2, 2, 2], vars=40, stencil=27, stages=3, seed=31),
    dict(np=[2, 2, 2], n=[32, 32, 32], b=[1, 1, 2], vars=40, stencil=7, stages=3, seed=32),
    dict(np=[2, 2, 2], n=[10, 10, 10], b=[2, 2, 2], vars=160, comm_vars=40, stencil=27, stages=3, seed=33),
    dict(np=[2, 2, 1], n=[10, 10, 10], b=[2, 2, 3], vars=40, stencil=7, stages=3, seed=34, permute=1),
]


@pytest.mark.parametrize("case", range(len(BASELINE_SHAPES)))
def test_baseline_variable_counts_on_eight_ranks(case):
    run_uniform_case(BASELINE_SHAPES[case])


def test_block_migration_between_ranks_on_one_gpu():
    """mamr_stage_send_block / mamr_stage_recv_block / mamr_flush_block_moves over the windows
    (pull): 4 ranks, every rank sends two blocks to the next rank and one to the one after it.
    Half of the variables live in the second pool when the blocks move (rcb.c:207-337 payloads,
    pack.c:66-70 layout); the moved blocks then take part in one more stage."""
    import numpy as np
    from loopback import Ranks, connect
    from miniamr_b200.capi import DeviceMesh
    from oracle.oracle import OracleMesh

    world, (nx, ny, nz), V, MB = 4, (4, 6, 8), 5, 12
    shape = (V, nx + 2, ny + 2, nz + 2)

    def seed(rank, slot):
        return np.random.RandomState(1000*rank + slot + 7).random_sample(shape)

    def isolated(slots):
        n = len(slots)
        return (np.asarray(slots, np.int32), np.zeros(n, np.int32), np.full((n, 6), -2, np.int32),
                np.zeros((n, 6, 2, 2), np.int32))

    # the oracle, rank by rank: stage 0 on slots 0..5 (variables 0..2 only), the moves, stage 1
    orc = []
    for r in range(world):
        m = OracleMesh(nx, ny, nz, V, MB)
        m.set_topology(*isolated(range(6)))
        for s in range(6):                # slots that were never active hold zeros (both pools)
            m.data[s] = seed(r, s)
        m.comm(0, 3, 0)
        for v in range(3):
            m.stencil_driver(v, 0)
        orc.append(m)
    before = [{s: m.data[s].copy() for s in range(MB)} for m in orc]
    for r in range(world):
        m = orc[r]
        for dst, (src_rank, src_slot) in {6: ((r - 1) % world, 0), 7: ((r - 1) % world, 2),
                                          8: ((r - 2) % world, 4)}.items():
            m.data[dst][:, 1:-1, 1:-1, 1:-1] = before[src_rank][src_slot][:, 1:-1, 1:-1, 1:-1]
        m.set_topology(*isolated([1, 3, 5, 6, 7, 8]))
        m.stage(1)

    def rank_main(rank, ctx):
        d = DeviceMesh(nx, ny, nz, V, MB, device=0, rank=rank, num_ranks=world)
        try:
            d.set_topology(*isolated(range(6)))
            connect(d, rank, ctx)
            for s in range(6):
                d.upload_block(s, seed(rank, s))
            d.comm(0, 3, 0)
            for v in range(3):
                d.stencil_driver(v, 0)                # variables 0..2 now live in the other pool
            nxt, nx2, prv, pr2 = (rank + 1) % world, (rank + 2) % world, (rank - 1) % world, (rank - 2) % world
            d.stage_send_block(0, nxt)
            d.stage_recv_block(6, prv)
            d.stage_send_block(2, nxt)
            d.stage_send_block(4, nx2)
            d.stage_recv_block(7, prv)
            d.stage_recv_block(8, pr2)
            assert d.pending_block_moves() == 6
            d.flush_block_moves()
            assert d.pending_block_moves() == 0
            d.sync()                                  # reports a wait that timed out
            d.set_topology(*isolated([1, 3, 5, 6, 7, 8]))
            d.stage(1)
            for s in (1, 3, 5, 6, 7, 8):
                bad = d.download_block(s).view(np.uint64) != orc[rank].data[s].view(np.uint64)
                assert not bad.any(), f"rank {rank} slot {s}: {int(bad.sum())} cells differ, first {np.argwhere(bad)[0]}"
            assert d.counters()["migrate_bytes"] == 3*V*nx*ny*nz*8
            d.flush_block_moves()                     # a round without moves on any rank
            ctx.barrier()
        finally:
            try:
                ctx.barrier()
            except Exception:
                pass
            d.close()

    Ranks(world).run(rank_main)
