"""Host-side logic of the fused path, on the CPU: the halo plan (csrc/plan.cu,
through the C ABI's host-only mamr_plan_* calls) executed with numpy must give
exactly the ghost cells the oracle's three-phase comm() gives — on refined,
uniform, anisotropic and --permute meshes — and, at world_size 2 over gloo, the
pack / message / receive-buffer layout must reproduce a single-rank run of the
same global mesh."""
import os
import sys

import numpy as np
import pytest

from goldenutil import NAMES, Golden
from miniamr_b200 import build
from miniamr_b200.capi import HaloPlan, MamrError
from oracle.oracle import OracleMesh
from planexec import from_pool, run_halo, run_pack, to_pool

build.build()


def oracle_from_golden(g):
    m = OracleMesh(g.nx, g.ny, g.nz, g.num_vars, g.max_blocks, stencil=g.stencil,
                   comm_vars=g.comm_vars, permute=g.permute)
    m.set_topology(g.slots, g.level, g.nei_level, g.nei)
    g.apply_stencil0(m)
    for s, tiles in g.seeded_blocks():
        m.data[s] = tiles
    return m


# (the *_v40 / *_v160 fixtures share their topologies with cfg1_like / cfg2_like; the plan does not
# depend on the number of variables, and the numpy executor would take minutes on them)
@pytest.mark.parametrize("name", [n for n in NAMES if n not in ("cfg3_like_ring", "ring27", "cfg1_v40", "cfg2_v40",
                                                                "cfg3_v40", "cfg5_v160")])
def test_plan_equals_three_phase_comm(name):
    g = Golden(name)
    m = oracle_from_golden(g)
    used = int(max(g.slots)) + 1
    for stage in range(min(g.stages, 6 if g.permute else 2)):
        plan = HaloPlan(g.nx, g.ny, g.nz, g.num_vars, g.max_blocks, g.slots, g.level, g.nei_level,
                        g.nei, stencil=g.stencil, comm_vars=g.comm_vars, permute=g.permute,
                        stage=stage)
        pool = to_pool(m.data[:used], g.nx, g.ny, g.nz)
        out = run_halo(plan, g.slots, pool, [None]*3, 0, g.num_vars, g.nx, g.ny, g.nz)
        got = from_pool(out, used, g.nx, g.ny, g.nz)
        m.comm(0, g.num_vars, stage)
        for s in g.slots:
            bad = got[s].view(np.uint64) != m.data[s].view(np.uint64)
            assert not bad.any(), f"{name} stage {stage} slot {s}: first {np.argwhere(bad)[0]}"
        for v in range(g.num_vars):
            m.stencil_driver(v, stage)


def test_every_ghost_cell_has_exactly_one_op():
    g = Golden("amr7_moved_permute")
    plan = HaloPlan(g.nx, g.ny, g.nz, g.num_vars, g.max_blocks, g.slots, g.level, g.nei_level, g.nei,
                    stencil=7)
    tile = (g.nx + 2)*(g.ny + 2)*(g.nz + 2)
    from planexec import _dst_index
    for a in range(len(g.slots)):
        cover = np.zeros(tile, int)
        for op in plan.halo[plan.begin[a]:plan.begin[a + 1]]:
            np.add.at(cover, _dst_index(op).ravel(), 1)
        c = cover.reshape(g.nx + 2, g.ny + 2, g.nz + 2)
        assert (c[1:-1, 1:-1, 1:-1] == 0).all()
        c[1:-1, 1:-1, 1:-1] = 1
        assert (c == 1).all()


def test_unsupported_chain_is_reported():
    """27-point on a refined mesh: ghosts would have to travel through a level
    boundary (the reference itself rejects it, main.c:709-710) -> no plan."""
    g = Golden("amr7_aniso")
    with pytest.raises(MamrError, match="level boundary|7-point"):
        HaloPlan(g.nx, g.ny, g.nz, g.num_vars, g.max_blocks, g.slots, g.level, g.nei_level, g.nei,
                 stencil=27)


def test_misconnected_mesh_is_reported():
    nl = np.full((1, 6), 5, np.int32)
    with pytest.raises(MamrError, match="misconnected"):
        HaloPlan(4, 4, 4, 1, 4, [0], [0], nl, np.zeros((1, 6, 2, 2), np.int32))


# ---- world_size 2 over gloo -------------------------------------------------
def _block_data(seed, gx, gy, gz, shape):
    rs = np.random.RandomState((seed*1000003 + gx*10007 + gy*101 + gz) % (2**31 - 1))
    return rs.random_sample(shape)


def _rank_main(rank, world, port, cfg, q):
    import torch
    import torch.distributed as dist
    from miniamr_b200.mesh import rank_coords, uniform_mesh
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        npx, npy, npz = cfg["np"]
        nx, ny, nz = cfg["n"]
        bx, by, bz = cfg["b"]
        V, stencil, stages, permute = cfg["vars"], cfg["stencil"], cfg["stages"], cfg["permute"]
        cv = cfg["comm_vars"] or V
        shape = (V, nx + 2, ny + 2, nz + 2)
        nb = bx*by*bz
        top = uniform_mesh(bx, by, bz, npx, npy, npz, rank, nx, ny, nz, comm_vars=cv, stencil=stencil)
        rx, ry, rz = rank_coords(rank, npx, npy, npz)
        loc = OracleMesh(nx, ny, nz, V, nb, stencil=stencil)   # storage + stencil only
        loc.slots = top["slots"]
        for s in range(nb):
            lx, ly, lz = s % bx, (s//bx) % by, s//(bx*by)
            loc.data[s] = _block_data(cfg["seed"], rx*bx + lx, ry*by + ly, rz*bz + lz, shape)
        GX, GY, GZ = bx*npx, by*npy, bz*npz
        gtop = uniform_mesh(GX, GY, GZ)
        glob = OracleMesh(nx, ny, nz, V, GX*GY*GZ, stencil=stencil, comm_vars=cfg["comm_vars"],
                          permute=permute)
        glob.set_topology(gtop["slots"], gtop["level"], gtop["nei_level"], gtop["nei"])
        for s in range(GX*GY*GZ):
            glob.data[s] = _block_data(cfg["seed"], s % GX, (s//GX) % GY, s//(GX*GY), shape)
        dirs = top["dirs"]
        size = [int(max([0] + [o + z for o, z in zip(D["send_off"][D["index"]], D["send_size"])]))
                if len(D["partner"]) else 0 for D in dirs]
        nbytes = 0
        for st in range(stages):
            plan = HaloPlan(nx, ny, nz, V, nb, top["slots"], top["level"], top["nei_level"],
                            top["nei"], dirs=dirs, stencil=stencil, comm_vars=cfg["comm_vars"],
                            permute=permute, stage=st, rank=rank, num_ranks=world)
            for start in range(0, V, cv):
                num = min(cv, V - start)
                pool = to_pool(loc.data, nx, ny, nz)
                send = [np.zeros(max(z, 1)) for z in size]
                recv = [np.zeros(max(z, 1)) for z in size]
                for o in range(3):
                    d = plan.dirs[o]
                    D = dirs[d]
                    if not len(D["partner"]):
                        continue
                    run_pack(plan.pack[o], pool, send, recv, start, num)
                    for i, p in enumerate(D["partner"]):     # one message per partner (comm.c:146)
                        so, ro = D["send_off"][D["index"][i]], D["recv_off"][D["index"][i]]
                        out = torch.from_numpy(send[d][so:so + D["send_size"][i]].copy())
                        inn = torch.zeros(int(D["recv_size"][i]), dtype=torch.float64)
                        if rank < p:
                            dist.send(out, int(p), tag=d); dist.recv(inn, int(p), tag=d)
                        else:
                            dist.recv(inn, int(p), tag=d); dist.send(out, int(p), tag=d)
                        recv[d][ro:ro + D["recv_size"][i]] = inn.numpy()
                        nbytes += out.numel()*8
                newp = run_halo(plan, top["slots"], pool, recv, start, num, nx, ny, nz)
                loc.data[:] = from_pool(newp, nb, nx, ny, nz)
                for v in range(start, start + num):
                    loc.stencil_driver(v, st)
            glob.stage(st)
        for s in range(nb):
            lx, ly, lz = s % bx, (s//bx) % by, s//(bx*by)
            gs = (rx*bx + lx) + GX*((ry*by + ly) + GY*(rz*bz + lz))
            bad = loc.data[s].view(np.uint64) != glob.data[gs].view(np.uint64)
            assert not bad.any(), f"rank {rank} slot {s}: first {np.argwhere(bad)[0]}"
        assert nbytes > 0
        q.put((rank, "ok"))
    except Exception as e:      # noqa: BLE001
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


GLOO_CASES = [
    dict(np=[2, 1, 1], n=[4, 6, 8], b=[2, 2, 2], vars=3, stencil=7, stages=3, seed=1, comm_vars=0, permute=0),
    dict(np=[2, 1, 1], n=[4, 6, 4], b=[2, 3, 2], vars=2, stencil=27, stages=3, seed=2, comm_vars=0, permute=0),
    dict(np=[1, 2, 1], n=[6, 4, 4], b=[2, 2, 3], vars=4, stencil=27, stages=3, seed=3, comm_vars=3, permute=0),
    dict(np=[1, 1, 2], n=[4, 4, 6], b=[3, 2, 2], vars=2, stencil=27, stages=7, seed=4, comm_vars=0, permute=1),
]


@pytest.mark.parametrize("case", range(len(GLOO_CASES)))
def test_two_ranks_over_gloo_match_single_rank_oracle(case):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + case
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, GLOO_CASES[case], q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
