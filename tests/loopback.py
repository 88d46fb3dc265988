"""TEST INFRASTRUCTURE — N ranks of the multi-GPU path on ONE GPU: one DeviceMesh per rank,
one host thread per rank, windows of the peer-memory transport connected inside the process
(miniamr_b200.h: mamr_p2p_get_handle / mamr_p2p_connect).  Everything the off-rank path does
on a real multi-GPU box runs here too -- pack kernels, pushes into the partner's receive
buffers, arrival / credit flags, receive-buffer reads of the fused kernels, the check_sum
all-reduce -- only the NVLink hop is a hop through local memory."""
import threading
import traceback

import numpy as np


def block_data(seed, gx, gy, gz, shape):
    rs = np.random.RandomState((seed*1000003 + gx*10007 + gy*101 + gz) % (2**31 - 1))
    return rs.random_sample(shape)


class Ranks:
    """run fn(rank, ctx) on `world` threads; ctx.barrier() is a host barrier, ctx.share is a dict"""

    def __init__(self, world, timeout=300):
        self.world = world
        self._bar = threading.Barrier(world, timeout=timeout)
        self.share = {}
        self.errors = []
        self.timeout = timeout

    def barrier(self):
        self._bar.wait()

    def run(self, fn):
        def main(rank):
            try:
                fn(rank, self)
            except BaseException:      # noqa: BLE001 -- report every rank's failure
                self.errors.append((rank, traceback.format_exc()))
                try:
                    self._bar.abort()
                except Exception:
                    pass
        ts = [threading.Thread(target=main, args=(r,), daemon=True) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(self.timeout)
        alive = [t for t in ts if t.is_alive()]
        assert not alive, f"{len(alive)} rank thread(s) still running after {self.timeout} s"
        first = [e for e in self.errors if "BrokenBarrierError" not in e[1]] or self.errors
        assert not self.errors, "rank %d:\n%s" % first[0]


def connect(d, rank, ctx, key="handles"):
    """exchange window handles between the rank threads and map them"""
    ctx.share.setdefault(key, {})[rank] = d.p2p_handle()
    ctx.barrier()
    d.p2p_connect([ctx.share[key][r] for r in range(ctx.world)])
    ctx.barrier()


def run_uniform_case(cfg, device=0):
    """tests/mgpu_worker.py on threads: every rank owns a sub-cube of a uniform global mesh and
    compares ITS blocks bit for bit with the CPU oracle run on the whole (single-rank) mesh."""
    from miniamr_b200.capi import DeviceMesh
    from miniamr_b200.mesh import rank_coords, uniform_mesh
    from oracle.oracle import OracleMesh

    npx, npy, npz = cfg["np"]
    world = npx*npy*npz
    nx, ny, nz = cfg["n"]
    bx, by, bz = cfg["b"]
    V, stencil, stages = cfg["vars"], cfg["stencil"], cfg["stages"]
    comm_vars, permute = cfg.get("comm_vars", 0), cfg.get("permute", 0)
    cv = comm_vars if 0 < comm_vars <= V else V
    shape = (V, nx + 2, ny + 2, nz + 2)
    nb = bx*by*bz
    check_comm = bool(cfg.get("check_comm"))

    # ---- the oracle on the whole mesh, one rank: snapshots the ranks compare with ------
    GX, GY, GZ = bx*npx, by*npy, bz*npz
    gtop = uniform_mesh(GX, GY, GZ)
    m = OracleMesh(nx, ny, nz, V, GX*GY*GZ, stencil=stencil, comm_vars=comm_vars, permute=permute)
    m.set_topology(gtop["slots"], gtop["level"], gtop["nei_level"], gtop["nei"])
    for s in range(GX*GY*GZ):
        m.data[s] = block_data(cfg["seed"], s % GX, (s//GX) % GY, s//(GX*GY), shape)
    snap_comm = {}       # start -> data right after comm(start) of stage 0
    sums = []
    for st in range(stages):
        if check_comm and st == 0:
            for start in range(0, V, cv):
                num = min(cv, V - start)
                m.comm(start, num, st)
                snap_comm[start] = {s: m.data[s].copy() for s in range(GX*GY*GZ)}
                for v in range(start, start + num):
                    m.stencil_driver(v, st)
        else:
            m.stage(st)
        sums.append([m.check_sum(v) for v in range(V)])
    final = {s: m.data[s] for s in range(GX*GY*GZ)}
    out = {}

    def rank_main(rank, ctx):
        top = uniform_mesh(bx, by, bz, npx, npy, npz, rank, nx, ny, nz, comm_vars=cv, stencil=stencil)
        d = DeviceMesh(nx, ny, nz, V, nb, stencil=stencil, comm_vars=comm_vars, permute=permute,
                       device=device, rank=rank, num_ranks=world)
        try:
            d.set_topology(top["slots"], top["level"], top["nei_level"], top["nei"])
            connect(d, rank, ctx)
            d.set_comm_lists(top["dirs"])
            rx, ry, rz = rank_coords(rank, npx, npy, npz)

            def gslot(s):
                lx, ly, lz = s % bx, (s//bx) % by, s//(bx*by)
                return (rx*bx + lx) + GX*((ry*by + ly) + GY*(rz*bz + lz))

            for s in range(nb):
                lx, ly, lz = s % bx, (s//bx) % by, s//(bx*by)
                d.upload_block(s, block_data(cfg["seed"], rx*bx + lx, ry*by + ly, rz*bz + lz, shape))
            ctx.barrier()

            def compare(what, want):
                for s in range(nb):
                    got = d.download_block(s)
                    bad = got.view(np.uint64) != want[gslot(s)].view(np.uint64)
                    if bad.any():
                        raise AssertionError(f"rank {rank} {what}: slot {s}: {int(bad.sum())} cells differ, "
                                             f"first {np.argwhere(bad)[0]}")

            for st in range(stages):
                for start in range(0, V, cv):
                    num = min(cv, V - start)
                    d.comm(start, num, st)
                    if check_comm and st == 0:
                        compare(f"comm stage {st} start {start}", snap_comm[start])   # materialised ghosts
                    for v in range(start, start + num):
                        d.stencil_driver(v, st)
                got = [d.check_sum(v) for v in range(V)]
                assert np.allclose(got, sums[st], rtol=1e-13, atol=0), (rank, st, got, sums[st])
            compare("final", final)
            c = d.counters()
            assert sum(c["size_mesg_send"]) > 0 and sum(c["counter_face_recv"]) > 0
            out[rank] = c
            ctx.barrier()
        finally:
            try:
                ctx.barrier()
            except Exception:
                pass
            d.close()

    Ranks(world).run(rank_main)
    return out
