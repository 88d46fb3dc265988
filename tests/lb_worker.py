"""TEST INFRASTRUCTURE — the ranks of a loopback run (tests/loopback.py), no torch.

    python tests/lb_worker.py KIND RANK WORLD DIR JSON_CFG     one rank, a process of its own
    python tests/lb_worker.py KIND all  WORLD DIR JSON_CFG     all ranks, one thread each

Processes: rank r runs on GPU r % device_count; window handles and barriers go through files
in a scratch directory, windows are mapped with CUDA IPC, and on one GPU the ranks' kernels are
time-sliced (correct but slow: a rank that waits for a peer burns its whole time slice).
Threads: the ranks share GPU 0 and one CUDA context, their kernels run concurrently, windows are plain
pointers.  A fresh process per run keeps every stream on a hardware queue of its own
(CUDA_DEVICE_MAX_CONNECTIONS=32), so that a kernel spinning on a peer's flag never sits in front
of that peer's work, and CUDA_MODULE_LOADING=EAGER keeps first launches from waiting for the
device while a peer already spins.
"""
import json
import os
import sys
import threading
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class Host:
    def __init__(self, rank, world, d):
        self.rank, self.world, self.dir, self.n = rank, world, d, 0

    def _wait(self, names, timeout=180):
        t0 = time.time()
        while not all(os.path.exists(os.path.join(self.dir, n)) for n in names):
            if time.time() - t0 > timeout:
                raise TimeoutError(f"rank {self.rank}: peers did not reach {names[0][:-2]}")
            time.sleep(0.002)

    def barrier(self):
        self.n += 1
        open(os.path.join(self.dir, f"bar{self.n}_{self.rank}"), "w").close()
        self._wait([f"bar{self.n}_{r}" for r in range(self.world)])

    def allgather(self, tag, blob):
        tmp = os.path.join(self.dir, f"{tag}_{self.rank}.tmp")
        with open(tmp, "wb") as f:
            f.write(blob)
        os.rename(tmp, os.path.join(self.dir, f"{tag}_{self.rank}"))
        self._wait([f"{tag}_{r}" for r in range(self.world)])
        return [open(os.path.join(self.dir, f"{tag}_{r}"), "rb").read() for r in range(self.world)]


class ThreadHost:
    """the same for ranks that are threads of this process"""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self.sh = rank, world, shared

    def barrier(self):
        self.sh["bar"].wait()

    def allgather(self, tag, blob):
        self.sh.setdefault(tag, {})[self.rank] = blob
        self.sh["bar"].wait()
        out = [self.sh[tag][r] for r in range(self.world)]
        self.sh["bar"].wait()
        return out


THREADS = False      # all ranks are threads of this process: they share GPU 0


def device_for(rank):
    from miniamr_b200.capi import load_library
    n = load_library().mamr_device_count()
    assert n > 0, "no CUDA device"
    return 0 if THREADS else rank % n


def connect(d, host):
    d.p2p_connect(host.allgather("handle", d.p2p_handle()))
    host.barrier()


def block_data(seed, gx, gy, gz, shape):
    rs = np.random.RandomState((seed*1000003 + gx*10007 + gy*101 + gz) % (2**31 - 1))
    return rs.random_sample(shape)


def uniform(rank, world, host, cfg):
    """tests/mgpu_worker.py without torch: this rank's sub-cube of a uniform global mesh against
    the CPU oracle run on the whole (single-rank) mesh, bit for bit."""
    from miniamr_b200.capi import DeviceMesh
    from miniamr_b200.mesh import rank_coords, uniform_mesh
    from oracle.oracle import OracleMesh

    npx, npy, npz = cfg["np"]
    assert npx*npy*npz == world
    nx, ny, nz = cfg["n"]
    bx, by, bz = cfg["b"]
    V, stencil, stages = cfg["vars"], cfg["stencil"], cfg["stages"]
    comm_vars, permute = cfg.get("comm_vars", 0), cfg.get("permute", 0)
    cv = comm_vars if 0 < comm_vars <= V else V
    shape = (V, nx + 2, ny + 2, nz + 2)
    nb = bx*by*bz
    top = uniform_mesh(bx, by, bz, npx, npy, npz, rank, nx, ny, nz, comm_vars=cv, stencil=stencil)
    d = DeviceMesh(nx, ny, nz, V, nb, stencil=stencil, comm_vars=comm_vars, permute=permute,
                   device=device_for(rank), rank=rank, num_ranks=world)
    d.set_topology(top["slots"], top["level"], top["nei_level"], top["nei"])
    connect(d, host)
    d.set_comm_lists(top["dirs"])
    rx, ry, rz = rank_coords(rank, npx, npy, npz)
    for s in range(nb):
        lx, ly, lz = s % bx, (s//bx) % by, s//(bx*by)
        d.upload_block(s, block_data(cfg["seed"], rx*bx + lx, ry*by + ly, rz*bz + lz, shape))

    GX, GY, GZ = bx*npx, by*npy, bz*npz
    gtop = uniform_mesh(GX, GY, GZ)
    m = OracleMesh(nx, ny, nz, V, GX*GY*GZ, stencil=stencil, comm_vars=comm_vars, permute=permute)
    m.set_topology(gtop["slots"], gtop["level"], gtop["nei_level"], gtop["nei"])
    for s in range(GX*GY*GZ):
        m.data[s] = block_data(cfg["seed"], s % GX, (s//GX) % GY, s//(GX*GY), shape)

    def compare(what):
        for s in range(nb):
            lx, ly, lz = s % bx, (s//bx) % by, s//(bx*by)
            gs = (rx*bx + lx) + GX*((ry*by + ly) + GY*(rz*bz + lz))
            bad = d.download_block(s).view(np.uint64) != m.data[gs].view(np.uint64)
            if bad.any():
                raise AssertionError(f"rank {rank} {what}: slot {s}: {int(bad.sum())} cells differ, "
                                     f"first {np.argwhere(bad)[0]}")

    for st in range(stages):
        for start in range(0, V, cv):
            num = min(cv, V - start)
            d.comm(start, num, st)
            if cfg.get("check_comm") and st == 0:
                m.comm(start, num, st)
                compare(f"comm stage {st}")        # forces the split (materialised) path
                for v in range(start, start + num):
                    d.stencil_driver(v, st)
                    m.stencil_driver(v, st)
                continue
            for v in range(start, start + num):
                d.stencil_driver(v, st)
        if not (cfg.get("check_comm") and st == 0):
            m.stage(st)
        sums = [d.check_sum(v) for v in range(V)]
        want = [m.check_sum(v) for v in range(V)]
        assert np.allclose(sums, want, rtol=1e-13, atol=0), (rank, st, sums, want)
    compare("final")
    c = d.counters()
    assert sum(c["size_mesg_send"]) > 0 and sum(c["counter_face_recv"]) > 0
    d.sync()
    host.barrier()
    d.close()


def migration(rank, world, host, cfg):
    """mamr_stage_send_block / _recv_block / mamr_flush_block_moves over the windows: every rank
    sends two blocks to the next rank and one to the one after it while half of the variables
    live in the second pool; the moved blocks then take part in one more stage."""
    from miniamr_b200.capi import DeviceMesh
    from oracle.oracle import OracleMesh

    (nx, ny, nz), V, MB = cfg["n"], cfg["vars"], 12
    shape = (V, nx + 2, ny + 2, nz + 2)

    def seed(r, slot):
        return np.random.RandomState(1000*r + slot + 7).random_sample(shape)

    def isolated(slots):
        n = len(slots)
        return (np.asarray(slots, np.int32), np.zeros(n, np.int32), np.full((n, 6), -2, np.int32),
                np.zeros((n, 6, 2, 2), np.int32))

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
    m = orc[rank]
    for dst, (src_rank, src_slot) in {6: ((rank - 1) % world, 0), 7: ((rank - 1) % world, 2),
                                      8: ((rank - 2) % world, 4)}.items():
        m.data[dst][:, 1:-1, 1:-1, 1:-1] = before[src_rank][src_slot][:, 1:-1, 1:-1, 1:-1]
    m.set_topology(*isolated([1, 3, 5, 6, 7, 8]))
    m.stage(1)

    d = DeviceMesh(nx, ny, nz, V, MB, device=device_for(rank), rank=rank, num_ranks=world)
    d.set_topology(*isolated(range(6)))
    connect(d, host)
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
        bad = d.download_block(s).view(np.uint64) != m.data[s].view(np.uint64)
        assert not bad.any(), f"rank {rank} slot {s}: {int(bad.sum())} cells differ, first {np.argwhere(bad)[0]}"
    assert d.counters()["migrate_bytes"] == 3*V*nx*ny*nz*8
    d.flush_block_moves()                     # a round without moves on any rank
    d.sync()
    host.barrier()
    d.close()


if __name__ == "__main__":
    kind, who, world, scratch, cfg = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], json.loads(sys.argv[5])
    fn = {"uniform": uniform, "migration": migration}[kind]
    if who != "all":
        fn(int(who), world, Host(int(who), world, scratch), cfg)
        print(f"LB_OK {who}")
        sys.exit(0)
    THREADS = True
    shared = {"bar": threading.Barrier(world, timeout=120)}
    errors = []

    def main(rank):
        try:
            fn(rank, world, ThreadHost(rank, world, shared), cfg)
        except BaseException:      # noqa: BLE001 -- every rank's failure is reported
            errors.append((rank, traceback.format_exc()))
            shared["bar"].abort()

    ts = [threading.Thread(target=main, args=(r,), daemon=True) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(240)
    if any(t.is_alive() for t in ts):
        print("rank thread(s) still running", file=sys.stderr)
        os._exit(3)
    first = [e for e in errors if "BrokenBarrierError" not in e[1]] or errors
    if first:
        print("rank %d:\n%s" % first[0], file=sys.stderr)
        os._exit(2)
    print("LB_OK all")
