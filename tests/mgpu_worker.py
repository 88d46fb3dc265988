"""Worker of tests/test_multi_gpu.py: run under torch.distributed.run with N ranks,
one GPU each.  Every rank owns a sub-cube of a uniform global mesh (sharded as the
reference shards blocks over MPI ranks, init.c:156-190), runs the stage loop
through the C ABI with NCCL ghost exchange and compares ITS blocks bit for bit
with the CPU oracle run on the whole (single-rank) global mesh."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def block_data(seed, gx, gy, gz, shape):
    rs = np.random.RandomState((seed*1000003 + gx*10007 + gy*101 + gz) % (2**31 - 1))
    return rs.random_sample(shape)


def main():
    import torch
    import torch.distributed as dist
    from miniamr_b200.capi import DeviceMesh
    from miniamr_b200.mesh import rank_coords, uniform_mesh
    from oracle.oracle import OracleMesh

    cfg = json.loads(sys.argv[1])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    npx, npy, npz = cfg["np"]
    assert npx*npy*npz == world
    nx, ny, nz = cfg["n"]
    bx, by, bz = cfg["b"]
    V, stencil, stages = cfg["vars"], cfg["stencil"], cfg["stages"]
    comm_vars, permute = cfg.get("comm_vars", 0), cfg.get("permute", 0)
    cv = comm_vars if 0 < comm_vars <= V else V
    shape = (V, nx + 2, ny + 2, nz + 2)

    # ---- this rank's shard on the device ---------------------------------
    top = uniform_mesh(bx, by, bz, npx, npy, npz, rank, nx, ny, nz, comm_vars=cv, stencil=stencil)
    nb = bx*by*bz
    d = DeviceMesh(nx, ny, nz, V, nb, stencil=stencil, comm_vars=comm_vars, permute=permute,
                   device=local, rank=rank, num_ranks=world)
    d.set_topology(top["slots"], top["level"], top["nei_level"], top["nei"])
    if os.environ.get("MAMR_TRANSPORT", "p2p") == "nccl":
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(DeviceMesh.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        d.nccl_init(bytes(uid.cpu().numpy().tobytes()))
    else:
        # peer-memory transport: window handles all-gathered over the host channel (CUDA IPC)
        hs = [None]*world
        dist.all_gather_object(hs, d.p2p_handle())
        d.p2p_connect(hs)
    d.set_comm_lists(top["dirs"])
    rx, ry, rz = rank_coords(rank, npx, npy, npz)
    for s in range(nb):
        lx, ly, lz = s % bx, (s//bx) % by, s//(bx*by)
        d.upload_block(s, block_data(cfg["seed"], rx*bx + lx, ry*by + ly, rz*bz + lz, shape))

    # ---- the oracle on the whole mesh, one rank ----------------------------
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
            got = d.download_block(s)
            want = m.data[gs]
            # ghost cells facing another rank hold the same values in both runs; compare all
            bad = got.view(np.uint64) != want.view(np.uint64)
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
    d.close()
    dist.barrier()
    if rank == 0:
        print("MGPU_OK", json.dumps({"nvlink_bytes": sum(c["size_mesg_send"]),
                                     "launches": c["kernel_launches"]}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
