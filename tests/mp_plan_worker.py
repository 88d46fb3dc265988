"""TEST INFRASTRUCTURE — one rank of the multi-rank planner check, started by

    minimpi/_bin/minimpirun -n N python tests/mp_plan_worker.py OUTDIR STAGES <miniAMR args>

The UNMODIFIED reference (ref_mp) runs its own driver() at N ranks (refinement,
load balancing, migration).  On the mesh it ends with, this rank hands the
reference's topology and off-rank comm lists to the host-only halo planner of
the C ABI (csrc/plan.cu), executes the resulting pack ops and halo plan with
numpy (tests/planexec.py — the descriptors the CUDA kernels consume), sends the
packed messages to the partner ranks over the host channel exactly as
comm.c:71-84,120-151 does, and compares every ghost cell with what the
reference's own comm() produces from the same state.  No GPU involved."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refharness  # noqa: E402
from miniamr_b200.capi import HaloPlan  # noqa: E402
from planexec import from_pool, run_halo, run_pack, to_pool  # noqa: E402


def main():
    outdir, stages, args = sys.argv[1], int(sys.argv[2]), sys.argv[3:]
    r = refharness.RefMiniAMR(args, variant="ref_mp", run_driver=True)
    p = r.p
    nx, ny, nz, V = p["nx"], p["ny"], p["nz"], p["num_vars"]
    cv = p["comm_vars"] if 0 < p["comm_vars"] <= V else V
    slots, level, nei_level, nei = r.topology()
    dirs = r.comm_lists()
    used = int(max(slots)) + 1 if len(slots) else 1
    # direction-wide buffer sizes: the last face's offset + its message share
    ssize, rsize = [], []
    for D in dirs:
        s = r_ = 1
        for i in range(len(D["partner"])):
            f0 = D["index"][i]
            s = max(s, int(D["send_off"][f0]) + int(D["send_size"][i]))
            r_ = max(r_, int(D["recv_off"][f0]) + int(D["recv_size"][i]))
        ssize.append(s)
        rsize.append(r_)
    # --code 1|2: the reference additionally writes ghost edges/corners the 7-point
    # stencil never reads; everything the stencil reads must still agree
    mask = np.ones((nx + 2, ny + 2, nz + 2), bool)
    if p["code"] != 0 and p["stencil"] == 7:
        from mputil import defined_mask
        mask = defined_mask(nx, ny, nz, 7)
    cells = 0
    for st in range(stages):
        plan = HaloPlan(nx, ny, nz, V, p["max_blocks"], slots, level, nei_level, nei, dirs=dirs,
                        stencil=p["stencil"], comm_vars=p["comm_vars"], permute=p["permute"], stage=st,
                        rank=p["my_pe"], num_ranks=p["num_pes"])
        for start in range(0, V, cv):
            num = min(cv, V - start)
            data = np.zeros((used, V, nx + 2, ny + 2, nz + 2))
            for s in slots:
                data[s] = r.get_slot(int(s))
            pool = to_pool(data, nx, ny, nz)
            send = [np.zeros(z) for z in ssize]
            recv = [np.zeros(z) for z in rsize]
            for o in range(3):
                d = plan.dirs[o]
                run_pack(plan.pack[o], pool, send, recv, start, num)
                r.exchange_dir(d, send[d], recv[d])          # collective over the partners
            got = from_pool(run_halo(plan, slots, pool, recv, start, num, nx, ny, nz), used, nx, ny, nz)
            r.comm(start, num, st)                            # the reference's own exchange
            for s in slots:
                want = r.get_slot(int(s))
                bad = (got[s, start:start + num].view(np.uint64) != want[start:start + num].view(np.uint64)) & mask[None]
                assert not bad.any(), (f"rank {p['my_pe']} stage {st} vars {start}+{num} slot {s}: "
                                       f"{int(bad.sum())} cells differ, first {np.argwhere(bad)[0]}")
                cells += bad.size
            for v in range(start, start + num):
                r.stencil_driver(v, st)
    nfaces = sum(len(D["block"]) for D in dirs)
    with open(os.path.join(outdir, f"rank{p['my_pe']}.txt"), "w") as f:
        f.write(f"PLAN_OK blocks={len(slots)} offrank_faces={nfaces} cells={cells} "
                f"levels={sorted(set(int(x) for x in level))}\n")


if __name__ == "__main__":
    main()
