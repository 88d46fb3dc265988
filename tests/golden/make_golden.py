"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

    python tests/golden/make_golden.py

Each fixture holds: the reference command line, the mesh topology the reference
built (block.h:36-53 fields, sorted_list order), and — for block data seeded
with numpy RandomState(seed) over every active tile in sorted_list order — the
per-stage check_sum() of every variable and the SHA-256 of all active tiles
after `stages` stages of driver.c:73-89.  /root/reference is only needed here,
never at test time.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refharness import RefMiniAMR  # noqa: E402

SPHERE = "--num_objects 1 --object 2 0 0.3 0.3 0.3 0.01 0.01 0.01 0.25 0.25 0.25 0 0 0"
MOVING = "--num_objects 1 --object 2 0 0.2 0.2 0.2 0.09 0.07 0.05 0.2 0.2 0.2 0 0 0"

CASES = {
    # name: (args, extra move+refine steps, stages, seed)
    "amr7_aniso": (f"--nx 4 --ny 6 --nz 8 --num_vars 3 --num_refine 2 --max_blocks 600 {SPHERE}", 0, 3, 11),
    "amr7_moved_permute": (f"--nx 4 --ny 4 --nz 4 --num_vars 4 --comm_vars 3 --num_refine 3 --block_change 1 "
                           f"--max_blocks 3000 --refine_freq 1 --permute {MOVING}", 5, 7, 12),
    "uni27_aniso": ("--nx 4 --ny 6 --nz 4 --num_vars 3 --stencil 27 --uniform_refine 1 --num_refine 1 "
                    "--init_x 2 --init_y 1 --init_z 2 --max_blocks 100", 0, 3, 13),
    "uni27_permute": ("--nx 4 --ny 4 --nz 6 --num_vars 2 --stencil 27 --uniform_refine 1 --num_refine 2 "
                      "--max_blocks 100 --permute", 0, 7, 14),
    "cfg1_like": (f"--nx 10 --ny 10 --nz 10 --num_vars 2 --stencil 7 --num_refine 4 --max_blocks 4000 {SPHERE}", 0, 2, 15),
    "cfg2_like": ("--nx 16 --ny 16 --nz 16 --num_vars 2 --stencil 27 --uniform_refine 1 --num_refine 1 "
                  "--max_blocks 20", 0, 2, 16),
    "cfg3_like_ring": ("--nx 32 --ny 32 --nz 32 --num_vars 2 --stencil 7 --uniform_refine 1 --num_refine 1 "
                       "--max_blocks 20", 0, 2, 17),
    "ring27": ("--nx 32 --ny 32 --nz 32 --num_vars 1 --stencil 27 --uniform_refine 1 --num_refine 1 "
               "--max_blocks 20", 0, 2, 18),
    # --stencil 0, the variable-work mix (stencil.c:147-983): 13 stages = every update kind twice;
    # the fixture also holds mat, a1, a0[] as init() drew them (init.c:418-423)
    # BASELINE variable counts (the bench's per-CTA variable pipelines run 10-20 tiles deep, cfg5 uses
    # four receive-buffer sets): same shapes as BASELINE.json configs[1], [2], [4], [0]
    "cfg2_v40": ("--nx 16 --ny 16 --nz 16 --num_vars 40 --stencil 27 --uniform_refine 1 --num_refine 2 "
                 "--max_blocks 80", 0, 2, 21),
    "cfg3_v40": ("--nx 32 --ny 32 --nz 32 --num_vars 40 --stencil 7 --uniform_refine 1 --num_refine 1 "
                 "--max_blocks 20", 0, 2, 22),
    "cfg5_v160": ("--nx 10 --ny 10 --nz 10 --num_vars 160 --comm_vars 40 --stencil 27 --uniform_refine 1 "
                  "--num_refine 1 --init_x 2 --init_y 2 --init_z 2 --max_blocks 80", 0, 2, 23),
    "cfg1_v40": (f"--nx 10 --ny 10 --nz 10 --num_vars 40 --stencil 7 --num_refine 4 --max_blocks 4000 {SPHERE}", 0, 2, 24),
    "uni0_variable_work": ("--nx 6 --ny 4 --nz 8 --num_vars 14 --comm_vars 5 --stencil 0 --uniform_refine 1 "
                           "--num_refine 1 --init_x 2 --init_y 1 --init_z 2 --max_blocks 80", 0, 13, 19),
}


def seed_data(ref, seed):
    rs = np.random.RandomState(seed)
    shape = (ref.p["num_vars"],) + ref.tile_shape
    for s in ref.sorted_slots():
        ref.set_slot(int(s), rs.random_sample(shape))


def digest(ref):
    h = hashlib.sha256()
    for s in ref.sorted_slots():
        h.update(np.ascontiguousarray(ref.get_slot(int(s))).tobytes())
    return h.hexdigest()


def main():
    only = set(sys.argv[1:])            # regenerate just these (default: all)
    for name, (args, moves, stages, seed) in CASES.items():
        if only and name not in only:
            continue
        r = RefMiniAMR(args.split())
        r.init()
        r.refine(0)
        for ts in range(1, moves + 1):
            r.move(1.0)
            r.refine(ts)
        slots, lev, nl, ne = r.topology()
        # entries of nei the reference never reads are uninitialised: clear them
        ne = ne.copy()
        for a in range(len(slots)):
            for l in range(6):
                if nl[a, l] == -2:
                    ne[a, l] = 0
                elif nl[a, l] != lev[a] + 1:
                    k = ne[a, l, 0, 0]
                    ne[a, l] = k
        seed_data(r, seed)
        p = r.p
        sums = np.zeros((stages, p["num_vars"]))
        for st in range(stages):
            r.stage(st)
            for v in range(p["num_vars"]):
                sums[st, v] = r.check_sum(v)
        out = os.path.join(HERE, name + ".npz")
        extra = {}
        if p["stencil"] == 0:
            mat, a1, a0 = r.stencil0()
            extra = dict(s0_mat=mat, s0_a1=a1, s0_a0=a0)
        np.savez_compressed(out, args=np.array(args), seed=seed, stages=stages, **extra,
                            params=np.array([p[k] for k in ("nx", "ny", "nz", "num_vars", "comm_vars",
                                                             "max_blocks", "stencil", "permute")], np.int32),
                            slots=slots, level=lev, nei_level=nl, nei=ne,
                            check_sums=sums, sha256=np.array(digest(r)))
        print(f"{name}: {len(slots)} blocks, levels {sorted(set(lev.tolist()))}, "
              f"{os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
