"""Host-side synthetic meshes for bench.py and the tests: a uniform (single
level) mesh of blocks on the unit cube with reflective domain boundaries,
optionally sharded over an npx x npy x npz rank grid exactly as the reference
shards blocks over MPI ranks (init.c:156-190, 453-468): rank r owns one
contiguous sub-cube of the block grid.  Produces what the device path consumes:
the per-block topology (block.h:36-53 fields) and, for N > 1, the per-direction
comm lists (comm.h:38-55) with the reference's message layout (one message per
direction and partner, faces at a fixed stride of comm_vars*len, comm_util.c).

This is integer bookkeeping for synthetic inputs; refinement, load balancing and
the real comm-list maintenance stay in the reference's host code
(integration/glue.c feeds the device path from its globals).
"""
from __future__ import annotations

import numpy as np


def face_len(d, nx, ny, nz, stencil):
    """msg_len[d][case 0/1] (init.c:81-90): whole-face message length per var."""
    wide = stencil != 7
    if d == 0:
        return ny * nz
    if d == 1:
        return (nx + 2 if wide else nx) * nz
    return (nx + 2 if wide else nx) * (ny + 2 if wide else ny)


def rank_coords(rank, npx, npy, npz):
    return rank % npx, (rank // npx) % npy, rank // (npx * npy)


def uniform_mesh(bx, by, bz, npx=1, npy=1, npz=1, rank=0, nx=0, ny=0, nz=0,
                 comm_vars=1, stencil=7):
    """Topology (+ comm lists) of rank `rank`'s bx x by x bz sub-cube.

    Returns dict(slots, level, nei_level, nei, dirs) with arrays in active order
    (slot = lx + bx*(ly + by*lz)); `dirs` is the 3-element comm-list structure
    DeviceMesh.set_comm_lists() takes (empty lists when there is one rank)."""
    nb = bx * by * bz
    rx, ry, rz = rank_coords(rank, npx, npy, npz)
    rc = (rx, ry, rz)
    npd = (npx, npy, npz)
    bd = (bx, by, bz)
    slots = np.arange(nb, dtype=np.int32)
    level = np.zeros(nb, np.int32)
    nei_level = np.zeros((nb, 6), np.int32)
    nei = np.zeros((nb, 6, 2, 2), np.int32)
    lx, ly, lz = np.meshgrid(np.arange(bx), np.arange(by), np.arange(bz), indexing="ij")
    lc = [lx.transpose(2, 1, 0).reshape(-1), ly.transpose(2, 1, 0).reshape(-1),
          lz.transpose(2, 1, 0).reshape(-1)]          # slot-ordered local coords
    stride = (1, bx, bx * by)
    faces = [[] for _ in range(3)]                     # per dir: (partner, pos, sign, slot)
    for d in range(3):
        for side in (0, 1):
            l = 2 * d + side
            step = -1 if side == 0 else 1
            c = lc[d] + step
            inside = (c >= 0) & (c < bd[d])
            nbr_rank_c = rc[d] + step
            has_rank = 0 <= nbr_rank_c < npd[d]
            for s in range(nb):
                if inside[s]:
                    nei_level[s, l] = 0
                    nei[s, l, 0, 0] = s + step * stride[d]
                elif has_rank:
                    prc = list(rc)
                    prc[d] = nbr_rank_c
                    partner = prc[0] + npx * (prc[1] + npy * prc[2])
                    nei_level[s, l] = 0
                    nei[s, l, :, :] = -1 - partner
                    sa, fa = ((1, 2), (0, 2), (0, 1))[d]
                    pos = lc[sa][s] * bd[fa] + lc[fa][s]   # in-face position, slow then fast
                    faces[d].append((partner, pos, side, s))
                else:
                    nei_level[s, l] = -2
    dirs = []
    f = 0 if stencil == 7 else 1
    for d in range(3):
        ln = face_len(d, nx, ny, nz, stencil) if nx else 0
        fl = sorted(faces[d])
        D = dict(partner=[], index=[], num=[], send_size=[], recv_size=[], block=[],
                 face_case=[], send_off=[], recv_off=[])
        off = 0
        for partner, pos, side, s in fl:
            if not D["partner"] or D["partner"][-1] != partner:
                D["partner"].append(partner)
                D["index"].append(len(D["block"]))
                D["num"].append(0)
                D["send_size"].append(0)
                D["recv_size"].append(0)
            D["num"][-1] += 1
            D["send_size"][-1] += comm_vars * ln
            D["recv_size"][-1] += comm_vars * ln
            D["block"].append(s)
            D["face_case"].append(f + 10 * side)
            D["send_off"].append(off)
            D["recv_off"].append(off)
            off += comm_vars * ln
        dirs.append({k: np.asarray(v, np.int32) for k, v in D.items()})
    return dict(slots=slots, level=level, nei_level=nei_level, nei=nei, dirs=dirs)
