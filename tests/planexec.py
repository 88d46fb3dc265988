"""TEST INFRASTRUCTURE: a numpy executor of the halo plan (miniamr_b200.capi.HaloPlan),
i.e. of exactly the descriptors the fused CUDA kernel and the pack kernel consume.
It lets the CPU suite check the host-side resolution logic (plan.cu) against the
oracle, including the multi-rank message layout over gloo, without a GPU."""
import numpy as np

from miniamr_b200.capi import PLAN_FIELDS

F = {k: i for i, k in enumerate(PLAN_FIELDS)}


def tile_stride(nx, ny, nz):
    t = (nx + 2)*(ny + 2)*(nz + 2)
    return (t + 15)//16*16


def to_pool(data, nx, ny, nz):
    """oracle layout data[slot, var, i, j, k] -> pool[var, slot*tile_stride + cell]"""
    nslot, V = data.shape[:2]
    ts = tile_stride(nx, ny, nz)
    pool = np.zeros((V, nslot*ts))
    flat = data.reshape(nslot, V, -1)
    for s in range(nslot):
        pool[:, s*ts:s*ts + flat.shape[2]] = flat[s]
    return pool


def from_pool(pool, nslot, nx, ny, nz):
    V = pool.shape[0]
    ts = tile_stride(nx, ny, nz)
    t = (nx + 2)*(ny + 2)*(nz + 2)
    out = np.zeros((nslot, V, nx + 2, ny + 2, nz + 2))
    for s in range(nslot):
        out[s] = pool[:, s*ts:s*ts + t].reshape(V, nx + 2, ny + 2, nz + 2)
    return out


def _values(op, src):
    a, b, c = np.meshgrid(np.arange(op[F["e0"]]), np.arange(op[F["e1"]]), np.arange(op[F["e2"]]),
                          indexing="ij")
    s0, s1, s2 = op[F["ss0"]], op[F["ss1"]], op[F["ss2"]]
    mode = op[F["mode"]]
    base = op[F["src_base"]]
    if mode in (0, 1):
        v = src[base + a*s0 + b*s1 + c*s2]
        return v/4.0 if mode == 1 else v
    if mode in (2, 3):
        v = src[base + (a >> 1)*s0 + (b >> 1)*s1 + (c >> 1)*s2]
        return v/4.0 if mode == 2 else v
    p = base + 2*a*s0 + 2*b*s1 + 2*c*s2
    S, Fs = op[F["S"]], op[F["F"]]
    return ((src[p] + src[p + Fs]) + src[p + S]) + src[p + S + Fs]


def _dst_index(op):
    a, b, c = np.meshgrid(np.arange(op[F["e0"]]), np.arange(op[F["e1"]]), np.arange(op[F["e2"]]),
                          indexing="ij")
    return op[F["dst_base"]] + a*op[F["ds0"]] + b*op[F["ds1"]] + c*op[F["ds2"]]


def run_pack(ops, pool, send, recv, start, num):
    """execute the pack ops of one phase: fill send[d] for variables start..start+num-1"""
    for op in ops:
        for v in range(start, start + num):
            src = pool[v] if op[F["src_mem"]] == 0 else \
                recv[op[F["src_mem"]] - 1][(v - start)*op[F["src_vs"]]:]
            dst = send[op[F["dst_mem"]] - 1]
            dst[(v - start)*op[F["dst_vs"]] + _dst_index(op)] = _values(op, src)


def run_halo(plan, slots, pool, recv, start, num, nx, ny, nz):
    """what the fused kernel does before the stencil: every ghost cell of every
    active tile from its resolved origin; returns the new pool (interiors kept)"""
    ts = tile_stride(nx, ny, nz)
    out = pool.copy()
    for a, slot in enumerate(slots):
        for op in plan.halo[plan.begin[a]:plan.begin[a + 1]]:
            for v in range(start, start + num):
                src = pool[v] if op[F["src_mem"]] == 0 else \
                    recv[op[F["src_mem"]] - 1][(v - start)*op[F["src_vs"]]:]
                out[v][slot*ts + _dst_index(op)] = _values(op, src)
    return out
