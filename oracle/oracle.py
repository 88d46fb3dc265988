"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/liboracle.so (oracle.c).

`OracleMesh` holds one rank's block pool as a numpy array
data[slot, var, i, j, k] plus the slot-indexed topology arrays and offers the
reference's call surface (comm / stencil_driver / check_sum / split /
consolidate / pack_block) on it.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", HERE, path])
        L = C.CDLL(path)
        L.orc_check_sum.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleMesh:
    def __init__(self, nx, ny, nz, num_vars, max_blocks, stencil=7, comm_vars=0,
                 permute=0):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.num_vars = num_vars
        self.comm_vars = comm_vars if 0 < comm_vars <= num_vars else num_vars
        self.max_blocks = max_blocks
        self.stencil = stencil
        self.permute = permute
        self.data = np.zeros((max_blocks, num_vars, nx + 2, ny + 2, nz + 2))
        self.level = np.zeros(max_blocks, np.int32)
        self.nei_level = np.zeros((max_blocks, 6), np.int32)
        self.nei = np.zeros((max_blocks, 6, 2, 2), np.int32)
        self.slots = np.zeros(0, np.int32)
        self.L = lib()

    @property
    def dims(self):
        return (self.nx, self.ny, self.nz, self.num_vars)

    def set_topology(self, slots, level, nei_level, nei):
        """arrays in sorted_list order (as RefMiniAMR.topology() returns)."""
        self.slots = np.ascontiguousarray(slots, np.int32)
        for a, s in enumerate(self.slots):
            self.level[s] = level[a]
            self.nei_level[s] = nei_level[a]
            self.nei[s] = nei[a]

    def comm_dir_local(self, d, start, num_comm):
        rc = self.L.orc_comm_dir_local(_p(self.data), *self.dims, self.stencil,
                                       len(self.slots), _p(self.slots), _p(self.level),
                                       _p(self.nei_level), _p(self.nei), int(d),
                                       int(start), int(num_comm), None)
        if rc:
            raise RuntimeError("ERROR: misconnected block")

    def phase_order(self, stage):
        o = (C.c_int * 3)()
        self.L.orc_phase_order(int(self.permute), int(stage), o)
        return list(o)

    def comm(self, start, num_comm, stage=0):
        counters = np.zeros(9, np.int32)
        rc = self.L.orc_comm_local(_p(self.data), *self.dims, self.stencil,
                                   len(self.slots), _p(self.slots), _p(self.level),
                                   _p(self.nei_level), _p(self.nei), int(start),
                                   int(num_comm), int(stage), int(self.permute),
                                   _p(counters))
        if rc:
            raise RuntimeError("ERROR: misconnected block")
        return counters.reshape(3, 3)  # [dir][same, diff, bc]

    def set_stencil0(self, mat, a1, a0):
        """--stencil 0 coefficients (init.c:418-423)"""
        self.s0 = (int(mat), float(a1), np.ascontiguousarray(a0, np.float64))
        self.flops = np.zeros(3)          # adds, muls, divs as stencil.c books them

    def stencil_driver(self, var, stage=0):
        if self.stencil == 0:
            mat, a1, a0 = self.s0
            self.L.orc_stencil0_driver.argtypes = [C.c_void_p] + [C.c_int]*5 + [C.c_void_p] + [C.c_int]*3 + \
                [C.c_double, C.c_void_p, C.c_void_p]
            self.L.orc_stencil0_driver(self.data.ctypes.data, *self.dims, len(self.slots),
                                       self.slots.ctypes.data, int(var), int(stage), mat, a1,
                                       a0.ctypes.data, self.flops.ctypes.data)
            return
        self.L.orc_stencil_calc(_p(self.data), *self.dims, len(self.slots),
                                _p(self.slots), int(var), self.stencil)

    def check_sum(self, var):
        return float(self.L.orc_check_sum(_p(self.data), *self.dims, len(self.slots),
                                          _p(self.slots), int(var)))

    def stage(self, stage=0):
        if self.stencil == 0:             # driver.c:75-89 with the stage-dependent update
            for start in range(0, self.num_vars, self.comm_vars):
                num = min(self.comm_vars, self.num_vars - start)
                self.comm(start, num, stage)
                for v in range(start, start + num):
                    self.stencil_driver(v, stage)
            return
        rc = self.L.orc_stage_local(_p(self.data), *self.dims, self.comm_vars,
                                    self.stencil, len(self.slots), _p(self.slots),
                                    _p(self.level), _p(self.nei_level), _p(self.nei),
                                    int(stage), int(self.permute))
        if rc:
            raise RuntimeError("ERROR: misconnected block")

    def pack_face(self, slot, face_case, d, start, num_comm):
        buf = np.zeros(num_comm * (max(self.nx, self.ny, self.nz) + 2) ** 2)
        n = self.L.orc_pack_face(_p(self.data), *self.dims, self.stencil, _p(buf),
                                 int(slot), int(face_case), int(d), int(start),
                                 int(num_comm))
        return buf[:n].copy()

    def unpack_face(self, buf, slot, face_case, d, start, num_comm):
        b = np.ascontiguousarray(buf, np.float64)
        return self.L.orc_unpack_face(_p(self.data), *self.dims, self.stencil, _p(b),
                                      int(slot), int(face_case), int(d), int(start),
                                      int(num_comm))

    def split_block(self, parent_slot, child_slots):
        c = np.ascontiguousarray(child_slots, np.int32)
        self.L.orc_split_block(_p(self.data), *self.dims, int(parent_slot), _p(c))

    def consolidate_block(self, child_slots, parent_slot):
        c = np.ascontiguousarray(child_slots, np.int32)
        self.L.orc_consolidate_block(_p(self.data), *self.dims, _p(c), int(parent_slot))

    def pack_block(self, slot):
        out = np.zeros(self.num_vars * self.nx * self.ny * self.nz)
        self.L.orc_pack_block(_p(self.data), *self.dims, int(slot), _p(out))
        return out

    def unpack_block(self, slot, payload):
        b = np.ascontiguousarray(payload, np.float64)
        self.L.orc_unpack_block(_p(self.data), *self.dims, int(slot), _p(b))
