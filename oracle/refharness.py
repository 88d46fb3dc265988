"""TEST INFRASTRUCTURE — ctypes driver for oracle/_ref/libminiamr_{ref,omp}.so.

The library is the UNMODIFIED reference (compiled by oracle/Makefile from
/root/reference where it lies) plus oracle/ref_harness.c.  Every `RefMiniAMR`
instance loads a private copy of the shared object, so several independent
reference instances (each with its own set of miniAMR globals) can live in one
Python process.

Only tests/, bench.py's cpu_baseline / --impl reference leg and
__graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
INT_DIR = os.path.join(os.path.dirname(HERE), "integration", "_bin")

P_NAMES = ["nx", "ny", "nz", "num_vars", "comm_vars", "max_blocks", "stencil",
           "num_refine", "num_active", "max_active_block", "code", "permute",
           "num_pes", "my_pe", "uniform_refine", "stages_per_ts", "num_tsteps",
           "checksum_freq", "refine_freq", "num_parents", "max_active_parent",
           "error_tol"]


def lib_path(variant: str) -> str:
    """"ref" / "omp": the unmodified reference; "int": the same host code linked
    against the CUDA stage path through integration/glue.c.  "ref_mp" / "int_mp":
    the same two linked with the multi-process minimpi back-end (one instance per
    rank, started by minimpi/_bin/minimpirun: tests/mp_worker.py)."""
    if variant in ("int", "int_mp"):
        return os.path.join(INT_DIR, f"libminiamr_{variant}.so")
    return os.path.join(REF_DIR, f"libminiamr_{variant}.so")


def available(variant: str = "ref") -> bool:
    return os.path.exists(lib_path(variant))


class RefMiniAMR:
    """One live instance of the reference program, stopped before driver()."""

    def __init__(self, args, variant: str = "ref", run_driver: bool = False,
                 quiet: bool = True):
        src = lib_path(variant)
        if not os.path.exists(src):
            raise FileNotFoundError(f"{src} missing: run `make -C oracle` / `make -C integration`")
        if variant in ("int", "int_mp"):
            # the harness looks at blocks[].array (sync_host, get_slot): keep the reference's
            # host allocation instead of the drop-in's stub tables (integration/glue.c)
            os.environ["MAMR_LEAN_HOST"] = "0"
            # the private copy below cannot use its $ORIGIN-relative rpath: make the
            # CUDA library resident first, the copy then binds to it by soname
            C.CDLL(os.path.join(os.path.dirname(HERE), "miniamr_b200", "libminiamr_b200.so"),
                   mode=C.RTLD_GLOBAL)
        fd, self._copy = tempfile.mkstemp(suffix=".so", prefix="miniamr_ref_")
        os.close(fd)
        shutil.copyfile(src, self._copy)
        self.lib = C.CDLL(self._copy)
        os.unlink(self._copy)  # mapping stays valid
        L = self.lib
        L.refh_check_sum.restype = C.c_double
        L.refh_check_sum.argtypes = [C.c_int]
        L.refh_calc_time_step.restype = C.c_double
        L.refh_move.argtypes = [C.c_double]
        L.refh_get_block.restype = C.c_longlong
        L.refh_get_parent.restype = C.c_longlong
        L.refh_get_grid_sum.restype = C.c_double
        L.refh_global_active.restype = C.c_longlong
        argv = ["miniAMR.x"] + [str(a) for a in args]
        if quiet and "--report_perf" not in argv:
            argv += ["--report_perf", "0"]
        arr = (C.c_char_p * (len(argv) + 1))()
        for i, a in enumerate(argv):
            arr[i] = a.encode()
        self._argv = arr
        L.refh_reseed()   # rand() state is process-wide; every instance starts at seed 1
        L.refh_start(len(argv), arr, 1 if run_driver else 0)
        self.refresh()

    # ---- parameters / topology -------------------------------------------
    def refresh(self):
        buf = (C.c_int * 64)()
        self.lib.refh_get_params(buf)
        self.p = {k: buf[i] for i, k in enumerate(P_NAMES)}
        p = self.p
        self.tile_shape = (p["nx"] + 2, p["ny"] + 2, p["nz"] + 2)
        self.tile = int(np.prod(self.tile_shape))
        return p

    def init(self):
        self.lib.refh_reseed()
        self.lib.refh_init()
        self.refresh()

    def refine(self, ts: int):
        self.lib.refh_refine(int(ts))
        self.refresh()

    def move(self, delta: float = 1.0):
        self.lib.refh_move(float(delta))

    def sorted_slots(self) -> np.ndarray:
        n = self.lib.refh_get_sorted(None)
        out = np.zeros(max(n, 1), dtype=np.int32)
        self.lib.refh_get_sorted(out.ctypes.data_as(C.POINTER(C.c_int)))
        return out[:n]

    def block(self, slot: int) -> dict:
        buf = (C.c_int * 40)()
        number = self.lib.refh_get_block(int(slot), buf)
        a = np.frombuffer(buf, dtype=np.int32)
        return dict(number=int(number), level=int(a[0]), refine=int(a[1]),
                    nei_level=a[2:8].copy(), nei=a[8:32].reshape(6, 2, 2).copy(),
                    cen=a[32:35].copy())

    def parent(self, p: int) -> dict:
        buf = (C.c_int * 24)()
        number = self.lib.refh_get_parent(int(p), buf)
        a = np.frombuffer(buf, dtype=np.int32)
        return dict(number=int(number), level=int(a[0]), refine=int(a[1]),
                    child=a[2:10].copy(), child_node=a[10:18].copy())

    def topology(self):
        """(slots[num_active], level[num_active], nei_level[num_active,6],
        nei[num_active,6,2,2]) in sorted_list order."""
        slots = self.sorted_slots()
        lev = np.zeros(len(slots), np.int32)
        nl = np.zeros((len(slots), 6), np.int32)
        ne = np.zeros((len(slots), 6, 2, 2), np.int32)
        for a, s in enumerate(slots):
            b = self.block(s)
            lev[a] = b["level"]
            nl[a] = b["nei_level"]
            ne[a] = b["nei"]
        return slots, lev, nl, ne

    # ---- data -------------------------------------------------------------
    def get_tile(self, slot: int, var: int) -> np.ndarray:
        out = np.empty(self.tile, dtype=np.float64)
        self.lib.refh_get_tile(int(slot), int(var), out.ctypes.data_as(C.c_void_p))
        return out.reshape(self.tile_shape)

    def set_tile(self, slot: int, var: int, data: np.ndarray):
        d = np.ascontiguousarray(data, dtype=np.float64).reshape(-1)
        assert d.size == self.tile
        self.lib.refh_set_tile(int(slot), int(var), d.ctypes.data_as(C.c_void_p))

    def get_slot(self, slot: int) -> np.ndarray:
        out = np.empty((self.p["num_vars"],) + self.tile_shape, dtype=np.float64)
        self.lib.refh_get_slot(int(slot), out.ctypes.data_as(C.c_void_p))
        return out

    def set_slot(self, slot: int, data: np.ndarray):
        d = np.ascontiguousarray(data, dtype=np.float64).reshape(-1)
        assert d.size == self.tile * self.p["num_vars"]
        self.lib.refh_set_slot(int(slot), d.ctypes.data_as(C.c_void_p))

    def sync_host(self):
        """integration build: copy the device-resident block data back to blocks[].array"""
        self.lib.refh_sync_host()

    def get_active(self) -> dict:
        """{slot: array[num_vars, nx+2, ny+2, nz+2]} for every active block."""
        return {int(s): self.get_slot(int(s)) for s in self.sorted_slots()}

    # ---- the hot path, one call at a time ----------------------------------
    def comm(self, start: int, num_comm: int, stage: int):
        self.lib.refh_comm(int(start), int(num_comm), int(stage))

    def stencil_driver(self, var: int, stage: int = 0):
        self.lib.refh_stencil_driver(int(var), int(stage))

    def check_sum(self, var: int) -> float:
        return float(self.lib.refh_check_sum(int(var)))

    def stage(self, stage: int):
        self.lib.refh_stage(int(stage))

    def pack_block(self, slot: int) -> np.ndarray:
        n = 50 + self.p["num_vars"] * self.p["nx"] * self.p["ny"] * self.p["nz"]
        out = np.empty(n, dtype=np.float64)
        self.lib.refh_pack_block(int(slot), out.ctypes.data_as(C.c_void_p), n)
        return out

    def unpack_block(self, slot: int, msg: np.ndarray):
        d = np.ascontiguousarray(msg, dtype=np.float64)
        self.lib.refh_unpack_block(int(slot), d.ctypes.data_as(C.c_void_p))

    def pack_face(self, slot, face_case, d, start, num_comm) -> np.ndarray:
        p = self.p
        buf = np.full(num_comm * (max(p["nx"], p["ny"], p["nz"]) + 2) ** 2, np.nan)
        self.lib.refh_pack_face(buf.ctypes.data_as(C.c_void_p), int(slot),
                                int(face_case), int(d), int(start), int(num_comm))
        return buf[~np.isnan(buf)]

    def unpack_face(self, buf, slot, face_case, d, start, num_comm):
        b = np.ascontiguousarray(buf, np.float64)
        self.lib.refh_unpack_face(b.ctypes.data_as(C.c_void_p), int(slot),
                                  int(face_case), int(d), int(start), int(num_comm))

    def timers(self) -> dict:
        buf = (C.c_double * 16)()
        self.lib.refh_get_timers(buf)
        names = ["calc", "comm", "checksum", "refine", "all", "total_blocks",
                 "num_tsteps", "fp_adds", "fp_divs"]
        return {k: buf[i] for i, k in enumerate(names)}

    def counters(self) -> dict:
        buf = (C.c_int * 9)()
        self.lib.refh_get_counters(buf)
        return dict(same=list(buf[0:3]), diff=list(buf[3:6]), bc=list(buf[6:9]))

    def comm_lists(self):
        """the off-rank comm lists of comm.h:38-55 as the `dirs` argument of
        miniamr_b200.capi (three dicts of int32 arrays)"""
        names = ["partner", "index", "num", "send_size", "recv_size", "block", "face_case",
                 "send_off", "recv_off"]
        dirs = []
        for d in range(3):
            D = {}
            for w, k in enumerate(names):
                n = self.lib.refh_get_comm_list(d, w, None)
                a = np.zeros(max(n, 1), np.int32)
                self.lib.refh_get_comm_list(d, w, a.ctypes.data_as(C.POINTER(C.c_int)))
                D[k] = a[:n].copy()
            dirs.append(D)
        return dirs

    def exchange_dir(self, d: int, send: np.ndarray, recv: np.ndarray):
        """one message per partner of direction d over the host MPI (comm.c:71-84,120-151)"""
        assert send.dtype == np.float64 and recv.dtype == np.float64
        self.lib.refh_exchange_dir(int(d), send.ctypes.data_as(C.c_void_p),
                                   recv.ctypes.data_as(C.c_void_p))

    def stencil0(self):
        """(mat, a1, a0[mat]) of a --stencil 0 run (init.c:418-423), mat = 0 otherwise"""
        a1 = C.c_double()
        a0 = np.zeros(max(1, self.p["num_vars"]), np.float64)
        mat = self.lib.refh_get_stencil0(C.byref(a1), a0.ctypes.data_as(C.c_void_p))
        return int(mat), float(a1.value), a0[:mat].copy()

    def flops(self):
        buf = (C.c_double * 3)()
        self.lib.refh_get_flops(buf)
        return dict(adds=buf[0], muls=buf[1], divs=buf[2])

    def global_active(self) -> int:
        return int(self.lib.refh_global_active())
