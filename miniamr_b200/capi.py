"""ctypes binding of include/miniamr_b200.h and a thin host-side mirror of the
reference's call surface for the stage hot path (same names, argument meaning
and error behaviour as ref/proto.h: comm, stencil_driver, check_sum,
pack_block, unpack_block, split/consolidate data movement).

The CUDA library is the product; this module never computes anything itself
and raises if the library (or a CUDA device) is missing — there is no CPU
fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# (MAMR_LIB_PATH: an experimental build of the same library, for A/B measurements)
LIB_PATH = os.environ.get("MAMR_LIB_PATH") or os.path.join(HERE, "libminiamr_b200.so")


class MamrError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("nx", "ny", "nz", "num_vars", "comm_vars", "max_blocks", "stencil",
                 "code", "permute", "device", "rank", "num_ranks")]


class Block(C.Structure):
    _fields_ = [("slot", C.c_int), ("level", C.c_int), ("nei_level", C.c_int * 6),
                ("nei", C.c_int * 24)]


class CommDir(C.Structure):
    _fields_ = [("num_partners", C.c_int),
                ("partner", C.POINTER(C.c_int)), ("index", C.POINTER(C.c_int)),
                ("num", C.POINTER(C.c_int)), ("send_size", C.POINTER(C.c_int)),
                ("recv_size", C.POINTER(C.c_int)),
                ("num_cases", C.c_int),
                ("block", C.POINTER(C.c_int)), ("face_case", C.POINTER(C.c_int)),
                ("send_off", C.POINTER(C.c_int)), ("recv_off", C.POINTER(C.c_int))]


class Counters(C.Structure):
    _fields_ = [("counter_same", C.c_longlong * 3), ("counter_diff", C.c_longlong * 3),
                ("counter_bc", C.c_longlong * 3),
                ("counter_halo_send", C.c_longlong * 3), ("counter_halo_recv", C.c_longlong * 3),
                ("counter_face_send", C.c_longlong * 3), ("counter_face_recv", C.c_longlong * 3),
                ("size_mesg_send", C.c_double * 3), ("size_mesg_recv", C.c_double * 3),
                ("total_fp_adds", C.c_double), ("total_fp_divs", C.c_double),
                ("total_red", C.c_longlong), ("kernel_launches", C.c_longlong),
                ("migrate_bytes", C.c_double), ("ghost_regens", C.c_longlong),
                ("total_fp_muls", C.c_double)]


EXPORTS = [
    "mamr_abi_version", "mamr_create", "mamr_destroy", "mamr_last_error", "mamr_sync",
    "mamr_get_counters", "mamr_reset_counters", "mamr_tile_doubles", "mamr_pool_bytes",
    "mamr_upload_block", "mamr_download_block", "mamr_upload_tile",
    "mamr_download_tile", "mamr_zero_block", "mamr_upload_vars", "mamr_download_vars", "mamr_upload_interiors", "mamr_set_topology", "mamr_set_comm_lists",
    "mamr_comm", "mamr_stencil_driver", "mamr_stencil_calc", "mamr_stencil_vars", "mamr_set_stencil0",
    "mamr_check_sum", "mamr_check_sum_vars", "mamr_stage", "mamr_split_block",
    "mamr_consolidate_block", "mamr_pack_block", "mamr_unpack_block", "mamr_send_block",
    "mamr_recv_block", "mamr_stage_send_block", "mamr_stage_recv_block", "mamr_flush_block_moves",
    "mamr_pending_block_moves", "mamr_device_count", "mamr_set_message_mode",
    "mamr_nccl_get_unique_id", "mamr_nccl_init", "mamr_p2p_get_handle", "mamr_p2p_connect",
    "mamr_timer_begin",
    "mamr_timer_end", "mamr_kernel_timing", "mamr_kernel_time_ms", "mamr_get_device_times",
    "mamr_plan_create", "mamr_plan_phase_dir", "mamr_plan_num_ops", "mamr_plan_get_ops",
    "mamr_plan_block_begin", "mamr_plan_destroy",
]

_LIB = None


def load_library(path: str = LIB_PATH):
    """dlopen the CUDA library; fail loudly if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(path):
        raise MamrError(f"{path} not built: run `python -m miniamr_b200.build` "
                        "(there is no CPU fallback)")
    L = C.CDLL(path)
    L.mamr_last_error.restype = C.c_char_p
    L.mamr_tile_doubles.restype = C.c_longlong
    L.mamr_pool_bytes.restype = C.c_longlong
    L.mamr_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("mamr_create", "mamr_abi_version", "mamr_last_error",
                        "mamr_nccl_get_unique_id"):
            fn.argtypes = None
    if L.mamr_abi_version() != 1:
        raise MamrError("libminiamr_b200.so ABI version mismatch")
    _LIB = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _block_array(slots, level, nei_level, nei):
    n = len(slots)
    arr = (Block * max(n, 1))()
    packed = np.zeros((max(n, 1), 32), np.int32)
    if n:
        packed[:n, 0] = slots
        packed[:n, 1] = level
        packed[:n, 2:8] = np.asarray(nei_level, np.int32).reshape(n, 6)
        packed[:n, 8:32] = np.asarray(nei, np.int32).reshape(n, 24)
    C.memmove(arr, packed.ctypes.data, packed.nbytes)
    return arr


def _comm_dirs(dirs):
    arr = (CommDir * 3)()
    keep = []
    for d in range(3):
        D = dirs[d] if dirs else {}
        a = {k: np.ascontiguousarray(D.get(k, []), np.int32) for k in
             ("partner", "index", "num", "send_size", "recv_size", "block", "face_case",
              "send_off", "recv_off")}
        keep.append(a)
        arr[d].num_partners = len(a["partner"])
        arr[d].num_cases = len(a["block"])
        for k in a:
            setattr(arr[d], k, _ip(a[k]))
    return arr, keep


PLAN_FIELDS = ("dst_base src_base dst_vs src_vs e0 e1 e2 ds0 ds1 ds2 ss0 ss1 ss2 S F first mode "
               "dst_mem src_mem").split()


class HaloPlan:
    """Host-only view of what one comm() call resolves to (no device needed):
    .halo  int64[n_ops, 19] + .begin int32[num_active+1]; .pack[o] per phase;
    .dirs[o] the direction of phase o.  Field names in PLAN_FIELDS."""

    def __init__(self, nx, ny, nz, num_vars, max_blocks, slots, level, nei_level, nei, dirs=None,
                 stencil=7, comm_vars=0, permute=0, stage=0, rank=0, num_ranks=1):
        L = load_library()
        prm = Params(nx, ny, nz, num_vars, comm_vars, max_blocks, stencil, 0, permute, -1, rank,
                     num_ranks)
        blocks = _block_array(slots, level, nei_level, nei)
        darr, keep = _comm_dirs(dirs)
        h = C.c_void_p()
        if L.mamr_plan_create(C.byref(prm), len(slots), blocks, darr, int(stage), C.byref(h)):
            raise MamrError(L.mamr_last_error().decode())
        try:
            def ops(which):
                n = L.mamr_plan_num_ops(h, which)
                out = np.zeros((max(n, 0), len(PLAN_FIELDS)), np.int64)
                if n > 0 and L.mamr_plan_get_ops(h, which, _dp(out)):
                    raise MamrError(L.mamr_last_error().decode())
                return out
            self.halo = ops(0)
            self.pack = [ops(1), ops(2), ops(3)]
            self.begin = np.zeros(len(slots) + 1, np.int32)
            L.mamr_plan_block_begin(h, _ip(self.begin))
            self.dirs = [L.mamr_plan_phase_dir(h, o) for o in range(3)]
        finally:
            L.mamr_plan_destroy(h)


class DeviceMesh:
    """One rank's device-resident block pool + the hot-path calls on it."""

    def __init__(self, nx, ny, nz, num_vars, max_blocks, stencil=7, comm_vars=0,
                 permute=0, code=0, device=-1, rank=0, num_ranks=1):
        self.L = load_library()
        self.params = Params(nx, ny, nz, num_vars, comm_vars, max_blocks, stencil, code,
                             permute, device, rank, num_ranks)
        self.nx, self.ny, self.nz, self.num_vars = nx, ny, nz, num_vars
        self.comm_vars = comm_vars if 0 < comm_vars <= num_vars else num_vars
        self.max_blocks = max_blocks
        self.stencil = stencil
        self.tile_shape = (nx + 2, ny + 2, nz + 2)
        h = C.c_void_p()
        rc = self.L.mamr_create(C.byref(self.params), C.byref(h))
        if rc:
            raise MamrError(self._err())
        self.h = h
        self.num_active = 0
        self._keep = None

    def _err(self):
        return self.L.mamr_last_error().decode()

    def _ck(self, rc):
        if rc:
            raise MamrError(self._err())

    def close(self):
        if getattr(self, "h", None):
            self.L.mamr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data ----------------------------------------------------------------
    def upload_block(self, slot, tiles):
        t = np.ascontiguousarray(tiles, np.float64)
        assert t.shape == (self.num_vars,) + self.tile_shape, t.shape
        self._ck(self.L.mamr_upload_block(self.h, int(slot), _dp(t)))

    def download_block(self, slot):
        out = np.empty((self.num_vars,) + self.tile_shape, np.float64)
        self._ck(self.L.mamr_download_block(self.h, int(slot), _dp(out)))
        return out

    def upload_tile(self, slot, var, tile):
        t = np.ascontiguousarray(tile, np.float64)
        assert t.shape == self.tile_shape
        self._ck(self.L.mamr_upload_tile(self.h, int(slot), int(var), _dp(t)))

    def download_tile(self, slot, var):
        out = np.empty(self.tile_shape, np.float64)
        self._ck(self.L.mamr_download_tile(self.h, int(slot), int(var), _dp(out)))
        return out

    def zero_block(self, slot):
        self._ck(self.L.mamr_zero_block(self.h, int(slot)))

    def upload_vars(self, var_start, num, num_slots, host_ptr):
        """host_ptr: address of [num][num_slots][tile] doubles (pinned => asynchronous)."""
        self._ck(self.L.mamr_upload_vars(self.h, int(var_start), int(num), int(num_slots),
                                         C.c_void_p(int(host_ptr))))

    def upload_interiors(self, var_start, num, num_slots, host_ptr):
        """host_ptr: address of [num_slots][num][nx][ny][nz] doubles (interiors only, ghost layer
        becomes zero; pinned => asynchronous)."""
        self._ck(self.L.mamr_upload_interiors(self.h, int(var_start), int(num), int(num_slots),
                                              C.c_void_p(int(host_ptr))))

    def download_vars(self, var_start, num, num_slots, host_ptr):
        self._ck(self.L.mamr_download_vars(self.h, int(var_start), int(num), int(num_slots),
                                           C.c_void_p(int(host_ptr))))

    def pool_bytes(self):
        return int(self.L.mamr_pool_bytes(self.h))

    # ---- topology ------------------------------------------------------------
    def set_topology(self, slots, level, nei_level, nei):
        """Arrays in sorted_list order: slots[n], level[n], nei_level[n,6],
        nei[n,6,2,2] (block.h:36-53)."""
        n = len(slots)
        arr = (Block * max(n, 1))()
        packed = np.zeros((max(n, 1), 32), np.int32)
        if n:
            packed[:n, 0] = slots
            packed[:n, 1] = level
            packed[:n, 2:8] = np.asarray(nei_level, np.int32).reshape(n, 6)
            packed[:n, 8:32] = np.asarray(nei, np.int32).reshape(n, 24)
        C.memmove(arr, packed.ctypes.data, packed.nbytes)
        self._ck(self.L.mamr_set_topology(self.h, n, arr))
        self.num_active = n

    def set_comm_lists(self, dirs):
        """dirs: 3 dicts with int arrays partner, index, num, send_size, recv_size,
        block, face_case, send_off, recv_off (comm.h:38-55)."""
        arr = (CommDir * 3)()
        keep = []
        for d in range(3):
            D = dirs[d]
            a = {k: np.ascontiguousarray(D.get(k, []), np.int32) for k in
                 ("partner", "index", "num", "send_size", "recv_size", "block", "face_case",
                  "send_off", "recv_off")}
            keep.append(a)
            arr[d].num_partners = len(a["partner"])
            arr[d].num_cases = len(a["block"])
            for k in a:
                setattr(arr[d], k, _ip(a[k]))
        self._keep = keep
        self._ck(self.L.mamr_set_comm_lists(self.h, arr))

    # ---- the reference's call surface ------------------------------------------
    def comm(self, start, num_comm, stage=0):
        self._ck(self.L.mamr_comm(self.h, int(start), int(num_comm), int(stage)))

    def stencil_driver(self, var, calc_stage=0):
        self._ck(self.L.mamr_stencil_driver(self.h, int(var), int(calc_stage)))

    def set_stencil0(self, mat, a1, a0):
        a = np.ascontiguousarray(a0, np.float64)
        self._keep_a0 = a
        self.L.mamr_set_stencil0.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        self._ck(self.L.mamr_set_stencil0(self.h, int(mat), float(a1), a.ctypes.data))

    def stencil_calc(self, var):
        self._ck(self.L.mamr_stencil_calc(self.h, int(var)))

    def stencil_vars(self, start, num):
        self._ck(self.L.mamr_stencil_vars(self.h, int(start), int(num)))

    def check_sum(self, var):
        s = C.c_double()
        self._ck(self.L.mamr_check_sum(self.h, int(var), C.byref(s)))
        return s.value

    def check_sum_vars(self, start, num):
        out = np.zeros(num, np.float64)
        self._ck(self.L.mamr_check_sum_vars(self.h, int(start), int(num), _dp(out)))
        return out

    def stage(self, stage=0):
        self._ck(self.L.mamr_stage(self.h, int(stage)))

    def split_block(self, parent_slot, child_slots):
        c = np.ascontiguousarray(child_slots, np.int32)
        assert c.size == 8
        self._ck(self.L.mamr_split_block(self.h, int(parent_slot), _ip(c)))

    def consolidate_block(self, child_slots, parent_slot):
        c = np.ascontiguousarray(child_slots, np.int32)
        assert c.size == 8
        self._ck(self.L.mamr_consolidate_block(self.h, _ip(c), int(parent_slot)))

    def pack_block(self, slot):
        out = np.empty(self.num_vars * self.nx * self.ny * self.nz, np.float64)
        self._ck(self.L.mamr_pack_block(self.h, int(slot), _dp(out)))
        return out

    def unpack_block(self, slot, payload):
        p = np.ascontiguousarray(payload, np.float64)
        assert p.size == self.num_vars * self.nx * self.ny * self.nz
        self._ck(self.L.mamr_unpack_block(self.h, int(slot), _dp(p)))

    def send_block(self, slot, dest):
        self._ck(self.L.mamr_send_block(self.h, int(slot), int(dest)))

    def recv_block(self, slot, src):
        self._ck(self.L.mamr_recv_block(self.h, int(slot), int(src)))

    # staged form (what integration/glue.c uses under exchange(), rcb.c:207-337)
    def stage_send_block(self, slot, dest):
        self._ck(self.L.mamr_stage_send_block(self.h, int(slot), int(dest)))

    def stage_recv_block(self, slot, src):
        self._ck(self.L.mamr_stage_recv_block(self.h, int(slot), int(src)))

    def flush_block_moves(self):
        self._ck(self.L.mamr_flush_block_moves(self.h))

    def pending_block_moves(self):
        return int(self.L.mamr_pending_block_moves(self.h))

    def sync(self):
        self._ck(self.L.mamr_sync(self.h))

    # ---- multi-GPU -------------------------------------------------------------
    @staticmethod
    def nccl_unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(128)
        if L.mamr_nccl_get_unique_id(buf):
            raise MamrError(L.mamr_last_error().decode())
        return buf.raw

    def nccl_init(self, uid: bytes):
        assert len(uid) == 128
        self._ck(self.L.mamr_nccl_init(self.h, C.create_string_buffer(uid, 128)))

    # peer-memory transport: handle out, all ranks' handles (rank order) in
    def p2p_handle(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._ck(self.L.mamr_p2p_get_handle(self.h, buf))
        return buf.raw

    def p2p_connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == 128*self.params.num_ranks
        self._ck(self.L.mamr_p2p_connect(self.h, C.create_string_buffer(blob, len(blob))))

    # ---- measurement -----------------------------------------------------------
    def counters(self):
        c = Counters()
        self._ck(self.L.mamr_get_counters(self.h, C.byref(c)))
        out = {}
        for name, _ in Counters._fields_:
            v = getattr(c, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out

    def reset_counters(self):
        self._ck(self.L.mamr_reset_counters(self.h))

    def timer_begin(self):
        self._ck(self.L.mamr_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_float()
        self._ck(self.L.mamr_timer_end(self.h, C.byref(ms)))
        return ms.value

    def kernel_timing(self, enable=True):
        self._ck(self.L.mamr_kernel_timing(self.h, 1 if enable else 0))

    def device_times(self, wait=True):
        class T(C.Structure):
            _fields_ = [(n, C.c_double) for n in
                        ("fused_ms", "stencil_ms", "split_ghost_ms", "pack_ms", "exchange_ms", "unpack_ms",
                         "regen_ms", "checksum_ms", "allreduce_ms", "halo_fraction")]
        t = T()
        self._ck(self.L.mamr_get_device_times(self.h, 1 if wait else 0, C.byref(t)))
        return {n: getattr(t, n) for n, _ in T._fields_}

    def kernel_times(self):
        s, g, c = C.c_float(), C.c_float(), C.c_float()
        ns, ng, nc = C.c_longlong(), C.c_longlong(), C.c_longlong()
        self._ck(self.L.mamr_kernel_time_ms(self.h, C.byref(s), C.byref(g), C.byref(c),
                                            C.byref(ns), C.byref(ng), C.byref(nc)))
        return dict(stencil_ms=s.value, ghost_ms=g.value, checksum_ms=c.value,
                    stencil_launches=ns.value, ghost_launches=ng.value,
                    checksum_launches=nc.value)
