"""farms_b200 -- thin ctypes view of libfarms_b200.so (include/farms_b200.h) for the test and bench harness.

The product is the C-ABI shared library and the FARMS_Flow CLI; this module only marshals numpy arrays
and torch CUDA tensors into that ABI.  There is no fallback: if the library is missing, or no B200 is
visible, calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FARMS_B200_LIB", os.path.join(_HERE, "libfarms_b200.so"))

OK, ERR_ARG, ERR_RANGE, ERR_CUDA, ERR_NOMEM, ERR_STATE, ERR_COMM = 0, -1, -2, -3, -4, -5, -6
COMM_ID_BYTES = 128
COMM_LOCAL = 1
IO_INPUT_ON_DEVICE, IO_OUTPUT_ON_DEVICE = 1, 2
FLAG_DEBUG_DET = 1
FLAG_EXACT_POOLING = 2
FLAG_GENERIC_POOLING = FLAG_EXACT_POOLING
FLAG_SERIAL_SEMANTICS = 4
POOLK_TILE_DENSE, POOLK_TILE_SPARSE, POOLK_TILE_SECOND, POOLK_TILE_ONE_CTA, POOLK_BITS, POOLK_ANY = 1, 2, 4, 8, 16, 32
POOLK_WARP_DENSE, POOLK_WARP_SPARSE, POOLK_WARP_SECOND = 64, 128, 256
POOLK_TILE16_DENSE, POOLK_TILE16_SPARSE, POOLK_TILE16_SECOND, POOLK_TILE16_XCULL = 512, 1024, 2048, 4096
POOL_VARIANTS = {"tile": 1, "bits": 2, "tile1": 3, "warp": 4, "tile16": 5, "tile16x4": 6, "tile16x3": 7, "tile16c": 8}

EXPORTS = [
    "farms_abi_version", "farms_build_is_checked", "farms_create", "farms_destroy", "farms_reset", "farms_last_error", "farms_normalize_filtersize", "farms_get_params",
    "farms_process_host", "farms_process_device", "farms_reserve", "farms_num_events", "farms_get_timings", "farms_set_t0",
    "farms_state_export", "farms_state_fold", "farms_slice_surface", "farms_pack4_f32",
    "farms_slice_surface_host", "farms_state_fold_host",
    "farms_host_alloc", "farms_host_free", "farms_host_register", "farms_host_unregister",
    "farms_comm_unique_id", "farms_comm_create", "farms_comm_destroy", "farms_comm_info", "farms_comm_process", "farms_comm_phases",
]


class Config(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("filtersize", C.c_int32),
                ("inlier_check", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32),
                ("max_batch", C.c_uint64), ("reorder_slack_us", C.c_uint32), ("pool_variant", C.c_uint32),
                ("fit_chunk", C.c_uint32), ("slab_target", C.c_uint32), ("reserved", C.c_uint32 * 4)]


class Out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("t_rel", "global_r", "global_theta", "vx", "vy", "local_r",
                                          "local_theta", "scale", "valid", "best_window", "inliers", "det")]


OUT_DTYPES = {"t_rel": np.uint32, "global_r": np.float64, "global_theta": np.float64, "vx": np.float64,
              "vy": np.float64, "local_r": np.float64, "local_theta": np.float64, "scale": np.uint8,
              "valid": np.uint8, "best_window": np.int8, "inliers": np.uint16, "det": np.float64}


class Gather(C.Structure):
    _fields_ = [("root", C.c_int32), ("dst", C.c_void_p), ("counts", C.c_void_p)]


class Timings(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("h2d_ms", C.c_float), ("ingest_ms", C.c_float),
                ("index_ms", C.c_float), ("fit_ms", C.c_float), ("bin_ms", C.c_float), ("pool_ms", C.c_float),
                ("d2h_ms", C.c_float), ("events", C.c_uint64), ("valid_events", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("pool_candidates", C.c_uint64), ("pool_kernels", C.c_uint64),
                ("pool_events", C.c_uint64 * 3)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "pool_events"}
        d["pool_events_first"], d["pool_events_second"], d["pool_events_general"] = (int(v) for v in self.pool_events)
        return d


_lib = None


def lib():
    """Load the C-ABI library (built in-tree by `make` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.farms_abi_version.restype = C.c_int
        L.farms_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Config)]
        L.farms_destroy.argtypes = [C.c_void_p]
        L.farms_destroy.restype = None
        L.farms_reset.argtypes = [C.c_void_p]
        L.farms_last_error.argtypes = [C.c_void_p]
        L.farms_last_error.restype = C.c_char_p
        L.farms_normalize_filtersize.argtypes = [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.farms_get_params.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
        for f in (L.farms_process_host, L.farms_process_device):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(Out)]
        L.farms_num_events.argtypes = [C.c_void_p]
        L.farms_num_events.restype = C.c_uint64
        L.farms_get_timings.argtypes = [C.c_void_p, C.POINTER(Timings)]
        L.farms_set_t0.argtypes = [C.c_void_p, C.c_uint64]
        L.farms_state_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.farms_state_fold.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.farms_slice_surface.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                          C.c_void_p, C.c_void_p]
        L.farms_pack4_f32.argtypes = [C.c_void_p] * 5 + [C.c_uint64, C.c_void_p]
        L.farms_comm_unique_id.argtypes = [C.c_void_p]
        L.farms_comm_create.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
        L.farms_comm_destroy.argtypes = [C.c_void_p]
        L.farms_comm_destroy.restype = None
        L.farms_comm_info.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 3
        L.farms_comm_process.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                         C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(Out), C.POINTER(Gather)]
        _lib = L
    return _lib


class FarmsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"farms_b200 error {code}: {msg}")
        self.code = code


class Farms:
    """Mirror of `vFlowManager(height, width, filterSize, minEvtsOnPlane, ...)` + `runFileCopy`
    (reference include/vFlow.h:99-104) on top of the C ABI."""

    def __init__(self, width, height, filtersize=3, inlier_check=5, device=0, flags=0, max_batch=0,
                 reorder_slack_us=0, pool_variant=0, fit_chunk=0, slab_target=0):
        L = lib()
        cfg = Config(width=width, height=height, filtersize=filtersize, inlier_check=inlier_check, device=device,
                     flags=flags, max_batch=max_batch, reorder_slack_us=reorder_slack_us,
                     pool_variant=POOL_VARIANTS.get(pool_variant, pool_variant), fit_chunk=fit_chunk,
                     slab_target=slab_target)
        self._h = C.c_void_p()
        rc = L.farms_create(C.byref(self._h), C.byref(cfg))
        if rc != OK:
            raise FarmsError(rc, "farms_create failed (no usable sm_100 CUDA device?)" if rc == ERR_CUDA else "farms_create")
        self.width, self.height, self.device, self.flags = width, height, device, flags

    def close(self):
        if getattr(self, "_h", None):
            lib().farms_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != OK:
            raise FarmsError(rc, lib().farms_last_error(self._h).decode())

    def params(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(lib().farms_get_params(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"filtersize": a.value, "radius": b.value, "plane_size": c.value}

    def reset(self):
        self._check(lib().farms_reset(self._h))

    def set_t0(self, t0):
        self._check(lib().farms_set_t0(self._h, int(t0)))

    def num_events(self):
        return lib().farms_num_events(self._h)

    def timings(self):
        t = Timings()
        self._check(lib().farms_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    def _columns(self, columns):
        cols = list(columns) if columns is not None else [k for k in OUT_DTYPES if k != "det"]
        if columns is None and self.flags & FLAG_DEBUG_DET:
            cols.append("det")
        return cols

    def process(self, x, y, t, p=None, columns=None, out=None):
        """Host arrays in (any integer dtype; copied to u16/u16/u64), dict of numpy columns out."""
        x = np.ascontiguousarray(x, dtype=np.uint16)
        y = np.ascontiguousarray(y, dtype=np.uint16)
        t = np.ascontiguousarray(t, dtype=np.uint64)
        n = len(x)
        assert len(y) == n and len(t) == n
        cols = self._columns(columns)
        res = out if out is not None else {k: np.empty(n, OUT_DTYPES[k]) for k in cols}
        o = Out()
        for k in cols:
            setattr(o, k, res[k].ctypes.data)
        self._check(lib().farms_process_host(self._h, x.ctypes.data, y.ctypes.data, t.ctypes.data, None, n, C.byref(o)))
        return res

    def process_device(self, x, y, t, columns=None, out=None):
        """torch CUDA tensors in (uint16, uint16, uint64 viewed as int64 is fine), dict of CUDA tensors out."""
        import torch
        n = x.numel()
        cols = self._columns(columns)
        tdt = {np.uint32: torch.int32, np.float64: torch.float64, np.uint8: torch.uint8, np.int8: torch.int8,
               np.uint16: torch.int16}
        res = out if out is not None else {k: torch.empty(n, dtype=tdt[OUT_DTYPES[k]], device=x.device) for k in cols}
        o = Out()
        for k in cols:
            setattr(o, k, res[k].data_ptr())
        torch.cuda.current_stream(x.device).synchronize()
        self._check(lib().farms_process_device(self._h, x.data_ptr(), y.data_ptr(), t.data_ptr(), None, n, C.byref(o)))
        return res

    # --- state hand-over (device pointers) ---
    def state_export(self, d_last_t, d_hit):
        self._check(lib().farms_state_export(self._h, d_last_t.data_ptr(), d_hit.data_ptr()))

    def state_fold(self, d_last_t, d_hit):
        self._check(lib().farms_state_fold(self._h, d_last_t.data_ptr(), d_hit.data_ptr()))

    def slice_surface(self, x, y, t, t0, d_last_t, d_hit):
        self._check(lib().farms_slice_surface(self._h, x.data_ptr(), y.data_ptr(), t.data_ptr(), x.numel(), int(t0),
                                              d_last_t.data_ptr(), d_hit.data_ptr()))

    def pack4_f32(self, a, b, c, d, out4):
        """a..d: f64 CUDA tensors of n entries; out4: float32 CUDA tensor of shape (>= n, 4)."""
        self._check(lib().farms_pack4_f32(self._h, a.data_ptr(), b.data_ptr(), c.data_ptr(), d.data_ptr(), a.numel(),
                                          out4.data_ptr()))


def comm_unique_id():
    """128 bytes identifying a new group (rank 0 makes them, every rank gets a copy)."""
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    rc = lib().farms_comm_unique_id(buf)
    if rc != OK:
        raise FarmsError(rc, "farms_comm_unique_id: NCCL is not available")
    return bytes(buf)


class Comm:
    """farms_comm: the time-sliced multi-GPU run of include/farms_b200.h on top of one Farms context per rank."""

    def __init__(self, farms, nranks, rank, unique_id, local=False):
        self.farms, self.nranks, self.rank = farms, nranks, rank
        self._h = C.c_void_p()
        idbuf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id).ljust(COMM_ID_BYTES, b"\0"))
        rc = lib().farms_comm_create(C.byref(self._h), farms._h, nranks, rank, idbuf, COMM_LOCAL if local else 0)
        if rc != OK:
            raise FarmsError(rc, "farms_comm_create: " + lib().farms_last_error(farms._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            lib().farms_comm_destroy(self._h)
            self._h = None

    __del__ = close

    def phases(self):
        """Host wall clock (ms) of the last process() call: slice upload + surface, exchange + fold, event loop, drain."""
        a = (C.c_float * 4)()
        lib().farms_comm_phases.argtypes = [C.c_void_p, C.c_void_p]
        lib().farms_comm_phases(self._h, a)
        return {"surface_ms": a[0], "exchange_ms": a[1], "event_loop_ms": a[2], "drain_ms": a[3]}

    def transport(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        lib().farms_comm_info(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {0: "none", 1: "nccl", 2: "local", 3: "nccl+peer-memory gather"}[c.value]

    def process(self, x, y, t, n_halo, n_surface, t0, out=None, gather_dst=None, root=0, device=False):
        """One collective time-sliced pass.  x, y, t: this rank's slice including its halo -- numpy arrays
        (device=False) or torch CUDA tensors (device=True).  out: dict of columns (numpy / CUDA tensors) for the
        owned events, or None.  gather_dst: float32 CUDA tensor (total owned events, 4) on the root, or None for no
        gather (ranks other than the root pass any non-None placeholder to take part).  Returns the per-rank counts."""
        if device:
            n = x.numel()
            px, py, pt = x.data_ptr(), y.data_ptr(), t.data_ptr()
        else:
            x = np.ascontiguousarray(x, dtype=np.uint16)
            y = np.ascontiguousarray(y, dtype=np.uint16)
            t = np.ascontiguousarray(t, dtype=np.uint64)
            n = len(x)
            px, py, pt = x.ctypes.data, y.ctypes.data, t.ctypes.data
        o = Out()
        flags = IO_INPUT_ON_DEVICE if device else 0
        if out is not None:
            dev_out = any(hasattr(v, "data_ptr") for v in out.values())
            flags |= IO_OUTPUT_ON_DEVICE if dev_out else 0
            for k, v in out.items():
                setattr(o, k, v.data_ptr() if hasattr(v, "data_ptr") else v.ctypes.data)
        counts = np.zeros(self.nranks, np.uint64)
        g = None
        if gather_dst is not None:
            g = Gather(root=root, dst=gather_dst.data_ptr() if self.rank == root else None, counts=counts.ctypes.data)
        rc = lib().farms_comm_process(self._h, px, py, pt, n, int(n_halo), int(n_surface), int(t0), flags,
                                      C.byref(o) if out is not None else None, C.byref(g) if g is not None else None)
        self.farms._check(rc)
        return counts
