"""ctypes view of tools/libfarms_synth.so (tools/farms_synth.h): deterministic synthetic event streams."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(_HERE, "libfarms_synth.so"))
        L.farms_synth_open.restype = C.c_void_p
        L.farms_synth_open.argtypes = [C.c_int, C.c_uint64]
        L.farms_synth_close.argtypes = [C.c_void_p]
        L.farms_synth_info.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.farms_synth_range.restype = C.c_int64
        L.farms_synth_range.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64] + [C.c_void_p] * 4 + [C.c_int64, C.c_int]
        _lib = L
    return _lib


class Synth:
    def __init__(self, config, seed=0):
        self._h = lib().farms_synth_open(config, seed)
        if not self._h:
            raise ValueError(f"unknown synthetic config {config}")
        w, h, fs, r = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        lib().farms_synth_info(self._h, C.byref(w), C.byref(h), C.byref(fs), C.byref(r))
        self.width, self.height, self.filtersize, self.rate = w.value, h.value, fs.value, r.value
        self.config = config

    def __del__(self):
        if getattr(self, "_h", None):
            lib().farms_synth_close(self._h)
            self._h = None

    def time_range(self, t_begin_us, t_end_us, nthreads=0, pinned=False):
        """All events with stream time in [t_begin_us, t_end_us): (x u16, y u16, t u64, p u8)."""
        cap = int(self.rate * (t_end_us - t_begin_us) * 1e-6 * 1.15) + 65536
        while True:
            if pinned:
                import torch
                bufs = [torch.empty(cap, dtype=d).pin_memory() for d in (torch.uint16, torch.uint16, torch.int64, torch.uint8)]
                arrs = [b.numpy() for b in bufs]
                arrs[2] = arrs[2].view(np.uint64)
            else:
                arrs = [np.empty(cap, d) for d in (np.uint16, np.uint16, np.uint64, np.uint8)]
            n = lib().farms_synth_range(self._h, int(t_begin_us), int(t_end_us), *[a.ctypes.data for a in arrs], cap, nthreads)
            if n >= 0:
                return tuple(a[:n] for a in arrs)
            cap = -n + 1024

    def first(self, n, t_begin_us=0, nthreads=0, pinned=False):
        """The first n events at or after t_begin_us."""
        span = int(n / self.rate * 1e6 * 1.05) + 2048
        while True:
            x, y, t, p = self.time_range(t_begin_us, t_begin_us + span, nthreads, pinned)
            if len(x) >= n:
                return x[:n], y[:n], t[:n], p[:n]
            span = int(span * 1.5) + 4096
