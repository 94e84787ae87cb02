"""FARMS_Flow --gpus N on a multi-GPU box against the one-GPU run (binary side-format, synthetic 1280x720 stream)."""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from farms_synth import Synth

PKG = os.path.join(ROOT, "aperture-robust-multiscale-optical-flow_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
gpus = int(sys.argv[2]) if len(sys.argv) > 2 else 2
syn = Synth(4)
x, y, t, p = syn.first(n)
L = C.CDLL(os.path.join(PKG, "libfarms_textio.so"))
L.farms_bin_write_events.argtypes = [C.c_char_p, C.c_uint64] + [C.c_void_p] * 4


def load(path):
    raw = np.fromfile(path, np.uint8)
    cnt = int(raw[8:16].view(np.uint64)[0])
    off, cols = 16, {}
    for k, dt in (("x", np.uint16), ("y", np.uint16), ("t", np.uint32), ("p", np.uint8), ("scale", np.uint8),
                  ("gr", np.float64), ("gth", np.float64), ("vx", np.float64), ("vy", np.float64),
                  ("lr", np.float64), ("lth", np.float64)):
        nb = cnt * np.dtype(dt).itemsize
        cols[k] = raw[off:off + nb].view(dt)
        off += nb
    return cols


with tempfile.TemporaryDirectory() as d:
    res = {}
    for g in (1, gpus):
        base = os.path.join(d, f"ev{g}")
        assert L.farms_bin_write_events((base + ".evb").encode(), n, x.ctypes.data, y.ctypes.data, t.ctypes.data,
                                        p.ctypes.data) == 0
        t0 = time.time()
        out = subprocess.run([os.path.join(PKG, "FARMS_Flow"), "--width", "1280", "--height", "720", "--filtersize", "5",
                              "--filename", base, "--binary", "1", "--SERIAL", "0", "--gpus", str(g)], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        print(f"--gpus {g}: wall {time.time() - t0:.2f} s;", [ln for ln in out.stdout.splitlines() if "Benchmark" in ln or "farms_b200" in ln])
        res[g] = load(base + "_FARMSOut_.bin")
    a, b = res[1], res[gpus]
    for k in ("x", "y", "t", "p", "scale", "vx", "vy", "lr", "lth"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    v = a["lr"] > 0
    rel = np.abs(a["gr"][v] - b["gr"][v]) / np.abs(a["gr"][v])
    print(f"events {n}, with flow {int(v.sum())}, scale identical, max rel diff of globalR {rel.max():.2e}")
    assert rel.max() < 1e-4
