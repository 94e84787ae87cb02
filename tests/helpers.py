"""Shared test helpers: oracle binding (checker only), stream builders, comparison rules."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# ---------------------------------------------------------------------------------------------------
# tier-2 oracle (oracle/farms_oracle.c) -- TEST INFRASTRUCTURE, never imported by the product
# ---------------------------------------------------------------------------------------------------


class _OracleOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("t_rel", "pol", "global_r", "global_theta", "vx", "vy", "local_r",
                                          "local_theta", "scale", "valid", "best_window", "inliers", "det")]


_ORACLE_DTYPES = {"t_rel": np.int32, "pol": np.int32, "global_r": np.float64, "global_theta": np.float64,
                  "vx": np.float64, "vy": np.float64, "local_r": np.float64, "local_theta": np.float64,
                  "scale": np.int32, "valid": np.uint8, "best_window": np.int8, "inliers": np.int32,
                  "det": np.float64}
_olib = None


def oracle_lib():
    global _olib
    if _olib is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L.farms_oracle_create.restype = C.c_void_p
        L.farms_oracle_create.argtypes = [C.c_int] * 4
        L.farms_oracle_destroy.argtypes = [C.c_void_p]
        L.farms_oracle_process.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_uint64, C.POINTER(_OracleOut)]
        L.farms_oracle_state.argtypes = [C.c_void_p] * 5
        L.farms_oracle_set_fast.argtypes = [C.c_void_p, C.c_int]
        L.farms_oracle_is_fast.argtypes = [C.c_void_p]
        L.farms_oracle_set_serial.argtypes = [C.c_void_p, C.c_int]
        _olib = L
    return _olib


class Oracle:
    def __init__(self, width, height, filtersize, inlier_check, fast=False, serial=False):
        """fast=True: the oracle's filtered pooling walk (oracle/farms_oracle.h: bit-identical outputs, ~10-20x
        faster on dense streams; the plain scan stays the witness it is checked against)."""
        self._h = oracle_lib().farms_oracle_create(width, height, filtersize, inlier_check)
        assert self._h
        self.npx = width * height
        if fast:
            assert oracle_lib().farms_oracle_set_fast(self._h, 1) == 0
        if serial:
            assert oracle_lib().farms_oracle_set_serial(self._h, 1) == 0

    def is_fast(self):
        return bool(oracle_lib().farms_oracle_is_fast(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            oracle_lib().farms_oracle_destroy(self._h)
            self._h = None

    def process(self, x, y, t, p=None):
        n = len(x)
        x = np.ascontiguousarray(x, np.int32)
        y = np.ascontiguousarray(y, np.int32)
        t = np.ascontiguousarray(np.asarray(t).astype(np.uint64) & 0xFFFFFFFF, np.uint32)
        p = np.ascontiguousarray(p if p is not None else np.ones(n), np.int32)
        res = {k: np.empty(n, d) for k, d in _ORACLE_DTYPES.items()}
        o = _OracleOut()
        for k in res:
            setattr(o, k, res[k].ctypes.data)
        rc = oracle_lib().farms_oracle_process(self._h, x.ctypes.data, y.ctypes.data, t.ctypes.data, p.ctypes.data, n,
                                               C.byref(o))
        assert rc == 0, "oracle: event outside the sensor"
        return res

    def state(self):
        lt = np.empty(self.npx, np.float64)
        hit = np.empty(self.npx, np.uint8)
        oracle_lib().farms_oracle_state(self._h, lt.ctypes.data, hit.ctypes.data, None, None)
        return lt, hit


def run_oracle(width, height, filtersize, inlier_check, x, y, t, p=None, fast=False, serial=False):
    return Oracle(width, height, filtersize, inlier_check, fast=fast, serial=serial).process(x, y, t, p)


# ---------------------------------------------------------------------------------------------------
# comparison rules (BASELINE.json north_star): masks / window / inliers bit-exact; R within 1e-4
# relative; angles within 1e-3 rad; divergent events counted and reported.
# ---------------------------------------------------------------------------------------------------
R_RTOL = 1e-4
ANGLE_ATOL = 1e-3


def angle_diff(a, b):
    d = np.abs(a - b) % (2 * np.pi)
    return np.minimum(d, 2 * np.pi - d)


def compare(got, ref, what=""):
    """Return a dict of divergence counts between a product result and the oracle's."""
    n = len(ref["valid"])
    rep = {"n": n, "valid_ref": int(ref["valid"].sum())}
    rep["t_rel"] = int(np.count_nonzero(got["t_rel"].astype(np.uint32) != ref["t_rel"].astype(np.uint32)))
    rep["valid"] = int(np.count_nonzero(got["valid"] != ref["valid"]))
    rep["best_window"] = int(np.count_nonzero(got["best_window"].astype(np.int32) != ref["best_window"].astype(np.int32)))
    rep["inliers"] = int(np.count_nonzero(got["inliers"].astype(np.int64) != ref["inliers"].astype(np.int64)))
    v = ref["valid"].astype(bool) & got["valid"].astype(bool)

    def rel_bad(a, b):
        with np.errstate(invalid="ignore", divide="ignore"):
            same = (a == b) | (np.isnan(a) & np.isnan(b))
            bad = ~same & ~(np.abs(a - b) <= R_RTOL * np.abs(b))
        return bad

    # Vx, Vy (columns 7, 8) are components of one vector: a component that is ~1e-16 of the length (flow along
    # an axis: cos(pi/2) = 6e-17) carries no digits, so each is checked against 1e-4 of the vector's length.
    with np.errstate(invalid="ignore", over="ignore"):
        mag = np.hypot(ref["vx"], ref["vy"])

    def comp_bad(a, b):
        with np.errstate(invalid="ignore"):
            same = (a == b) | (np.isnan(a) & np.isnan(b))
            return ~same & ~(np.abs(a - b) <= R_RTOL * mag)

    rep["vx"] = int(np.count_nonzero(comp_bad(got["vx"], ref["vx"])))
    rep["vy"] = int(np.count_nonzero(comp_bad(got["vy"], ref["vy"])))
    rep["local_r"] = int(np.count_nonzero(rel_bad(got["local_r"], ref["local_r"])))
    rep["local_theta"] = int(np.count_nonzero(angle_diff(got["local_theta"], ref["local_theta"])[v] > ANGLE_ATOL))
    rep["global_r"] = int(np.count_nonzero(rel_bad(got["global_r"], ref["global_r"])))
    rep["global_theta"] = int(np.count_nonzero(angle_diff(got["global_theta"], ref["global_theta"])[v] > ANGLE_ATOL))
    rep["scale"] = int(np.count_nonzero(got["scale"].astype(np.int32) != ref["scale"]))
    with np.errstate(invalid="ignore", divide="ignore"):
        rr = np.abs(got["global_r"] - ref["global_r"])[v] / np.abs(ref["global_r"][v]) if v.any() else np.zeros(1)
    rep["max_rel_global_r"] = float(np.nanmax(rr)) if rr.size else 0.0
    rep["what"] = what
    return rep


def assert_parity(rep, allow_scale_flips=0):
    exact = ["t_rel", "valid", "best_window", "inliers"]
    tol = ["vx", "vy", "local_r", "local_theta"]
    for k in exact + tol:
        assert rep[k] == 0, f"{rep['what']}: {k} diverges on {rep[k]} of {rep['n']} events: {rep}"
    # a flipped scale between near-tied means is the only tolerated divergence, and it is counted
    assert rep["scale"] <= allow_scale_flips, f"{rep['what']}: {rep}"
    assert rep["global_r"] <= rep["scale"] and rep["global_theta"] <= rep["scale"], f"{rep['what']}: {rep}"


def synth_stream(config, n, t_begin_us=0):
    from farms_synth import Synth
    s = Synth(config)
    x, y, t, p = s.first(n, t_begin_us)
    return s, x, y, t, p
