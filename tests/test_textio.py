"""CPU tests of the multi-threaded text reader/writer (include/farms_textio.h) against plain Python statements of
the reference's reader (src/vFlow.cpp:173-188) and writer (src/vFlow.cpp:436-440)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT

PKG = os.path.join(ROOT, "aperture-robust-multiscale-optical-flow_b200")
LIB = os.path.join(PKG, "libfarms_textio.so")


class Events(C.Structure):
    _fields_ = [("n", C.c_uint64), ("x", C.POINTER(C.c_uint16)), ("y", C.POINTER(C.c_uint16)),
                ("t", C.POINTER(C.c_uint64)), ("xi", C.POINTER(C.c_int32)), ("yi", C.POINTER(C.c_int32)),
                ("pol", C.POINTER(C.c_int32))]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", PKG, "libfarms_textio.so"])
    L = C.CDLL(LIB)
    L.farms_text_read.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(Events), C.c_char_p, C.c_size_t]
    L.farms_text_free.argtypes = [C.POINTER(Events)]
    L.farms_text_write.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64] + [C.c_void_p] * 11 + [C.c_int]
    return L


def read(lib, path, maxn=2**63, threads=0):
    ev = Events()
    err = C.create_string_buffer(256)
    rc = lib.farms_text_read(path.encode(), maxn, threads, C.byref(ev), err, 256)
    if rc != 0:
        return rc, err.value.decode()
    n = ev.n
    out = {k: np.ctypeslib.as_array(getattr(ev, k), shape=(max(n, 1),))[:n].copy() for k in ("x", "y", "t", "xi", "yi", "pol")}
    lib.farms_text_free(C.byref(ev))
    return 0, out


def reference_reader(text, maxn):
    """`while (getline(f, line) && n < N) { stream >> x >> y >> time_ >> pol; push }` with stale values."""
    x = y = t = p = 0
    rows = []
    lines = text.split("\n")
    if lines and lines[-1] == "":
        lines = lines[:-1]
    for ln in lines:
        if len(rows) >= maxn:
            break
        vals = []
        for tok in ln.split():
            try:
                vals.append(int(tok))
            except ValueError:
                break
        if len(vals) > 0: x = vals[0]
        if len(vals) > 1: y = vals[1]
        if len(vals) > 2: t = vals[2]
        if len(vals) > 3: p = vals[3]
        rows.append((x, y, t, max(p, 0)))
    return np.array(rows, dtype=np.int64).reshape(-1, 4)


def test_reader_matches_reference_semantics(lib, tmp_path):
    rng = np.random.default_rng(3)
    n = 300_000
    a = np.stack([rng.integers(0, 1280, n), rng.integers(0, 720, n), np.sort(rng.integers(1000, 4_000_000_000, n)),
                  rng.integers(-1, 2, n)], 1)
    lines = [" ".join(map(str, r)) for r in a]
    # blank lines, short lines, garbage, CRLF, tabs, no trailing newline
    lines[5] = ""
    lines[77] = "12 34"
    lines[1000] = "7 8 zzz 1"
    lines[2000] = lines[2000].replace(" ", "\t") + "\r"
    lines[150_000] = ""
    text = "\n".join(lines)
    path = str(tmp_path / "ev.txt")
    open(path, "w").write(text)
    want = reference_reader(text, 10**18)
    for threads in (1, 3, 8):
        rc, got = read(lib, path, threads=threads)
        assert rc == 0
        assert np.array_equal(got["xi"], want[:, 0]) and np.array_equal(got["yi"], want[:, 1])
        assert np.array_equal(got["t"], want[:, 2].astype(np.uint64)) and np.array_equal(got["pol"], want[:, 3])
        assert np.array_equal(got["x"], want[:, 0].astype(np.uint16))
    rc, got = read(lib, path, maxn=12345, threads=4)   # --numEvents
    assert rc == 0 and len(got["x"]) == 12345 and np.array_equal(got["yi"], want[:12345, 1])


def test_reader_errors_and_empty(lib, tmp_path):
    rc, msg = read(lib, str(tmp_path / "missing.txt"))
    assert rc == -1 and "Unable to open" in msg
    p = tmp_path / "bad.txt"
    p.write_text("1 2 3 1\n70000 2 4 1\n")
    rc, msg = read(lib, str(p))
    assert rc == -1 and "event 1" in msg
    e = tmp_path / "empty.txt"
    e.write_text("")
    rc, got = read(lib, str(e))
    assert rc == 0 and len(got["x"]) == 0


def test_writer_is_byte_identical_to_percent_g(lib, tmp_path):
    rng = np.random.default_rng(4)
    n = 200_000
    xi = rng.integers(0, 1280, n).astype(np.int32)
    yi = rng.integers(0, 720, n).astype(np.int32)
    tr = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)   # values >= 2^31 print negative like vector<int>
    pol = rng.integers(0, 2, n).astype(np.int32)
    cols = [np.where(rng.random(n) < 0.4, 0.0, rng.standard_normal(n) * 10.0 ** rng.integers(-8, 9, n)) for _ in range(6)]
    cols[2][:4] = [np.inf, -np.inf, 0.0, 1e-300]
    cols[3][:4] = [np.nan, 1e300, 123456.5, 0.1]
    scale = (rng.integers(0, 11, n) * 5).astype(np.uint8)
    p11, p8 = str(tmp_path / "o11.txt"), str(tmp_path / "o8.txt")
    args = [xi, yi, tr, pol] + cols + [scale]
    for threads in (1, 5):
        rc = lib.farms_text_write(p11.encode(), p8.encode(), n, *[a.ctypes.data for a in args], threads)
        assert rc == 0
        got11 = open(p11).read().splitlines()
        got8 = open(p8).read().splitlines()
        assert len(got11) == n and len(got8) == n
        ti = tr.astype(np.int32)
        for i in list(range(50)) + list(rng.integers(0, n, 3000)):
            d = ["%g" % c[i] for c in cols]
            if np.isnan(cols[3][i]):
                assert got11[i].split()[7] in ("nan", "-nan")
                continue
            assert got11[i] == f"{xi[i]} {yi[i]} {ti[i]} {pol[i]} {d[0]} {d[1]} {d[2]} {d[3]} {d[4]} {d[5]} {scale[i]}"
            assert got8[i] == f"{xi[i]} {yi[i]} {ti[i]} {pol[i]} {d[0]} {d[1]} {d[4]} {d[5]}"


def test_binary_side_format_round_trip(lib, tmp_path):
    """include/farms_textio.h binary side-format: events written, read back (with a count limit),
    results written and decoded column by column."""
    lib.farms_bin_read.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(Events), C.c_char_p, C.c_size_t]
    lib.farms_bin_write_events.argtypes = [C.c_char_p, C.c_uint64] + [C.c_void_p] * 4
    lib.farms_bin_write.argtypes = [C.c_char_p, C.c_uint64] + [C.c_void_p] * 11
    rng = np.random.default_rng(3)
    n = 1000
    x = rng.integers(0, 1280, n).astype(np.uint16)
    y = rng.integers(0, 720, n).astype(np.uint16)
    t = np.sort(rng.integers(1000, 2**40, n)).astype(np.uint64)
    p = rng.integers(0, 2, n).astype(np.uint8)
    path = str(tmp_path / "ev.evb")
    assert lib.farms_bin_write_events(path.encode(), n, x.ctypes.data, y.ctypes.data, t.ctypes.data, p.ctypes.data) == 0
    for take in (n, 137, 0):
        ev = Events()
        err = C.create_string_buffer(256)
        assert lib.farms_bin_read(path.encode(), take, C.byref(ev), err, 256) == 0, err.value
        assert ev.n == take
        for k, ref in (("x", x), ("y", y), ("t", t), ("xi", x), ("yi", y), ("pol", p)):
            got = np.ctypeslib.as_array(getattr(ev, k), shape=(max(take, 1),))[:take]
            assert np.array_equal(got, ref[:take].astype(got.dtype)), k
        lib.farms_text_free(C.byref(ev))
    bad = str(tmp_path / "bad.evb")
    open(bad, "wb").write(b"not an event file")
    ev = Events()
    err = C.create_string_buffer(256)
    assert lib.farms_bin_read(bad.encode(), n, C.byref(ev), err, 256) == -1 and b"FARMSEV1" in err.value
    # results
    cols = [rng.standard_normal(n) for _ in range(6)]
    t_rel = (t - t[0]).astype(np.uint32)
    scale = (rng.integers(0, 11, n) * 5).astype(np.uint8)
    xi, yi, pol = x.astype(np.int32), y.astype(np.int32), p.astype(np.int32)
    out = str(tmp_path / "res.bin")
    assert lib.farms_bin_write(out.encode(), n, xi.ctypes.data, yi.ctypes.data, t_rel.ctypes.data, pol.ctypes.data,
                               *[c.ctypes.data for c in cols], scale.ctypes.data) == 0
    raw = open(out, "rb").read()
    assert raw[:8] == b"FARMSOU1" and int(np.frombuffer(raw, np.uint64, 1, 8)[0]) == n
    off = 16
    for ref, dt in ((x, np.uint16), (y, np.uint16), (t_rel, np.uint32), (p, np.uint8), (scale, np.uint8)):
        got = np.frombuffer(raw, dt, n, off)
        assert np.array_equal(got, ref.astype(dt))
        off += n * np.dtype(dt).itemsize
    for c in cols:
        assert np.array_equal(np.frombuffer(raw, np.float64, n, off), c)
        off += 8 * n
    assert off == len(raw)
