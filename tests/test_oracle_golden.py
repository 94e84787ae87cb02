"""The CPU oracle (oracle/farms_oracle.c) pinned against golden vectors produced by the reference's own
sources (tests/golden/make_golden.py -> oracle/_ref/FARMS_Flow).  CPU only."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from helpers import ROOT, run_oracle
from kat_streams import (FIT_SYNTH_CASES, HASH_CASES, LONG_SYNTH_CASES, SERIAL_CASES, SERIAL_SYNTH_CASES, SYNTH_CASES,
                         TEXT_CASES, sweeps, write_txt)

GOLDEN = os.path.join(ROOT, "tests", "golden")
META = json.load(open(os.path.join(GOLDEN, "golden.json")))
META_SERIAL = json.load(open(os.path.join(GOLDEN, "golden_serial.json")))
META_FIT = json.load(open(os.path.join(GOLDEN, "golden_fit.json")))
CLI = os.path.join(ROOT, "oracle", "farms_oracle_cli")


@pytest.fixture(scope="module", autouse=True)
def _cli():
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "farms_oracle_cli"])


def oracle_text(tmp_path, name, w, h, fs, inl, x, y, t, p, fast=False, serial=False):
    base = str(tmp_path / name)
    write_txt(base + ".txt", x, y, t, p)
    subprocess.run([CLI, str(w), str(h), str(fs), str(inl), base], check=True, capture_output=True,
                   env=dict(os.environ, FARMS_ORACLE_FAST="1" if fast else "0", FARMS_ORACLE_SERIAL="1" if serial else "0"))
    return open(base + "_FARMSOut_oracle.txt", "rb").read()


@pytest.mark.parametrize("name", sorted(TEXT_CASES))
def test_text_golden(name, tmp_path):
    w, h, fs, inl, build = TEXT_CASES[name]
    got = oracle_text(tmp_path, name, w, h, fs, inl, *build()).decode().splitlines()
    ref = open(os.path.join(GOLDEN, name + ".ref.txt")).read().splitlines()
    assert len(got) == len(ref) == META[name]["rows"]
    if name == "kat_plane_16x12":
        # width > height on a heap-allocated surface: the reference reads past the end of its vectors
        # (src/vFlow.cpp:1000-1002) and its scale choice for 5 early events depends on heap contents.
        # Everything but the scale column must still agree; the scale column may differ on those rows only.
        diff = [i for i, (a, b) in enumerate(zip(got, ref)) if a != b]
        assert all(got[i].split()[:10] == ref[i].split()[:10] for i in diff)
        assert len(diff) <= 5 and all(int(ref[i].split()[2]) < 500 for i in diff)
    else:
        assert got == ref


def test_appendix_b_known_answer():
    """analytic plane a = 1e-4 s/px, b = 3e-5 s/px on a 16x12 sensor."""
    w, h, fs, inl, build = TEXT_CASES["kat_plane_16x12"]
    x, y, t, p = build()
    r = run_oracle(w, h, fs, inl, x, y, t, p)
    assert int(r["valid"].sum()) == 184
    inval = sorted((int(a), int(b)) for a, b, v in zip(x, y, r["valid"]) if not v)
    assert inval == [(0, 0), (0, 1), (0, 2), (0, 3), (1, 0), (1, 1), (1, 2), (1, 3)]
    v = r["valid"].astype(bool)
    a, b = 1e-4, 3e-5
    assert np.allclose(r["vx"][v], b / (a * a + b * b), rtol=1e-9)   # axes are swapped in the reference
    assert np.allclose(r["vy"][v], a / (a * a + b * b), rtol=1e-9)
    assert np.allclose(r["local_r"][v], 9578.26, rtol=1e-6)
    assert np.allclose(r["local_theta"][v], 1.27934, atol=1e-5)
    # pixel (0,4): exactly 5 inliers >= inlierCheck 5; pixel (1,1): singular AtA => DET < 1
    i04 = [i for i in range(len(x)) if (x[i], y[i]) == (0, 4)][0]
    i11 = [i for i in range(len(x)) if (x[i], y[i]) == (1, 1)][0]
    assert r["inliers"][i04] == 5 and r["valid"][i04] == 1
    assert r["det"][i11] < 1 and r["valid"][i11] == 0


@pytest.mark.parametrize("name", sorted(HASH_CASES))
def test_hash_golden(name, tmp_path):
    w, h, fs, inl, build = HASH_CASES[name]
    raw = oracle_text(tmp_path, name, w, h, fs, inl, *build())
    assert hashlib.sha256(raw).hexdigest() == META[name]["sha256"]


@pytest.mark.parametrize("name", sorted(SYNTH_CASES))
def test_synthetic_scene_golden(name, tmp_path):
    from farms_synth import Synth
    cfg, n, start = SYNTH_CASES[name]
    s = Synth(cfg)
    x, y, t, p = s.first(n, start)
    inp = np.stack([x.astype(np.int64), y.astype(np.int64), t.astype(np.int64), p.astype(np.int64)], 1)
    assert hashlib.sha256(inp.tobytes()).hexdigest() == META[name]["input_sha256"], "generator changed"
    raw = oracle_text(tmp_path, name, s.width, s.height, s.filtersize, 5, x, y, t, p)
    assert hashlib.sha256(raw).hexdigest() == META[name]["sha256"]
    assert raw.decode().splitlines()[:0] == []
    rows = raw.decode().splitlines()
    assert sum(1 for r in rows if float(r.split()[4]) > 0) == META[name]["valid"]


def test_oracle_streaming_equals_one_shot():
    from helpers import Oracle
    w, h, fs, inl, build = HASH_CASES["sweeps_160x120_fs5"]
    x, y, t, p = build()
    x, y, t, p = x[:20000], y[:20000], t[:20000], p[:20000]
    full = run_oracle(w, h, fs, inl, x, y, t, p)
    o = Oracle(w, h, fs, inl)
    # t0 is the first timestamp ever seen, so later pieces must be rebased by the caller: feed raw times
    parts = []
    o2 = Oracle(w, h, fs, inl)
    for a, b in [(0, 7000), (7000, 7001), (7001, 20000)]:
        parts.append(o2.process(x[a:b], y[a:b], t[a:b], p[a:b]))
    for k in ("valid", "inliers", "scale", "global_r", "vx"):
        assert np.array_equal(np.concatenate([q[k] for q in parts]), full[k], equal_nan=(k in ("global_r", "vx")))


@pytest.mark.parametrize("name", sorted(LONG_SYNTH_CASES))
def test_long_prefix_golden_with_the_fast_oracle(name, tmp_path):
    """Steady-state prefixes of the benchmark scenes (1-2 M events): the reference's own output (SHA-256 of its
    11-column text, recorded by make_golden.py) against the tier-2 oracle in its fast pooling mode."""
    from farms_synth import Synth
    cfg, n, start = LONG_SYNTH_CASES[name]
    s = Synth(cfg)
    x, y, t, p = s.first(n, start)
    inp = np.stack([x.astype(np.int64), y.astype(np.int64), t.astype(np.int64), p.astype(np.int64)], 1)
    assert hashlib.sha256(inp.tobytes()).hexdigest() == META[name]["input_sha256"], "generator changed"
    raw = oracle_text(tmp_path, name, s.width, s.height, s.filtersize, 5, x, y, t, p, fast=True)
    assert hashlib.sha256(raw).hexdigest() == META[name]["sha256"]
    # the point of these cases: most of the stream is in the steady state (valid fraction of the scene)
    assert META[name]["valid"] > 0.3 * n


@pytest.mark.parametrize("cfg,n,start", [(1, 60000, 0), (2, 80000, 30000), (3, 80000, 0), (4, 250000, 4000)])
def test_fast_oracle_mode_is_bit_identical_to_the_plain_scan(cfg, n, start):
    from farms_synth import Synth
    from helpers import Oracle
    s = Synth(cfg)
    x, y, t, p = s.first(n, start)
    slow = run_oracle(s.width, s.height, s.filtersize, 5, x, y, t, p)
    o = Oracle(s.width, s.height, s.filtersize, 5, fast=True)
    fast = o.process(x, y, t, p)
    assert o.is_fast()
    for k in slow:
        assert np.array_equal(slow[k].view(np.uint8), fast[k].view(np.uint8)), k


def test_fast_oracle_mode_falls_back_on_decreasing_timestamps():
    from farms_synth import Synth
    from helpers import Oracle
    s = Synth(1)
    x, y, t, p = s.first(30000, 0)
    rng = np.random.default_rng(11)
    t2 = t.astype(np.int64)
    t2[10000:] += rng.integers(-200, 200, len(t) - 10000)
    t2 = t2.astype(np.uint64)
    slow = run_oracle(s.width, s.height, 5, 5, x, y, t2, p)
    o = Oracle(s.width, s.height, 5, 5, fast=True)
    fast = o.process(x, y, t2, p)
    assert not o.is_fast()
    for k in slow:
        assert np.array_equal(slow[k].view(np.uint8), fast[k].view(np.uint8)), k


def test_fast_oracle_mode_on_aliasing_shapes():
    """width > height (flat-index aliasing of vFlow.cpp:1000) and width < height (row bound below the sensor)."""
    for (w, h) in [(64, 20), (150, 40), (20, 64), (48, 200)]:
        x, y, t, p = sweeps(w, h, slopes=((7, 3), (-5, 6)), gap=100)
        slow = run_oracle(w, h, 5, 5, x, y, t, p)
        fast = run_oracle(w, h, 5, 5, x, y, t, p, fast=True)
        for k in slow:
            assert np.array_equal(slow[k].view(np.uint8), fast[k].view(np.uint8)), (w, h, k)


def _pool_numpy(w, h, x, y, t_rel, lr, lth, valid, serial, first_raw_t=None):
    """Independent numpy restatement of computeTrueFlow over per-event local flows (vFlow.cpp:952-1094), in the
    batch or the serial update order of lastEventTime.  Small streams only."""
    length = np.zeros((w * h,))
    theta = np.zeros((w * h,))
    last = np.zeros((w * h,))
    out_r, out_th, out_scale = np.zeros(len(x)), np.zeros(len(x)), np.zeros(len(x), np.int32)
    npx = w * h
    for e in range(len(x)):
        f = int(x[e]) * h + int(y[e])
        if serial and e == 0:
            last[f] = first_raw_t
            continue
        te = float(t_rel[e])
        if not serial:
            last[f] = te
        if valid[e]:
            length[f], theta[f] = lr[e], lth[e]
            best, bestv = 0.0, None
            for k, s in enumerate(range(0, 51, 5)):
                sl = sx = sy = n = 0.0
                for i in range(max(0, x[e] - s), min(x[e] + s, w - 1) + 1):
                    for j in range(max(0, y[e] - s), min(y[e] + s, w - 1) + 1):
                        g = i * h + j
                        if g >= npx:
                            continue
                        if length[g] > 0 and abs(te - last[g]) < 500:
                            sl += length[g]
                            sx += length[g] * np.cos(theta[g])
                            sy += length[g] * np.sin(theta[g])
                            n += 1
                if n > 0 and sl / n > best:
                    best, bestv = sl / n, (sx / n, sy / n, s)
            if bestv is None:
                bestv = (length[f] * np.cos(theta[f]), length[f] * np.sin(theta[f]), 0)
            out_r[e] = np.hypot(bestv[0], bestv[1])
            out_th[e] = np.arctan2(bestv[1], bestv[0])
            out_scale[e] = bestv[2]
        else:
            length[f] = theta[f] = 0.0
        last[f] = te
    return out_r, out_th, out_scale


@pytest.mark.parametrize("serial", [False, True])
def test_serial_mode_of_the_oracle_against_an_independent_pooling(serial):
    """vFlowManager::run semantics (vFlow.cpp:465-826): first event only sets t0 and leaves its raw timestamp in
    lastEventTime; lastEventTime is written after pooling.  The oracle's serial mode is checked against a numpy
    restatement of the pooling fed with the oracle's own per-event local flows."""
    from helpers import Oracle
    w, h = 40, 30
    x, y, t, p = sweeps(w, h, slopes=((25, 9), (-6, 9)), jitter=2, gap=120)
    n = 700
    x, y, t, p = x[:n], y[:n], t[:n], p[:n]
    r = Oracle(w, h, 5, 5, serial=serial).process(x, y, t, p)
    if serial:
        assert r["valid"][0] == 0 and r["t_rel"][0] == 0 and r["best_window"][0] == -1
    gr, gth, sc = _pool_numpy(w, h, x, y, r["t_rel"], r["local_r"], r["local_theta"], r["valid"], serial, float(t[0]))
    v = r["valid"].astype(bool)
    assert v.sum() > 200
    assert np.allclose(r["global_r"][v], gr[v], rtol=1e-12)
    assert np.allclose(r["global_theta"][v], gth[v], atol=1e-12)
    assert np.array_equal(r["scale"][v], sc[v])
    if serial:
        # the serial order changes results: events whose pixel was last hit more than 500 us ago do not pool themselves
        b = Oracle(w, h, 5, 5).process(x, y, t, p)
        both = v & b["valid"].astype(bool)
        assert np.count_nonzero(r["global_r"][both] != b["global_r"][both]) > 10


# ---------------------------------------------------------------------------------------------------
# The serial mode against the reference's DEFAULT driver: vFlowManager::run writes no file, so its results were
# recorded through the call probe of oracle/serial_probe.cpp (tests/golden/make_golden_serial.py).
# ---------------------------------------------------------------------------------------------------
def _serial_rows(tmp_path, name, w, h, fs, inl, x, y, t, p):
    got = oracle_text(tmp_path, name, w, h, fs, inl, x, y, t, p, serial=True).decode().splitlines()
    assert got[0].split()[4:] == ["0"] * 7  # the line that only sets t0 (src/vFlow.cpp:531-558)
    k = META_SERIAL[name]["rows"]           # run() stops at numEvents <= filesize / 18 (:511); the oracle does not
    return got[1:1 + k]


@pytest.mark.parametrize("name", sorted(SERIAL_CASES))
def test_serial_mode_golden(name, tmp_path):
    """Event by event what the reference's own computeLocalFlow / computeTrueFlow returned inside run()."""
    w, h, fs, inl, build = SERIAL_CASES[name]
    got = _serial_rows(tmp_path, name, w, h, fs, inl, *build())
    m = META_SERIAL[name]
    assert len(got) == m["rows"]
    assert sum(1 for r in got if float(r.split()[4]) > 0) == m["valid"] > 0.5 * m["rows"]
    path = os.path.join(GOLDEN, name + ".ref.txt")
    if os.path.exists(path):
        assert got == open(path).read().splitlines()
    assert hashlib.sha256(("\n".join(got) + "\n").encode()).hexdigest() == m["sha256"]


@pytest.mark.parametrize("name", sorted(SERIAL_SYNTH_CASES))
def test_serial_mode_golden_on_a_benchmark_scene(name, tmp_path):
    from farms_synth import Synth
    cfg, n, start = SERIAL_SYNTH_CASES[name]
    s = Synth(cfg)
    x, y, t, p = s.first(n, start)
    inp = np.stack([x.astype(np.int64), y.astype(np.int64), t.astype(np.int64), p.astype(np.int64)], 1)
    m = META_SERIAL[name]
    assert hashlib.sha256(inp.tobytes()).hexdigest() == m["input_sha256"], "generator changed"
    got = _serial_rows(tmp_path, name, s.width, s.height, s.filtersize, 5, x, y, t, p)
    assert sum(1 for r in got if float(r.split()[4]) > 0) == m["valid"]
    assert hashlib.sha256(("\n".join(got) + "\n").encode()).hexdigest() == m["sha256"]


def test_serial_mode_differs_from_the_batch_driver_on_the_golden_streams(tmp_path):
    """The two drivers really disagree on these streams (else the goldens above would pin nothing new)."""
    name = "serial_sweeps_30x40_fs5"
    w, h, fs, inl, build = SERIAL_CASES[name]
    x, y, t, p = build()
    ser = _serial_rows(tmp_path, name, w, h, fs, inl, x, y, t, p)
    bat = oracle_text(tmp_path, name + "_b", w, h, fs, inl, x, y, t, p).decode().splitlines()[1:1 + len(ser)]
    assert sum(a != b for a, b in zip(ser, bat)) > 20


# ---------------------------------------------------------------------------------------------------
# Plane-fit intermediates the reference's files do not carry: the inlier count computeGrads returned and the window
# computeLocalFlow chose, recorded per event through the same probe under the batch driver.
# ---------------------------------------------------------------------------------------------------
def _oracle_fit_rows(tmp_path, name, w, h, fs, inl, x, y, t, p):
    base = str(tmp_path / name)
    write_txt(base + ".txt", x, y, t, p)
    subprocess.run([CLI, str(w), str(h), str(fs), str(inl), base], check=True, capture_output=True,
                   env=dict(os.environ, FARMS_ORACLE_DIAG="1"))
    rows = [ln.split() for ln in open(base + "_FARMSOut_oracle_diag.txt")]  # valid best_window inliers det
    return ["%s %s" % (r[2], r[1]) for r in rows]


@pytest.mark.parametrize("name", sorted(TEXT_CASES))
def test_inlier_counts_and_windows_golden(name, tmp_path):
    w, h, fs, inl, build = TEXT_CASES[name]
    got = _oracle_fit_rows(tmp_path, name, w, h, fs, inl, *build())
    ref = open(os.path.join(GOLDEN, name + ".fit.ref.txt")).read().splitlines()
    assert len(ref) == META_FIT[name]["rows"]
    assert got == ref
    assert len({r.split()[1] for r in ref}) >= 3  # several different windows win on these streams


@pytest.mark.parametrize("name", sorted(FIT_SYNTH_CASES))
def test_inlier_counts_and_windows_golden_on_benchmark_scenes(name, tmp_path):
    from farms_synth import Synth
    cfg, n, start = FIT_SYNTH_CASES[name]
    s = Synth(cfg)
    x, y, t, p = s.first(n, start)
    got = _oracle_fit_rows(tmp_path, name, s.width, s.height, s.filtersize, 5, x, y, t, p)
    assert len(got) == META_FIT[name]["rows"]
    assert hashlib.sha256(("\n".join(got) + "\n").encode()).hexdigest() == META_FIT[name]["sha256"]
