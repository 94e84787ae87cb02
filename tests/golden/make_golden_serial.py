#!/usr/bin/env python
"""Golden vectors for the semantics of the reference's DEFAULT driver, vFlowManager::run (--SERIAL 1).

That driver computes a local and a pooled flow per event and writes neither (src/vFlow.cpp:727-765 are commented
out).  `make -C oracle refserial` builds the UNMODIFIED reference sources as a shared library with the call probe of
oracle/serial_probe.cpp in front of computeLocalFlow / computeTrueFlow; this script runs that binary and turns the
probe's log into rows of the reference's 11-column batch format
    x y t p globalR globalTheta Vx Vy localR localTheta scale              (ostream default == "%g")
applying exactly what run() does with the two results (:648, :666-667, :705-706).  Row k of a golden file is event
k + 1 of the stream: the first line of a recording only sets t0 (:531-558) and calls neither function.

  python tests/golden/make_golden_serial.py          (needs /root/reference for the build only)

Cases of up to 3000 rows keep the rows (<name>.ref.txt), all keep their SHA-256; golden_serial.json also records
how many lines run() consumed, which pins its `numEvents <= filesize / 18` rule (:511) for the CLI.

The same probe also sits in front of computeGrads(subsurf, cen, ..): under the BATCH driver (--SERIAL 0) its log
gives, per event, the inlier count the plane fit returned (:1352-1369) and which of the 9 candidate windows won
(:870-910) -- two intermediate results that no output file of the reference carries, but that the parity tests of
the GPU path compare bit for bit (farms_out.inliers / best_window).  They go to <name>.fit.ref.txt as
"<inliers> <window>" per event ("0 -1" when no window fits inside the sensor and computeGrads is never reached)."""
import hashlib
import json
import math
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np  # noqa: E402
from kat_streams import FIT_SYNTH_CASES, SERIAL_CASES, SERIAL_SYNTH_CASES, TEXT_CASES, write_txt  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "FARMS_Flow_serial")


def run_probe(w, h, fs, inl, x, y, t, p, d, name):
    """-> (rows of the 11-column format for events 1..K, size of the input file, numEvents asked for)"""
    base = os.path.join(d, name)
    write_txt(base + ".txt", x, y, t, p)
    log = base + ".probe"
    ask = len(x)  # more than the file can satisfy: run() caps it at filesize / 18
    subprocess.run([REF, "--width", str(w), "--height", str(h), "--filtersize", str(fs), "--inlierCheck", str(inl),
                    "--filename", base, "--SERIAL", "1", "--numEvents", str(ask)], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=dict(os.environ, FARMS_SERIAL_PROBE_OUT=log))
    rows, k, pending = [], 0, None
    t0 = int(t[0])

    def flush():
        if pending is not None:
            rows.append(pending)

    for ln in open(log):
        f = ln.split()
        if f[0] == "G":  # (plane-fit intermediates: see run_probe_fit)
            continue
        if f[0] == "L":  # computeLocalFlow() of the next event (:629)
            flush()
            k += 1
            vx, vy = float(f[1]), float(f[2])
            pol = max(int(p[k]), 0)  # (:575)
            trel = (int(t[k]) - t0) & 0xFFFFFFFF
            # an event without valid local flow (:648): zeros, Vx / Vy as computed
            pending = "%d %d %d %d %g %g %g %g %g %g %d" % (x[k], y[k], trel, pol, 0, 0, vx, vy, 0, 0, 0)
            local = (vx, vy)
        else:  # computeTrueFlow(x, y, time_, pol) of the same event (:703)
            xx, yy, tt, pp = int(f[1]), int(f[2]), int(f[3]), int(f[4])
            assert (xx, yy) == (int(x[k]), int(y[k])) and tt == (int(t[k]) - t0) & 0xFFFFFFFF, (k, ln)
            tvx, tvy, scale = float(f[5]), float(f[6]), int(f[7])
            vx, vy = local
            length = math.sqrt(vx * vx + vy * vy)            # :666
            theta = math.atan2(vy, vx)                       # :667
            true_length = math.sqrt(tvy * tvy + tvx * tvx)   # :705
            true_angle = math.atan2(tvy, tvx)                # :706
            pending = "%d %d %d %d %g %g %g %g %g %g %d" % (xx, yy, tt, pp, true_length, true_angle, vx, vy, length,
                                                            theta, scale)
    flush()
    return rows, os.path.getsize(base + ".txt"), ask


def run_probe_fit(w, h, fs, inl, x, y, t, p, d, name):
    """Batch driver behind the probe -> one "<inliers> <window>" row per event."""
    base = os.path.join(d, name + "_fit")
    write_txt(base + ".txt", x, y, t, p)
    log = base + ".probe"
    subprocess.run([REF, "--width", str(w), "--height", str(h), "--filtersize", str(fs), "--inlierCheck", str(inl),
                    "--filename", base, "--SERIAL", "0"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                   env=dict(os.environ, FARMS_SERIAL_PROBE_OUT=log))
    rows, grads = [], "0 -1"
    for ln in open(log):
        f = ln.split()
        if f[0] == "G":    # inside computeLocalFlow, before it returns
            grads = "%d %d" % (int(f[1]), int(f[2]))
        elif f[0] == "L":
            rows.append(grads)
            grads = "0 -1"
    assert len(rows) == len(x)
    return rows


def summary(rows, fsize, ask):
    raw = ("\n".join(rows) + "\n").encode()
    return {"rows": len(rows), "valid": sum(1 for r in rows if float(r.split()[4]) > 0), "input_bytes": fsize,
            "input_lines": ask, "num_events_asked": ask, "sha256": hashlib.sha256(raw).hexdigest()}, raw


def main():
    if not os.path.exists(REF):
        raise SystemExit("oracle/_ref/FARMS_Flow_serial is missing: run `make -C oracle refserial` where /root/reference exists")
    from farms_synth import Synth
    meta = {}
    with tempfile.TemporaryDirectory() as d:
        for name, (w, h, fs, inl, build) in SERIAL_CASES.items():
            x, y, t, p = build()
            m, raw = summary(*run_probe(w, h, fs, inl, x, y, t, p, d, name))
            if m["rows"] <= 3000:  # (larger ones are pinned by their hash alone)
                open(os.path.join(HERE, name + ".ref.txt"), "wb").write(raw)
            meta[name] = m
        for name, (cfg, n, start) in SERIAL_SYNTH_CASES.items():
            s = Synth(cfg)
            x, y, t, p = s.first(n, start)
            t_start = time.time()
            m, _ = summary(*run_probe(s.width, s.height, s.filtersize, 5, x, y, t, p, d, name))
            m["reference_wall_s"] = round(time.time() - t_start, 1)
            m["input_sha256"] = hashlib.sha256(np.stack([x.astype(np.int64), y.astype(np.int64), t.astype(np.int64),
                                                          p.astype(np.int64)], 1).tobytes()).hexdigest()
            meta[name] = m
        fit = {}
        for name, (w, h, fs, inl, build) in TEXT_CASES.items():
            rows = run_probe_fit(w, h, fs, inl, *build(), d, name)
            raw = ("\n".join(rows) + "\n").encode()
            open(os.path.join(HERE, name + ".fit.ref.txt"), "wb").write(raw)
            fit[name] = {"rows": len(rows), "fitted": sum(1 for r in rows if not r.endswith("-1")),
                         "sha256": hashlib.sha256(raw).hexdigest()}
        for name, (cfg, n, start) in FIT_SYNTH_CASES.items():
            s = Synth(cfg)
            x, y, t, p = s.first(n, start)
            rows = run_probe_fit(s.width, s.height, s.filtersize, 5, x, y, t, p, d, name)
            raw = ("\n".join(rows) + "\n").encode()
            fit[name] = {"rows": len(rows), "fitted": sum(1 for r in rows if not r.endswith("-1")),
                         "sha256": hashlib.sha256(raw).hexdigest()}
    for k, v in fit.items():
        print(k, v["rows"], v["fitted"], v["sha256"][:12])
    json.dump(fit, open(os.path.join(HERE, "golden_fit.json"), "w"), indent=1, sort_keys=True)
    for k, v in meta.items():
        # run() reads the first line, then numEvents + 1 more with numEvents capped at filesize / 18 (:511, :565)
        assert v["rows"] == min(min(v["num_events_asked"], v["input_bytes"] // 18) + 1, v["input_lines"] - 1), (k, v)
        print(k, v["rows"], v["valid"], v["sha256"][:12])
    json.dump(meta, open(os.path.join(HERE, "golden_serial.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
