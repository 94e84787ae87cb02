#!/usr/bin/env python
"""farms_view.py -- text counterpart of the reference's showOpticalFlowOutputWithHistogram.m.

The MATLAB script loads `<name>_FARMSOut_batch.txt` and, per time window, draws quiver plots and two polar
histograms (bins of pi/50 over 0..2*pi): the local flow angle (column 10) and the aperture-corrected global
angle (column 6) of the events with flow (showOpticalFlowOutputWithHistogram.m:38-47, 121-160, 255-259, 348-355).
This prints the same two histograms (and the chosen-scale histogram) as text, for the whole file or per window.

  python tools/farms_view.py <name>_FARMSOut_batch.txt | <name>_FARMSOut_.bin  [--window-us 50000] [--bins 100]
"""
import argparse
import sys

import numpy as np


def load(path):
    if path.endswith(".bin"):
        raw = open(path, "rb").read()
        assert raw[:8] == b"FARMSOU1", "not a FARMSOU1 result file"
        n = int(np.frombuffer(raw, np.uint64, 1, 8)[0])
        off, cols = 16, {}
        for k, dt in (("x", np.uint16), ("y", np.uint16), ("t", np.uint32), ("p", np.uint8), ("scale", np.uint8),
                      ("global_r", np.float64), ("global_theta", np.float64), ("vx", np.float64), ("vy", np.float64),
                      ("local_r", np.float64), ("local_theta", np.float64)):
            cols[k] = np.frombuffer(raw, dt, n, off)
            off += n * np.dtype(dt).itemsize
        return cols
    a = np.loadtxt(path, ndmin=2)
    names = ["x", "y", "t", "p", "global_r", "global_theta", "vx", "vy", "local_r", "local_theta", "scale"]
    return {k: a[:, i] for i, k in enumerate(names)}


def bar(h, width=50):
    m = h.max() if h.size and h.max() > 0 else 1
    return ["#" * int(round(width * v / m)) for v in h]


def show(c, sel, bins, title):
    flow = sel & (c["local_r"] > 0)
    n = int(flow.sum())
    print(f"== {title}: {int(sel.sum())} events, {n} with flow")
    if not n:
        return
    edges = np.linspace(0.0, 2 * np.pi, bins + 1)
    hl, _ = np.histogram(np.mod(c["local_theta"][flow], 2 * np.pi), edges)
    hg, _ = np.histogram(np.mod(c["global_theta"][flow], 2 * np.pi), edges)
    bl, bg = bar(hl, 30), bar(hg, 30)
    print(f"{'angle(deg)':>10s} {'local':>8s} {'':30s} {'global':>8s}")
    for i in range(bins):
        if hl[i] or hg[i]:
            print(f"{np.degrees(edges[i]):10.1f} {hl[i]:8d} {bl[i]:30s} {hg[i]:8d} {bg[i]}")

    def spread(theta):  # circular standard deviation, degrees
        r = np.hypot(np.mean(np.cos(theta)), np.mean(np.sin(theta)))
        return float(np.degrees(np.sqrt(max(-2.0 * np.log(min(max(r, 1e-300), 1.0)), 0.0)))) + 0.0

    print(f"circular spread: local {spread(c['local_theta'][flow]):.1f} deg, global {spread(c['global_theta'][flow]):.1f} deg")
    hs = np.bincount((c["scale"][flow] // 5).astype(int), minlength=11)
    print("chosen scale (half-width 0,5,..,50):", " ".join(str(v) for v in hs))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--window-us", type=int, default=0, help="0 = whole file")
    ap.add_argument("--bins", type=int, default=100, help="the MATLAB script uses pi/50 => 100 bins")
    args = ap.parse_args()
    c = load(args.path)
    t = c["t"].astype(np.int64)
    if args.window_us <= 0:
        show(c, np.ones(len(t), bool), args.bins, "all events")
    else:
        for w0 in range(int(t.min()), int(t.max()) + 1, args.window_us):
            show(c, (t >= w0) & (t < w0 + args.window_us), args.bins, f"t in [{w0}, {w0 + args.window_us}) us")
    return 0


if __name__ == "__main__":
    sys.exit(main())
