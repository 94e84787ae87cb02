#!/usr/bin/env python
"""CPU prototype of the pooling algorithm DESIGN.md section 9 sizes as the next step (not part of the product).

The pooling kernel of this repository scans, for every event, the ~1400 staged flow events around it.  The
alternative: per time slab keep a summed-area table (SAT) of the flow events that are alive at the slab's start, and
correct a window sum taken from it by the few events that were born or died INSIDE the slab before the query:

    sum(window, event i) =   SAT_S(window)
                           - sum of base entries in the window that died by i   (superseded at their pixel by an
                                                                                  event of the slab with index <= i,
                                                                                  or 500 us old at t_i)
                           + sum of events j of the slab, i0 <= j <= i, in the window that are still the latest
                             event of their pixel at i and have flow

which is exact (the three sets partition the contributors of src/vFlow.cpp:996-1010).  The reference's flat-index
aliasing (rows bounded by width-1, :1000) is handled by building the table over LOGICAL cells (i, j) -> flat index
i*H + j, so an aliased pixel simply appears at two logical positions.

`pool_sat` implements this in numpy for small streams and is checked against the oracle by
tests/test_host_logic.py; `python tools/sat_pooling_prototype.py` prints, for a steady-state prefix of a benchmark
scene, how many explicit candidates a query has under this scheme compared with the contributors it pools."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALES = np.arange(0, 51, 5)
LIFE = 500.0


def logical_positions(g, w, h):
    """All logical cells (i, j), 0 <= i, j <= w-1, whose flat index i*h + j is g (src/vFlow.cpp:1000-1002)."""
    out = []
    i, j = g // h, g % h
    while i >= 0 and j <= w - 1:
        out.append((i, j))
        i, j = i - 1, j + h
    return out


def pool_sat(w, h, x, y, t_rel, lr, lth, valid, slab_us=100.0):
    """Batch-driver pooling (sorted timestamps) through per-slab summed-area tables.
    Returns (global_r, global_theta, scale, stats)."""
    n = len(x)
    npx = w * h
    t = np.asarray(t_rel, np.float64)
    assert np.all(np.diff(t) >= 0), "sorted timestamps"
    flat = np.asarray(x, np.int64) * h + np.asarray(y, np.int64)
    lcx, lcy = lr * np.cos(lth), lr * np.sin(lth)
    vals = np.stack([lr, lcx, lcy, np.ones(n)], 1) * (np.asarray(valid) != 0)[:, None]  # zero row = no flow
    # next event at the same pixel (index), n if none
    nxt = np.full(n, n, np.int64)
    seen = {}
    for i in range(n - 1, -1, -1):
        nxt[i] = seen.get(int(flat[i]), n)
        seen[int(flat[i])] = i
    latest = np.full(npx, -1, np.int64)  # latest event per pixel among those before the slab
    out_r, out_th, out_s = np.zeros(n), np.zeros(n), np.zeros(n, np.int32)
    stats = {"queries": 0, "explicit": 0, "contributors": 0, "slabs": 0}
    i0 = 0
    while i0 < n:
        t0 = t[i0]
        i1 = int(np.searchsorted(t, t0 + slab_us, side="left"))
        i1 = max(i1, i0 + 1)
        stats["slabs"] += 1
        # ---- base: the latest event of every pixel before the slab, if it has flow and can still be alive ----
        base = latest[(latest >= 0)]
        base = base[(vals[base, 3] > 0) & (t0 - t[base] < LIFE)]
        img = np.zeros((w + 1, w + 1, 4))
        d_i, d_j, d_sup, d_exp, d_val = [], [], [], [], []
        for j in base:
            sup = nxt[j] if nxt[j] < i1 else n            # superseded inside the slab?
            dies = sup < n or t[j] + LIFE <= t[i1 - 1]     # ... or 500 us old before the slab ends
            for (a, b) in logical_positions(int(flat[j]), w, h):
                img[a + 1, b + 1] += vals[j]
                if dies:
                    d_i.append(a); d_j.append(b); d_sup.append(sup); d_exp.append(t[j] + LIFE); d_val.append(vals[j])
        sat = img.cumsum(0).cumsum(1)
        d_i, d_j, d_sup, d_exp = (np.asarray(v) for v in (d_i, d_j, d_sup, d_exp))
        d_val = np.asarray(d_val).reshape(-1, 4)
        # ---- births: the slab's own flow events at their logical positions ----
        b_i, b_j, b_idx, b_end, b_val = [], [], [], [], []
        for j in range(i0, i1):
            if vals[j, 3] > 0:
                for (a, b) in logical_positions(int(flat[j]), w, h):
                    b_i.append(a); b_j.append(b); b_idx.append(j); b_end.append(nxt[j]); b_val.append(vals[j])
        b_i, b_j, b_idx, b_end = (np.asarray(v) for v in (b_i, b_j, b_idx, b_end))
        b_val = np.asarray(b_val).reshape(-1, 4)
        # ---- queries ----
        for i in range(i0, i1):
            if vals[i, 3] == 0:
                continue
            xi, yi, ti = int(x[i]), int(y[i]), t[i]
            dead = (d_sup <= i) | (d_exp <= ti) if len(d_i) else np.zeros(0, bool)
            born = (b_idx <= i) & (b_end > i) if len(b_i) else np.zeros(0, bool)
            best, bestv, prev_cnt = 0.0, None, 0
            for s in SCALES:
                a0, a1 = max(0, xi - s), min(xi + s, w - 1)
                c0, c1 = max(0, yi - s), min(yi + s, w - 1)
                if a0 > a1 or c0 > c1:
                    continue
                tot = sat[a1 + 1, c1 + 1] - sat[a0, c1 + 1] - sat[a1 + 1, c0] + sat[a0, c0]
                if len(d_i):
                    m = dead & (d_i >= a0) & (d_i <= a1) & (d_j >= c0) & (d_j <= c1)
                    tot = tot - d_val[m].sum(0)
                if len(b_i):
                    m = born & (b_i >= a0) & (b_i <= a1) & (b_j >= c0) & (b_j <= c1)
                    tot = tot + b_val[m].sum(0)
                cnt = int(round(tot[3]))
                # windows are nested: the same count as the scale before means the same contributors, whose mean the
                # reference reproduces bit for bit and then rejects with its strict '>' (:1054) -- a table difference
                # can be off in the last digit, so that case is decided by the count
                if cnt > 0 and cnt != prev_cnt and tot[0] / cnt > best:
                    best, bestv = tot[0] / cnt, (tot[1] / cnt, tot[2] / cnt, int(s))
                prev_cnt = cnt
                if s == SCALES[-1]:
                    stats["contributors"] += cnt
            if bestv is None:
                bestv = (lcx[i], lcy[i], 0)
            out_r[i] = np.hypot(bestv[0], bestv[1])
            out_th[i] = np.arctan2(bestv[1], bestv[0])
            out_s[i] = bestv[2]
            stats["queries"] += 1
            stats["explicit"] += len(d_i) + len(b_i)
        np.maximum.at(latest, flat[i0:i1], np.arange(i0, i1))
        i0 = i1
    return out_r, out_th, out_s, stats


def window_statistics(config=4, n=1_500_000, sample=400, slab_us=100.0, seed=1):
    """On a steady-state prefix of a benchmark scene: contributors a query pools against the explicit candidates
    (births + deaths of its slab inside its widest window) the table scheme would test."""
    for p in ("aperture-robust-multiscale-optical-flow_b200", "tools", "tests"):
        sys.path.insert(0, os.path.join(ROOT, p))
    from farms_synth import Synth
    from helpers import run_oracle
    s = Synth(config)
    x, y, t, p = s.first(n, 0)
    o = run_oracle(s.width, s.height, s.filtersize, 5, x, y, t, p, fast=True)
    valid = o["valid"].astype(bool)
    tr = o["t_rel"].astype(np.float64)
    xs, ys = x.astype(np.int64), y.astype(np.int64)
    flat = xs * s.height + ys
    nxt = np.full(n, n, np.int64)
    order = np.argsort(flat, kind="stable")
    same = flat[order][1:] == flat[order][:-1]
    nxt[order[:-1][same]] = order[1:][same]
    rng = np.random.default_rng(seed)
    q = rng.choice(np.nonzero(valid & (np.arange(n) > n // 2))[0], sample, replace=False)
    rows = []
    for i in q:
        t0 = np.floor(tr[i] / slab_us) * slab_us
        i0 = int(np.searchsorted(tr, t0, side="left"))
        i1 = int(np.searchsorted(tr, t0 + slab_us, side="left"))
        lo = int(np.searchsorted(tr, t0 - LIFE, side="right"))
        j = np.arange(lo, i1)
        inwin = (np.abs(xs[j] - xs[i]) <= 50) & (np.abs(ys[j] - ys[i]) <= 50)
        # what the event pools today: latest flow event of each pixel of the window, younger than 500 us
        alive = valid[j] & inwin & (j <= i) & (nxt[j] > i) & (tr[i] - tr[j] < LIFE)
        # table scheme: base entries of the window that die inside the slab + the slab's own flow events there
        base = valid[j] & inwin & (j < i0) & (nxt[j] >= i0)
        deaths = base & ((nxt[j] < i1) | (tr[j] + LIFE <= tr[i1 - 1]))
        births = valid[j] & inwin & (j >= i0)
        rows.append((int(alive.sum()), int(deaths.sum() + births.sum()), int(base.sum())))
    a = np.asarray(rows, float)
    return {"scene": f"configs[{config - 1}] {s.width}x{s.height}", "events": n, "queries_sampled": sample,
            "slab_us": slab_us, "contributors_per_query": a[:, 0].mean(), "explicit_candidates_per_query": a[:, 1].mean(),
            "base_entries_in_window": a[:, 2].mean()}


if __name__ == "__main__":
    import json
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    print(json.dumps(window_statistics(cfg)))
