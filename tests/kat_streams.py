"""Formula-generated known-answer event streams (no RNG library: a fixed LCG), shared by the golden-vector
generator (tests/golden/make_golden.py) and the tests."""
import numpy as np


def plane(width, height, ax=100, ay=30, t0=1000):
    """One event per pixel on the plane t = t0 + ax*x + ay*y, sorted by (t, x, y), polarity 1
    (SURVEY.md Appendix B uses 16x12, ax=100, ay=30)."""
    ev = sorted((t0 + ax * x + ay * y, x, y) for x in range(width) for y in range(height))
    a = np.array(ev, dtype=np.int64)
    return a[:, 1].copy(), a[:, 2].copy(), a[:, 0].copy(), np.ones(len(a), np.int64)


def _lcg(n, seed):
    out = np.empty(n, np.int64)
    s = seed & 0xFFFFFFFF
    for i in range(n):
        s = (1664525 * s + 1013904223) & 0xFFFFFFFF
        out[i] = s >> 16
    return out


def sweeps(width, height, slopes=((40, 11), (-25, 33), (17, -29)), jitter=3, gap=400, seed=7, drop=0.0):
    """Several plane sweeps over the sensor with different slopes, +-jitter us of timestamp noise, a fraction
    `drop` of the pixels skipped per sweep, polarity alternating per sweep.  Sorted by (t, x, y)."""
    n = width * height
    xs, ys, ts, ps = [], [], [], []
    base = 1000
    for k, (ax, ay) in enumerate(slopes):
        r = _lcg(2 * n, seed + 101 * k)
        off = min(0, ax * (width - 1)) + min(0, ay * (height - 1))
        i = 0
        tmax = 0
        for x in range(width):
            for y in range(height):
                j = int(r[2 * i] % (2 * jitter + 1)) - jitter if jitter else 0
                keep = (r[2 * i + 1] % 1000) >= drop * 1000
                i += 1
                if not keep:
                    continue
                t = base + ax * x + ay * y - off + jitter + j
                xs.append(x); ys.append(y); ts.append(t); ps.append(k % 2)
                tmax = max(tmax, t)
        base = tmax + gap
    a = np.array(sorted(zip(ts, xs, ys, ps)), dtype=np.int64)
    return a[:, 1].copy(), a[:, 2].copy(), a[:, 0].copy(), a[:, 3].copy()


def write_txt(path, x, y, t, p):
    np.savetxt(path, np.stack([np.asarray(x, np.int64), np.asarray(y, np.int64), np.asarray(t, np.int64),
                               np.asarray(p, np.int64)], 1), fmt="%d")


# name -> (width, height, filtersize, inlierCheck, builder)
TEXT_CASES = {
    # SURVEY.md Appendix B.  width > height on a heap-sized surface: the reference's own scale column depends on
    # heap garbage past the end of its vectors (see DESIGN.md "Oracle"), so only columns 1-10 are compared.
    "kat_plane_16x12": (16, 12, 5, 5, lambda: plane(16, 12)),
    "kat_plane_12x16": (12, 16, 5, 5, lambda: plane(12, 16)),
    "kat_sweeps_20x24_fs5": (20, 24, 5, 5, lambda: sweeps(20, 24)),
    "kat_sweeps_20x24_fs3_inl3": (20, 24, 3, 3, lambda: sweeps(20, 24, drop=0.2)),
    "kat_sweeps_18x30_fs7": (18, 30, 7, 5, lambda: sweeps(18, 30, jitter=1)),
}

# name -> (width, height, filtersize, inlierCheck, builder): only the SHA-256 of the reference's output is kept
HASH_CASES = {
    "sweeps_160x120_fs5": (160, 120, 5, 5, lambda: sweeps(160, 120, slopes=((12, 5), (-9, 14), (7, -6)), gap=200)),
    "sweeps_200x150_fs7": (200, 150, 7, 5, lambda: sweeps(200, 150, slopes=((9, 4), (-6, 10)), gap=300, drop=0.1)),
}

# synthetic benchmark scenes (tools/farms_synth): name -> (config, n events, stream start us)
SYNTH_CASES = {
    "synth_cfg1_320x320": (1, 20000, 0),
    "synth_cfg2_304x240": (2, 40000, 0),
    "synth_cfg3_346x260_fs7": (3, 40000, 0),
    "synth_cfg4_1280x720": (4, 150000, 0),
}

# long prefixes of the benchmark scenes (steady state: at 1280x720 the valid fraction only settles at ~54 % after
# ~1 M events); the reference needs minutes for each, the fast mode of the tier-2 oracle seconds.
LONG_SYNTH_CASES = {
    "synth_cfg2_304x240_1M": (2, 1_000_000, 0),
    "synth_cfg3_346x260_fs7_2M": (3, 2_000_000, 0),
    "synth_cfg4_1280x720_2M": (4, 2_000_000, 0),
}

# The reference's DEFAULT driver (vFlowManager::run, --SERIAL 1), observed through the call probe of
# oracle/serial_probe.cpp: name -> (width, height, filtersize, inlierCheck, builder).  width > height only on surfaces
# of >= 16,384 pixels (see TEXT_CASES about heap-sized ones: the reference reads past the end of its vectors there).
SERIAL_CASES = {
    "serial_sweeps_30x40_fs5": (30, 40, 5, 5, lambda: sweeps(30, 40, slopes=((25, 9), (-6, 9), (11, -7)), jitter=2, gap=120)),
    "serial_sweeps_20x24_fs3_inl3": (20, 24, 3, 3, lambda: sweeps(20, 24)),
    "serial_sweeps_18x30_fs7": (18, 30, 7, 5, lambda: sweeps(18, 30, jitter=1)),
    "serial_sweeps_200x100_fs5": (200, 100, 5, 5, lambda: sweeps(200, 100, slopes=((9, 4), (-6, 10)), gap=300, drop=0.1)),
}
# ... and on a synthetic benchmark scene: name -> (config, n events, stream start us); SHA-256 of the rows only
SERIAL_SYNTH_CASES = {
    "serial_synth_cfg2_304x240": (2, 60000, 0),
}
# plane-fit intermediates (inlier count, winning window) of the batch driver through the same probe; TEXT_CASES keep
# their rows, these keep a SHA-256: name -> (config, n events, stream start us)
FIT_SYNTH_CASES = {
    "fit_synth_cfg2_304x240": (2, 40000, 0),
    "fit_synth_cfg3_346x260_fs7": (3, 40000, 0),
}
