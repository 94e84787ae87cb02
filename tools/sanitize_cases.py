#!/usr/bin/env python
"""Small streams, one per pooling-kernel instantiation, for compute-sanitizer (tools/sanitize.sh).

  python tools/sanitize_cases.py <case> [--check]

Each case prints the FARMS_POOLK_* kernels that ran and the events pooled per path (farms_timings), so the
sanitizer log shows which code was covered; --check also compares with the CPU oracle (slow under a sanitizer)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("aperture-robust-multiscale-optical-flow_b200", "tools", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np  # noqa: E402


def squeeze(t, f):
    return (t[0] + ((t - t[0]).astype(np.float64) / f).astype(np.uint64)).astype(np.uint64)


def build(case):
    from farms_synth import Synth
    from kat_streams import sweeps
    if case == "dense":      # k_pool_tile16<8,512,3,2> + flagged second pass <16,960,4,1> + k_pool_any
        s = Synth(2)
        x, y, t, p = s.first(150_000, 0)
        return s.width, s.height, s.filtersize, x, y, squeeze(t, 8.0), {}
    if case == "sparse":     # k_pool_tile16<8,416,4,2>
        s = Synth(1)
        x, y, t, p = s.first(25_000, 0)
        return s.width, s.height, s.filtersize, x, y, t, {}
    if case in ("bits", "tile1", "tile", "warp", "tile16", "tile16x4", "tile16c"):
        s = Synth(2)
        x, y, t, p = s.first(100_000, 0)
        return s.width, s.height, s.filtersize, x, y, squeeze(t, 4.0), {"pool_variant": case}
    if case == "tall":       # owner tiles without reachable rows (width-1 row bound)
        x, y, t, p = sweeps(20, 160, slopes=((9, 2), (-7, 3)), gap=150)
        return 20, 160, 5, x, y, t.astype(np.uint64), {}
    if case in ("aliased", "aliased_c"):    # width > height: logical rows >= H alias the next column
        x, y, t, p = sweeps(150, 40, slopes=((9, 2), (-7, 3)), gap=150)
        return 150, 40, 5, x, y, t.astype(np.uint64), {"pool_variant": "tile16c"} if case == "aliased_c" else {}
    if case == "serial":     # FARMS_FLAG_SERIAL_SEMANTICS
        import farms_b200
        s = Synth(1)
        x, y, t, p = s.first(40_000, 0)
        return s.width, s.height, s.filtersize, x, y, t, {"flags": farms_b200.FLAG_SERIAL_SEMANTICS}
    if case in ("long4", "long3", "long4_c"):  # steady-state density, default kernels, several internal batches
        s = Synth(3 if case == "long3" else 4)
        x, y, t, p = s.first(1_500_000, 0)
        return s.width, s.height, s.filtersize, x, y, t, dict({"max_batch": 400_000},
                                                              **({"pool_variant": "tile16c"} if case == "long4_c" else {}))
    if case == "exact":      # k_pool_any for every event
        import farms_b200
        s = Synth(3)
        x, y, t, p = s.first(60_000, 0)
        return s.width, s.height, s.filtersize, x, y, t, {"flags": farms_b200.FLAG_EXACT_POOLING}
    raise SystemExit(f"unknown case {case}")


def main():
    import farms_b200
    case = sys.argv[1]
    w, h, fs, x, y, t, kw = build(case)
    kw.setdefault("max_batch", 70_000)
    f = farms_b200.Farms(w, h, fs, 5, **kw)
    got = f.process(x, y, t)
    tm = f.timings()
    rep = {"case": case, "checked_build": bool(farms_b200.lib().farms_build_is_checked()), "events": len(x),
           "valid": int(got["valid"].sum()),
           "pool_kernels": tm["pool_kernels"], "pool_events": [tm["pool_events_first"], tm["pool_events_second"],
                                                               tm["pool_events_general"]]}
    if "--check" in sys.argv:
        from helpers import assert_parity, compare, run_oracle
        serial = bool(kw.get("flags", 0) & farms_b200.FLAG_SERIAL_SEMANTICS)
        r = compare(got, run_oracle(w, h, fs, 5, x, y, t, fast=not serial, serial=serial), case)
        assert_parity(r)
        rep["oracle"] = "parity ok"
    print("sanitize_case", json.dumps(rep))


if __name__ == "__main__":
    main()
