#!/usr/bin/env python
"""Regenerate tests/golden/* from the tier-1 oracle: the UNMODIFIED reference sources built against
oracle/shim (oracle/_ref/FARMS_Flow, `make -C oracle ref`; needs /root/reference at build time only).

  python tests/golden/make_golden.py

Text cases keep the reference's full 11-column output; hash cases keep its SHA-256 and a few counters."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import numpy as np  # noqa: E402
from kat_streams import HASH_CASES, LONG_SYNTH_CASES, SYNTH_CASES, TEXT_CASES, write_txt  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "FARMS_Flow")


def run_ref(w, h, fs, inl, x, y, t, p, d, name):
    base = os.path.join(d, name)
    write_txt(base + ".txt", x, y, t, p)
    subprocess.run([REF, "--width", str(w), "--height", str(h), "--filtersize", str(fs), "--inlierCheck", str(inl),
                    "--filename", base, "--SERIAL", "0"], check=True, capture_output=True)
    return open(base + "_FARMSOut_batch.txt", "rb").read()


def summary(raw):
    rows = [ln.split() for ln in raw.decode().splitlines()]
    valid = sum(1 for r in rows if float(r[4]) > 0)
    hist = {}
    for r in rows:
        hist[r[10]] = hist.get(r[10], 0) + 1
    return {"rows": len(rows), "valid": valid, "scale_hist": hist, "sha256": hashlib.sha256(raw).hexdigest(),
            "first_valid_rows": [" ".join(r) for r in rows if float(r[4]) > 0][:3]}


def main():
    if not os.path.exists(REF):
        raise SystemExit("oracle/_ref/FARMS_Flow is missing: run `make -C oracle ref` where /root/reference exists")
    from farms_synth import Synth
    meta = {}
    with tempfile.TemporaryDirectory() as d:
        for name, (w, h, fs, inl, build) in TEXT_CASES.items():
            raw = run_ref(w, h, fs, inl, *build(), d, name)
            open(os.path.join(HERE, name + ".ref.txt"), "wb").write(raw)
            meta[name] = summary(raw)
        for name, (w, h, fs, inl, build) in HASH_CASES.items():
            meta[name] = summary(run_ref(w, h, fs, inl, *build(), d, name))
        for name, (cfg, n, start) in list(SYNTH_CASES.items()) + list(LONG_SYNTH_CASES.items()):
            s = Synth(cfg)
            x, y, t, p = s.first(n, start)
            t_start = time.time()
            m = summary(run_ref(s.width, s.height, s.filtersize, 5, x, y, t, p, d, name))
            m["reference_wall_s"] = round(time.time() - t_start, 1)
            m["input_sha256"] = hashlib.sha256(np.stack([x.astype(np.int64), y.astype(np.int64), t.astype(np.int64),
                                                          p.astype(np.int64)], 1).tobytes()).hexdigest()
            meta[name] = m
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    for k, v in meta.items():
        print(k, v["rows"], v["valid"], v["sha256"][:12])


if __name__ == "__main__":
    main()
