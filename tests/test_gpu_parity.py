"""GPU parity: the CUDA path through the C ABI against the CPU oracle on identical seeded inputs."""
import json

import numpy as np
import pytest

from helpers import assert_parity, compare, run_oracle, synth_stream

pytestmark = pytest.mark.gpu

CASES = [
    # (config, events, stream start us, filtersize override)
    (1, 30000, 0, None),
    (2, 30000, 0, None),
    (3, 30000, 0, None),
    (4, 120000, 0, None),
    (1, 20000, 100000, 3),
    (2, 20000, 50000, 7),
]


@pytest.mark.parametrize("config,n,start,fs", CASES)
def test_parity_synthetic(config, n, start, fs):
    import farms_b200
    s, x, y, t, p = synth_stream(config, n, start)
    fs = fs or s.filtersize
    ref = run_oracle(s.width, s.height, fs, 5, x, y, t, p)
    f = farms_b200.Farms(s.width, s.height, fs, 5)
    got = f.process(x, y, t)
    rep = compare(got, ref, f"cfg{config} n={n} fs={fs}")
    print(json.dumps(rep))
    assert rep["valid_ref"] > 0
    assert_parity(rep)
    tm = f.timings()
    assert tm["valid_events"] == rep["valid_ref"]
    assert tm["kernel_launches"] > 0
