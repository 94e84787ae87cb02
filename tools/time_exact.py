"""Stage times with FARMS_FLAG_EXACT_POOLING (development aid)."""
import sys
sys.path.insert(0, 'aperture-robust-multiscale-optical-flow_b200'); sys.path.insert(0, 'tools')
import numpy as np, torch
import farms_b200
from farms_synth import Synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
syn = Synth(4)
x, y, t, p = syn.first(n)
dev = torch.device('cuda', 0)
dx, dy = torch.from_numpy(x.copy()).to(dev), torch.from_numpy(y.copy()).to(dev)
dt = torch.from_numpy(t.copy().view(np.int64)).to(dev)
for flags in (0, farms_b200.FLAG_EXACT_POOLING):
    f = farms_b200.Farms(syn.width, syn.height, syn.filtersize, 5, flags=flags)
    for _ in range(2):
        f.reset()
        f.process_device(dx, dy, dt, columns=["global_r", "global_theta", "scale"])
    tm = f.timings()
    print("flags", flags, {k: round(v, 2) for k, v in tm.items() if k.endswith("_ms")})
