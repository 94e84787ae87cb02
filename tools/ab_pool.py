"""A/B timing of the pooling implementations on the bench workload (development aid, not a bench line)."""
import os
import subprocess
import sys

n = sys.argv[1] if len(sys.argv) > 1 else "20000000"
for impl in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("tile", "tile1", "bits")):
    env = dict(os.environ, FARMS_POOL_IMPL=impl)
    out = subprocess.run([sys.executable, "bench.py", "--events", n, "--steps", "2", "--warmup", "1", "--no-e2e",
                          "--no-cpu-baseline"], env=env, capture_output=True, text=True)
    print(impl, out.stdout.strip()[-900:], out.stderr.strip()[-600:])
