"""Join an `ncu --page source --csv` SASS listing with nvdisasm -g line info: executed warp instructions
and stall samples per CUDA source line (development aid).

  python tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel substring> [top N]
"""
import csv
import re
import subprocess
import sys
import tempfile
import os

rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
# instruction -> line, in order, for the kernel's section
lines = []
insec = False
cur = None
for ln in dis.splitlines():
    if ln.startswith("//---------------------"):
        insec = kern in ln and ".text." in ln
        continue
    if not insec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + os.environ.get("NCU_KERNEL", kern)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = hdr.index("Instructions Executed")
cs = hdr.index("# Samples")
body = rows[hi + 1:]
for j, r in enumerate(body):  # several kernels matched: keep the first one
    if r and r[0] == "Kernel Name":
        body = body[:j]
        break
if len(body) != len(lines):
    print(f"warning: {len(body)} SASS rows in the report vs {len(lines)} in the object", file=sys.stderr)
agg = {}
tot_i = tot_s = 0
for r, l in zip(body, lines):
    try:
        ni, ns = int(r[ci]), int(r[cs])
    except ValueError:
        continue
    a = agg.setdefault(l, [0, 0])
    a[0] += ni
    a[1] += ns
    tot_i += ni
    tot_s += ns
print(f"total warp instructions {tot_i}, samples {tot_s}")
srcfile = {}
for (l, (ni, ns)) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if l:
        path = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", l[0])
        if os.path.exists(path):
            srcfile.setdefault(path, open(path).read().splitlines())
            text = srcfile[path][l[1] - 1].strip()[:90]
    print(f"{100*ni/tot_i:5.1f}% inst {100*ns/max(tot_s,1):5.1f}% smp  {l}  {text}")
# per phase (ranges of source lines given as extra args "name:lo-hi")
for spec in sys.argv[5:]:
    name, rng = spec.split(":")
    lo, hi = map(int, rng.split("-"))
    ni = sum(v[0] for l, v in agg.items() if l and l[0] == "pooling.cu" and lo <= l[1] <= hi)
    ns = sum(v[1] for l, v in agg.items() if l and l[0] == "pooling.cu" and lo <= l[1] <= hi)
    print(f"phase {name:12s} {100*ni/tot_i:5.1f}% inst {100*ns/max(tot_s,1):5.1f}% samples")
