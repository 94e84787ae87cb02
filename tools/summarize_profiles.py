"""Turn the ncu captures under gpurun_out/ into the text summaries committed under profiles/ (development aid).

  python tools/summarize_profiles.py launches <launches.csv> <out.txt> "<command>"
  python tools/summarize_profiles.py kernel <report.ncu-rep> <out.txt> "<title>" [kernel index]
"""
import collections
import csv
import json
import re
import subprocess
import sys

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(path, out, command):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    per = {}
    for r in rows[1:]:
        d = per.setdefault(r[ii], {"k": r[ki]})
        v = float(r[vi].replace(",", ""))
        if "time" in r[mi]:
            d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r[ui]]
        else:
            d[r[mi]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[ui]]
    agg = collections.OrderedDict()
    for i in sorted(per, key=int):
        d = per[i]
        k = re.sub(r"\(.*", "", d["k"]).replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["ms"]
        a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({command})\n")
        f.write("# totals over the whole process; per-launch times are cold-cache and serialised: compare shares.\n")
        f.write(f"# DRAM GB/s = (dram__bytes_read + dram__bytes_write) / time; frac = of the measured HBM peak {PEAK} GB/s\n")
        f.write(f"{'kernel':58s} {'launches':>8s} {'total_ms':>10s} {'share':>7s} {'dram_GB':>9s} {'GB/s':>8s} {'frac':>6s}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            gbs = a[2] / max(a[1], 1e-9) / 1e6
            f.write(f"{k[:58]:58s} {a[0]:8d} {a[1]:10.3f} {100*a[1]/tot:6.1f}% {a[2]/1e9:9.3f} {gbs:8.1f} {gbs/PEAK:6.3f}\n")
        f.write(f"{'total':58s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}\n")
    # machine-readable copy for bench.py's `roofline_kernels` (kernels above 0.1 % of the step)
    kj = {"source": f"{out} ({command})",
          "kernels": [{"kernel": k, "launches": a[0], "ms": a[1], "share": a[1] / tot, "dram_bytes": a[2],
                       "gbs": a[2] / max(a[1], 1e-9) / 1e6}
                      for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]) if a[1] / tot > 1e-3]}
    import os
    with open(os.path.join(os.path.dirname(out) or ".", "kernels.json"), "w") as f:
        json.dump(kj, f, indent=1)


def kernel(rep, out, title, index=0):
    src = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + index]
    d = dict(zip(hdr, zip(units, vals)))
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none, {title}\n")
        f.write(f"{'Kernel Name':86s}{d['Kernel Name'][1]}\n")
        for m in METRICS:
            if m in d:
                f.write(f"{m:86s}{d[m][1]:>20s} {d[m][0]}\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(*sys.argv[2:5])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 0)
