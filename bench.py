#!/usr/bin/env python
"""bench.py -- FARMS event-flow throughput on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W]           # our CUDA path through the C ABI
  python bench.py --impl reference [...]                        # the reference's own CPU path

A "step" is one pass of the hot path (ingest -> history index -> plane fit -> pooling) over one synthetic
event stream of the BASELINE config the metric is quoted on: 1280x720, filtersize 5, 200 M events
(configs[3]; deterministic generator tools/farms_synth.cpp).  `value` is measured with the stream already
resident in HBM; `e2e` is the same pass through the C ABI with pinned HOST buffers, host<->device copies inside
the timed region.  N > 1: the stream is time-sliced, one slice per GPU, through farms_comm_process
(include/farms_b200.h, csrc/comm.cu): NCCL all-gather of per-slice "last event per pixel" surfaces, a 499-us causal
halo for pooling, and the contract columns of every rank gathered on rank 0 batch by batch (ncclSend/ncclRecv under
the next batch's kernels).  --scaling weak (default): --events per GPU; --scaling strong: --events in total
(configs[4] = 1 B events: `--scaling strong --events 1000000000`).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in ("aperture-robust-multiscale-optical-flow_b200", "tools"):
    sys.path.insert(0, os.path.join(ROOT, p))

METRIC = "farms_mevents_per_s"
UNIT = "Mevents/s"
HALO_US = 499                 # pooling admits |dt| < 500 us (reference src/vFlow.cpp:1002)
B_ALG_POOL = 54               # algorithmic HBM bytes per event of the pooling kernel
B_ALG_FIT = 35                # ... of the plane-fit kernel (K3)
# device-resident outputs of the timed pass: every column of the reference's 11-column row (src/vFlow.cpp:438)
DEV_COLUMNS = ["t_rel", "global_r", "global_theta", "vx", "vy", "local_r", "local_theta", "scale", "valid"]
# end-to-end pass: the README's 8-column contract (x y t p are echoes that stay with the caller) + the valid flag
E2E_COLUMNS = ["t_rel", "global_r", "global_theta", "local_r", "local_theta", "valid"]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [q.strip() for q in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_rate(sample_xytp, width, height, filtersize, tag):
    """Run oracle/_ref/FARMS_Flow (the unmodified reference sources, shim-built) on a bounded sample and
    return (events/s from its own loop timer, seconds).  Falls back to the C port when the binary is absent."""
    x, y, t, p = sample_xytp
    n = len(x)
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "FARMS_Flow")
    with tempfile.TemporaryDirectory() as d:
        base = os.path.join(d, f"sample_{tag}")
        arr = np.stack([x.astype(np.int64), y.astype(np.int64), t.astype(np.int64), p.astype(np.int64)], 1)
        np.savetxt(base + ".txt", arr, fmt="%d")
        if os.path.exists(ref_bin):
            out = subprocess.run([ref_bin, "--width", str(width), "--height", str(height), "--filtersize", str(filtersize),
                                  "--inlierCheck", "5", "--filename", base, "--SERIAL", "0"],
                                 capture_output=True, text=True, check=True).stdout
            m = re.search(r"Processing time\s*:\s*(\d+) usec", out)
            sec = int(m.group(1)) * 1e-6
            return n / sec, sec, "reference"
        cli = os.path.join(ROOT, "oracle", "farms_oracle_cli")
        out = subprocess.run([cli, str(width), str(height), str(filtersize), "5", base], capture_output=True, text=True,
                             check=True).stderr
        m = re.search(r"in ([\d.]+) s", out)
        sec = float(m.group(1))
        return n / sec, sec, "port"


def bind_to_gpu_numa_node(dev_index):
    """Run this rank (and so its pinned allocations, first-touch) on the CPUs of the GPU's NUMA node."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(dev_index).pci_bus_id
        dom = torch.cuda.get_device_properties(dev_index).pci_domain_id
        devid = torch.cuda.get_device_properties(dev_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return {"numa_node": node, "bound": False}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "bound": bool(allowed), "cpus": len(allowed)}
    except Exception as e:  # no sysfs entry (virtualised box): leave the affinity alone
        return {"numa_node": None, "bound": False, "why": str(e)[:80]}


def roofline_kernels(peak):
    """Per-kernel DRAM GB/s against the measured HBM peak, from the committed ncu launch list of this build
    (profiles/kernels.json, written by tools/summarize_profiles.py from `ncu --metrics gpu__time_duration.sum,
    dram__bytes_*`): cold-cache, serialised launches -- compare shares and fractions, not absolute times."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernels.json")) as fh:
            k = json.load(fh)
        return {"source": k.get("source"), "peak_gbs": peak,
                "kernels": [{"kernel": e["kernel"], "share_of_step": e["share"], "dram_gbs": e["gbs"],
                             "frac_of_hbm_peak": e["gbs"] / peak} for e in k["kernels"]]}
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events", type=int, default=200_000_000, help="events per GPU (weak) or in total (strong) per step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=450_000, help="events of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip exact-pooling / other-config / sliced-parity extras")
    ap.add_argument("--parity-events", type=int, default=2_000_000)
    ap.add_argument("--pool-variant", default="", help="A/B runs: tile | bits | tile1 | warp | tile16 ... (default: the library's choice)")
    ap.add_argument("--fit-chunk", type=int, default=0, help="A/B runs: events per plane-fit chunk (default: the library's choice)")
    ap.add_argument("--max-batch", type=int, default=0, help="A/B runs: events per internal batch (default: 32 Mi)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from farms_synth import Synth
    syn = Synth(args.config)
    W, H, FS = syn.width, syn.height, syn.filtersize
    per_gpu = args.events if args.scaling == "weak" else args.events // max(world, 1)
    cfg_name = f"configs[{args.config - 1}]" if not (args.scaling == "strong" and args.events >= 1_000_000_000) else "configs[4]"
    workload = (f"{cfg_name}: synthetic {'high-rate pan' if args.config >= 4 else 'scene'} {W}x{H}, "
                + (f"{args.events} events per GPU" if args.scaling == "weak" else f"{args.events} events in total")
                + f", filtersize {FS}, inlierCheck 5")

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = syn.first(args.cpu_sample)
        rates = []
        kind = "reference"
        for i in range(args.warmup + args.steps):
            r, sec, kind = reference_rate(sample, W, H, FS, f"ref{i}")
            if i >= args.warmup:
                rates.append((r, sec))
        total_sec = sum(s for _, s in rates)
        value = args.cpu_sample * len(rates) / total_sec / 1e6
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_sec / len(rates),
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "l2": "inputs_larger_than_l2"},  # the same object as the B200 arm's
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
                                 "timer": "reference's own loop timer (src/vFlow.cpp:214-423)",
                                 "steady_state_note": "2 M-event prefix of this workload (45 % valid): 7.3 k events/s "
                                                      "(BASELINE.md section 5); the cold prefix timed here is ~9x faster",
                                 "sample": f"first {args.cpu_sample} events of the workload stream per step "
                                           "(cold surface; the reference is single-threaded and cannot use more cores)"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (B200)
    import torch
    import farms_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def new_comm(farms):
        """farms_comm over the ranks of this job: rank 0 makes the NCCL id, torch.distributed only carries it."""
        if world == 1:
            return farms_b200.Comm(farms, 1, 0, b"")
        idt = torch.zeros(farms_b200.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(farms_b200.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return farms_b200.Comm(farms, world, rank, bytes(idt.cpu().numpy().tobytes()))

    import slicing
    tdt = {np.uint32: torch.int32, np.float64: torch.float64, np.uint8: torch.uint8}

    class Slice:
        """This rank's time slice of a stream of `span_us` microseconds cut `world` ways (+ causal halo)."""

        def __init__(self, synth, span_us, pinned):
            D = -(-span_us // world)
            self.D = D
            plan = slicing.slice_plan(rank, world, D, HALO_US)
            tg = time.time()
            self.x, self.y, self.t, self.p = synth.time_range(plan.t_lo, min(plan.t_end, span_us), pinned=pinned)
            self.gen_s = time.time() - tg
            self.n = len(self.x)
            self.n_halo, self.n_surf = slicing.split_counts(self.t.astype(np.int64) - 1000, plan)
            if rank == world - 1:
                self.n_surf = 0
            self.n_owned = self.n - self.n_halo
            t0_t = torch.tensor([int(self.t[0]) if rank == 0 else 0], dtype=torch.int64, device=dev)
            if dist:
                dist.broadcast(t0_t, 0)
            self.t0 = int(t0_t.item())

        def to_device(self):
            self.dx = torch.from_numpy(self.x).to(dev)
            self.dy = torch.from_numpy(self.y).to(dev)
            self.dt = torch.from_numpy(self.t.view(np.int64)).to(dev)

    def all_counts(n_owned):
        tot = torch.tensor([n_owned], dtype=torch.int64, device=dev)
        if dist:
            dist.all_reduce(tot)
        return int(tot.item())

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        wall = time.perf_counter() - w0
        # the library works on its own streams and synchronises them before returning, so wall clock between two
        # full synchronisations is the device-inclusive time; CUDA events on torch's stream bracket the same span
        dev_ms = e0.elapsed_time(e1)
        tmax = torch.tensor([max(wall, dev_ms * 1e-3)], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item())

    # ---------------------------------------------------------------- sliced parity (N > 1), before any timing
    sliced_parity = None
    if world > 1 and not args.no_extras:
        span = int(args.parity_events / syn.rate * 1e6)
        sl = Slice(syn, span, pinned=False)
        sl.to_device()
        fp = farms_b200.Farms(W, H, FS, 5, device=local_rank)
        cmp_ = new_comm(fp)
        total = all_counts(sl.n_owned)
        gathered = torch.zeros((total, 4), dtype=torch.float32, device=dev) if rank == 0 else torch.zeros(1, device=dev)
        counts = cmp_.process(sl.dx, sl.dy, sl.dt, sl.n_halo, sl.n_surf, sl.t0, out=None, gather_dst=gathered, device=True)
        barrier()
        if rank == 0:
            # the same events in one piece on this GPU alone
            x1, y1, t1, _ = syn.time_range(0, span)
            f1 = farms_b200.Farms(W, H, FS, 5, device=local_rank)
            one = f1.process(x1, y1, t1, columns=["global_r", "global_theta", "local_r", "local_theta", "valid"])
            g = gathered.cpu().numpy().astype(np.float64)
            ok_n = len(x1) == total == int(counts.sum())
            bad = np.zeros(min(len(x1), total), bool)
            maxrel = 0.0
            if ok_n:
                for j, k in enumerate(["global_r", "global_theta", "local_r", "local_theta"]):
                    a, b = g[:, j], one[k]
                    if k.endswith("_r"):
                        with np.errstate(invalid="ignore", divide="ignore"):
                            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
                        rel[b == 0] = np.abs(a[b == 0])
                        bad |= rel > 1e-4
                        maxrel = max(maxrel, float(rel.max()))
                    else:
                        d = np.abs(a - b) % (2 * np.pi)
                        bad |= np.minimum(d, 2 * np.pi - d) > 1e-3
            sliced_parity = {"events": int(total), "single_gpu_events": int(len(x1)), "mismatches": int(bad.sum()) if ok_n else -1,
                             "max_rel_r": maxrel, "valid": int(one["valid"].sum()), "slices": world,
                             "tolerance": "R 1e-4 relative, theta 1e-3 rad (float4 transport)", "transport": cmp_.transport()}
            f1.close()
        cmp_.close()
        fp.close()
        del gathered, sl
        barrier()

    # ---------------------------------------------------------------- the workload
    span_us = int(args.events / syn.rate * 1e6) * (world if args.scaling == "weak" else 1)
    sl = Slice(syn, span_us, pinned=True)
    sl.to_device()
    n_all, n_owned = sl.n, sl.n_owned
    total_events = all_counts(n_owned)
    f = farms_b200.Farms(W, H, FS, 5, device=local_rank, pool_variant=args.pool_variant or 0, fit_chunk=args.fit_chunk,
                         max_batch=args.max_batch)
    cm = new_comm(f)
    dev_out = {k: torch.empty(n_owned, dtype=tdt[farms_b200.OUT_DTYPES[k]], device=dev) for k in DEV_COLUMNS}
    gathered = torch.empty((total_events, 4), dtype=torch.float32, device=dev) if rank == 0 else torch.zeros(1, device=dev)
    host_out = None
    launches = [0]
    stage = {}

    def step_device():
        f.reset()
        cm.process(sl.dx, sl.dy, sl.dt, sl.n_halo, sl.n_surf, sl.t0, out=dev_out, gather_dst=gathered, device=True)
        tm = f.timings()
        if dist:
            for k, v in cm.phases().items():
                stage[k] = stage.get(k, 0) + v
        launches[0] += tm["kernel_launches"] + (2 + rank if dist else 0)
        for k, v in tm.items():
            stage[k] = stage.get(k, 0) + v

    def step_host():
        f.reset()
        # every rank's results land in pinned host memory of this node
        cm.process(sl.x, sl.y, sl.t, sl.n_halo, sl.n_surf, sl.t0, out=host_out, gather_dst=None, device=False)

    for _ in range(args.warmup):
        step_device()
    launches[0] = 0
    stage.clear()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sec = timed(step_device, args.steps)
    clocks = sampler.stop()
    launches_per_run = launches[0]
    stage_avg = {k: v / args.steps for k, v in stage.items()}
    value = total_events * args.steps / sec / 1e6

    e2e = None
    if not args.no_e2e:
        host_out = {}
        for k in E2E_COLUMNS:
            buf = torch.empty(n_owned, dtype=tdt[farms_b200.OUT_DTYPES[k]]).pin_memory()
            a = buf.numpy()
            host_out[k] = a.view(np.uint32) if farms_b200.OUT_DTYPES[k] is np.uint32 else a
        for _ in range(args.warmup):
            step_host()
        sec_h = timed(step_host, args.steps)
        h2d = n_all * (2 + 2 + 8)
        d2h = n_owned * sum(np.dtype(farms_b200.OUT_DTYPES[k]).itemsize for k in E2E_COLUMNS)
        e2e = {"value": total_events * args.steps / sec_h / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * sec_h / args.steps,
               "pcie_gbs_this_rank": (h2d + d2h) * args.steps / sec_h / 1e9, "numa": numa,
               "columns": E2E_COLUMNS,
               "api": ("farms_comm_process" if world > 1 else "farms_process_host") +
                      " (pinned host SoA in, pinned host columns out: the README's 8-column contract)"}

    # ---------------------------------------------------------------- extras on rank 0 (not part of `value`)
    extras = {}
    if rank == 0 and not args.no_extras and world == 1:
        nx = min(20_000_000, n_all)
        fx = farms_b200.Farms(W, H, FS, 5, device=local_rank, flags=farms_b200.FLAG_EXACT_POOLING)
        cols = {k: dev_out[k][:nx] for k in DEV_COLUMNS}
        fx.process_device(sl.dx[:nx], sl.dy[:nx], sl.dt[:nx], columns=DEV_COLUMNS, out=cols)
        torch.cuda.synchronize()
        ta = time.perf_counter()
        fx.reset()
        fx.process_device(sl.dx[:nx], sl.dy[:nx], sl.dt[:nx], columns=DEV_COLUMNS, out=cols)
        torch.cuda.synchronize()
        extras["exact_pooling"] = {"value": nx / (time.perf_counter() - ta) / 1e6, "unit": UNIT, "events": nx,
                                   "what": "FARMS_FLAG_EXACT_POOLING (FP64 sums for every event; the FARMS_Flow text "
                                           "mode's default), device-resident, one pass"}
        fx.close()
        other = []
        for cfg_i, n_i in ((1, 100_000), (2, 5_000_000), (3, 20_000_000)):
            si = Synth(cfg_i)
            xi, yi, ti, _ = si.first(n_i)
            dxi, dyi = torch.from_numpy(xi.copy()).to(dev), torch.from_numpy(yi.copy()).to(dev)
            dti = torch.from_numpy(ti.copy().view(np.int64)).to(dev)
            fi = farms_b200.Farms(si.width, si.height, si.filtersize, 5, device=local_rank)
            oi = {k: torch.empty(n_i, dtype=tdt[farms_b200.OUT_DTYPES[k]], device=dev) for k in DEV_COLUMNS}
            best = None
            for it in range(4):
                fi.reset()
                torch.cuda.synchronize()
                ta = time.perf_counter()
                fi.process_device(dxi, dyi, dti, columns=DEV_COLUMNS, out=oi)
                torch.cuda.synchronize()
                dtm = time.perf_counter() - ta
                if it:
                    best = dtm if best is None else min(best, dtm)
            other.append({"config": f"configs[{cfg_i - 1}]", "sensor": f"{si.width}x{si.height}", "filtersize": si.filtersize,
                          "events": n_i, "value": n_i / best / 1e6, "unit": UNIT, "ms": 1e3 * best,
                          "valid_events": int(fi.timings()["valid_events"])})
            fi.close()
        extras["other_configs"] = other

    if rank != 0:
        cm.close()
        if dist:
            dist.destroy_process_group()
        return 0

    # roofline of the dominant kernel (the stage with the most device time)
    peak, peak_src = measured_peak()
    nbatches = max(1, -(-n_all // (args.max_batch or (32 << 20))))
    if stage_avg.get("pool_ms", 0) >= stage_avg.get("fit_ms", 0):
        kname, kms, balg, nl = "k_pool_tile16", stage_avg["pool_ms"], B_ALG_POOL, nbatches
    else:
        kname, kms, balg, nl = "k_fit_gather", stage_avg["fit_ms"], B_ALG_FIT, nbatches
    achieved = balg * n_all / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    # DRAM traffic of that kernel from the committed `ncu --set full` capture (profiles/traffic.json), scaled
    # from the captured launch to this run's events per launch
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tr = json.load(fh)[kname]
        traffic = tr["dram_bytes_per_event"] * (n_all // nl)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_event": balg, "events_per_launch": n_all // nl,
                "avg_launch_ms": kms / nl, "traffic_source": "profiles/traffic.json (ncu dram__bytes_read+write per event x events per launch)",
                "note": "shared-memory/issue-bound gather kernel: HBM is <1% utilised by construction, see DESIGN.md"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        sample = syn.first(args.cpu_sample)
        r, s_sec, kind = reference_rate(sample, W, H, FS, "cpu")
        cpu_baseline = {"value": r / 1e6, "unit": UNIT, "cores": 1, "kind": kind, "seconds": s_sec,
                        "host_cores_available": os.cpu_count(),
                        "sample": f"first {args.cpu_sample} events of the workload stream (cold surface of active "
                                  "events); the reference is single-threaded"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64 plane fit and scale decisions; pooling sums f32 ring partials combined in f64 (exact f64 re-pool when undecided)",
            "data": "synthetic",
            "config": {"workload": workload, "l2": "inputs_larger_than_l2"},
            "run": {"events_per_step_total": total_events,
                    "slicing": f"time slices of {sl.D} us per GPU + {HALO_US} us causal halo (farms_comm_process, "
                               f"transport {cm.transport()})" if world > 1 else "none",
                    "generator_s": sl.gen_s,
                    "value_starts_from": "x/y/t resident in HBM (SURVEY 8(d) counts from pinned host memory: that is `e2e`)",
                    "delivered": "all 11-column outputs on the owning GPU (f64) + float4 contract columns of every "
                                 "rank gathered on rank 0"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_run),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "stages_ms_per_step": {k: stage_avg[k] for k in ("total_ms", "ingest_ms", "index_ms", "fit_ms", "bin_ms",
                                                              "pool_ms") if k in stage_avg},
            "comm_phases_ms_per_step_rank0": ({k: stage_avg[k] for k in ("surface_ms", "exchange_ms", "event_loop_ms", "drain_ms")
                                              if k in stage_avg} if world > 1 else None),
            "pool_paths_events_per_step": {k: int(stage_avg.get(k, 0)) for k in ("pool_events_first", "pool_events_second",
                                                                                 "pool_events_general")},
            "pool_kernels": int(f.timings()["pool_kernels"]),
            "valid_events_per_step": int(stage_avg.get("valid_events", 0)),
            "pool_candidates_per_step": int(stage_avg.get("pool_candidates", 0))}
    if sliced_parity is not None:
        line["sliced_parity"] = sliced_parity
    rk = roofline_kernels(peak)
    if rk:
        line["roofline_kernels"] = rk
    line.update(extras)
    print(json.dumps(line))
    cm.close()
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
