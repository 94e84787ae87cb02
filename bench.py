#!/usr/bin/env python
"""bench.py -- FARMS event-flow throughput on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py [--gpus N] [--steps K] [--warmup W]           # our CUDA path through the C ABI
  python bench.py --impl reference [...]                        # the reference's own CPU path

A "step" is one pass of the hot path (ingest -> history index -> plane fit -> pooling) over one synthetic
event stream of the BASELINE config the metric is quoted on: 1280x720, filtersize 5, 200 M events
(configs[3]; deterministic generator tools/farms_synth.cpp).  `value` is measured with the stream already
resident in HBM; `e2e` is the same pass through farms_process_host with pinned HOST buffers, host<->device
copies inside the timed region.  N > 1: the stream is time-sliced, one slice of the same length per GPU
(weak scaling), each rank rebuilding the surface of active events at its slice start from an NCCL
all-gather of per-slice "last event per pixel" surfaces, plus a 499-us causal halo for pooling.
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
E2E_COLUMNS = ["t_rel", "global_r", "global_theta", "vx", "vy", "local_r", "local_theta", "scale", "valid"]


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--events", type=int, default=200_000_000, help="events per GPU per step")
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=450_000, help="events of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from farms_synth import Synth
    syn = Synth(args.config)
    W, H, FS = syn.width, syn.height, syn.filtersize
    workload = (f"configs[{args.config - 1}]: synthetic {'high-rate pan' if args.config >= 4 else 'scene'} {W}x{H}, "
                f"{args.events} events per GPU, filtersize {FS}, inlierCheck 5")

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
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "timer": "reference's own loop timer (src/vFlow.cpp:214-423)"},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind,
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # this rank's time slice (stream microseconds), with the causal halo in front (slicing.py)
    import slicing
    D = int(args.events / syn.rate * 1e6)
    plan = slicing.slice_plan(rank, world, D, HALO_US)
    tg0 = time.time()
    x, y, t, p = syn.time_range(plan.t_lo, plan.t_end, pinned=True)
    gen_s = time.time() - tg0
    n_all = len(x)
    # outputs of [0, n_halo) are discarded; [0, n_surf) is this rank's share of the SAE exchange
    n_halo, n_surf = slicing.split_counts(t.astype(np.int64) - 1000, plan)
    n_owned = n_all - n_halo
    # global t0 = first timestamp of the whole stream (reference src/vFlow.cpp:194)
    t0_t = torch.tensor([int(t[0]) if rank == 0 else 0], dtype=torch.int64, device=dev)
    if dist:
        dist.broadcast(t0_t, 0)
    t0 = int(t0_t.item())

    hx, hy = torch.from_numpy(x), torch.from_numpy(y)
    ht = torch.from_numpy(t.view(np.int64))
    dx, dy, dt = hx.to(dev), hy.to(dev), ht.to(dev)
    npx = W * H
    f = farms_b200.Farms(W, H, FS, 5, device=local_rank)
    surf_t = torch.zeros(npx, dtype=torch.int32, device=dev)
    surf_hit = torch.zeros(npx, dtype=torch.uint8, device=dev)
    all_t = torch.zeros(world * npx, dtype=torch.int32, device=dev) if dist else None
    all_hit = torch.zeros(world * npx, dtype=torch.uint8, device=dev) if dist else None
    tdt = {np.uint32: torch.int32, np.float64: torch.float64, np.uint8: torch.uint8}
    dev_out = {k: torch.empty(n_all, dtype=tdt[farms_b200.OUT_DTYPES[k]], device=dev) for k in E2E_COLUMNS}
    host_out = None

    def exchange_state():
        """SAE at this rank's halo start = fold of earlier ranks' 'last event per pixel' surfaces."""
        f.set_t0(t0)
        if not dist:
            return
        f.slice_surface(dx[:n_surf], dy[:n_surf], dt[:n_surf], t0, surf_t, surf_hit)
        dist.all_gather_into_tensor(all_t, surf_t)
        dist.all_gather_into_tensor(all_hit, surf_hit)
        for r in range(rank):
            f.state_fold(all_t[r * npx:(r + 1) * npx], all_hit[r * npx:(r + 1) * npx])

    # final NCCL gather of the owned events' outputs to rank 0 (float4 per event); slices differ by a few
    # events, so every rank sends n_max rows and rank 0 keeps the first sizes[r]
    if dist:
        sz = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sz, torch.tensor([n_owned], dtype=torch.int64, device=dev))
        sizes = [int(q.item()) for q in sz]
        n_max = max(sizes)
        packed = torch.zeros((n_max, 4), dtype=torch.float32, device=dev)
        gathered = [torch.empty((n_max, 4), dtype=torch.float32, device=dev) for _ in range(world)] if rank == 0 else None

    def gather_outputs(cols):
        if not dist:
            return
        f.pack4_f32(cols["global_r"][n_halo:], cols["global_theta"][n_halo:], cols["local_r"][n_halo:],
                    cols["local_theta"][n_halo:], packed)
        dist.gather(packed, gathered, dst=0)

    launches = [0]
    stage = {}

    sect = {"exchange_s": 0.0, "process_s": 0.0, "gather_s": 0.0}

    def step_device():
        f.reset()
        ta = time.perf_counter()
        exchange_state()
        torch.cuda.synchronize()
        tb = time.perf_counter()
        f.process_device(dx, dy, dt, columns=E2E_COLUMNS, out=dev_out)
        tc = time.perf_counter()
        sect["exchange_s"] += tb - ta
        sect["process_s"] += tc - tb
        tm = f.timings()
        launches[0] += tm["kernel_launches"] + (3 + rank if dist else 0)
        for k, v in tm.items():
            stage[k] = stage.get(k, 0) + v
        td = time.perf_counter()
        gather_outputs(dev_out)
        torch.cuda.synchronize()
        sect["gather_s"] += time.perf_counter() - td

    def step_host():
        f.reset()
        exchange_state()
        # every rank's results land in pinned host memory of this node: nothing left to gather
        f.process(x, y, t, columns=E2E_COLUMNS, out=host_out)

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
        # the library works on its own stream and synchronises it before returning, so wall clock between two
        # full synchronisations is the device-inclusive time; CUDA events on torch's stream bracket the same span
        dev_ms = e0.elapsed_time(e1)
        tmax = torch.tensor([max(wall, dev_ms * 1e-3)], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item())

    for _ in range(args.warmup):
        step_device()
    launches[0] = 0
    stage.clear()
    for k in sect:
        sect[k] = 0.0
    sampler = ClockSampler(local_rank)
    sampler.start()
    sec = timed(step_device, args.steps)
    clocks = sampler.stop()
    launches_per_run = launches[0]
    stage_avg = {k: v / args.steps for k, v in stage.items()}

    tot = torch.tensor([n_owned], dtype=torch.int64, device=dev)
    if dist:
        dist.all_reduce(tot)
    total_events = int(tot.item())
    value = total_events * args.steps / sec / 1e6

    e2e = None
    if not args.no_e2e:
        host_out = {}
        for k in E2E_COLUMNS:
            buf = torch.empty(n_all, dtype=tdt[farms_b200.OUT_DTYPES[k]]).pin_memory()
            a = buf.numpy()
            host_out[k] = a.view(np.uint32) if farms_b200.OUT_DTYPES[k] is np.uint32 else a
        for _ in range(args.warmup):
            step_host()
        sec_h = timed(step_host, args.steps)
        h2d = n_all * (2 + 2 + 8)
        d2h = n_all * sum(np.dtype(farms_b200.OUT_DTYPES[k]).itemsize for k in E2E_COLUMNS)
        e2e = {"value": total_events * args.steps / sec_h / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * sec_h / args.steps,
               "api": "farms_process_host (pinned host SoA in, pinned host columns out)"}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0

    # roofline of the dominant kernel (the stage with the most device time)
    peak, peak_src = measured_peak()
    nbatches = max(1, -(-n_all // (16 << 20)))
    if stage_avg.get("pool_ms", 0) >= stage_avg.get("fit_ms", 0):
        kname, kms, balg, nl = "k_pool_tile", stage_avg["pool_ms"], B_ALG_POOL, nbatches
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
    if not args.no_cpu_baseline:
        sample = syn.first(args.cpu_sample)
        r, s_sec, kind = reference_rate(sample, W, H, FS, "cpu")
        cpu_baseline = {"value": r / 1e6, "unit": UNIT, "cores": 1, "kind": kind, "seconds": s_sec,
                        "host_cores_available": os.cpu_count(),
                        "sample": f"first {args.cpu_sample} events of the workload stream (cold surface of active "
                                  "events); the reference is single-threaded"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "events_per_step_total": total_events, "l2": "inputs_larger_than_l2",
                       "slicing": f"time slices of {D} us per GPU + {HALO_US} us causal halo" if world > 1 else "none",
                       "generator_s": gen_s},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_run),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "stages_ms_per_step": {k: stage_avg[k] for k in ("total_ms", "ingest_ms", "index_ms", "fit_ms", "bin_ms",
                                                              "pool_ms") if k in stage_avg},
            "rank0_sections_ms_per_step": {k: 1e3 * v / args.steps for k, v in sect.items()},
            "valid_events_per_step": int(stage_avg.get("valid_events", 0)),
            "pool_candidates_per_step": int(stage_avg.get("pool_candidates", 0))}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
