"""Host-side logic that needs no GPU: C-ABI exports, parameter normalisation, CLI contract, generators,
slicing plan (with a world_size-2 gloo run), bench reference arm."""
import ctypes
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT, Oracle

PKG = os.path.join(ROOT, "aperture-robust-multiscale-optical-flow_b200")
LIB = os.path.join(PKG, "libfarms_b200.so")
CLI = os.path.join(PKG, "FARMS_Flow")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not (os.path.exists(LIB) and os.path.exists(CLI)):
        subprocess.check_call(["make", "-C", PKG, "all"])


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "farms_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(farms_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 15
    L = ctypes.CDLL(LIB)
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/farms_b200.h but not exported"
    import farms_b200
    assert sorted(farms_b200.EXPORTS) == declared
    assert L.farms_abi_version() == 2


def test_filtersize_normalisation_matches_reference():
    """src/vFlow.cpp:32-38: <5 -> 3, even -> -1, fRad = fs/2, planeSize = fs^2."""
    L = ctypes.CDLL(LIB)
    r, p = ctypes.c_int32(), ctypes.c_int32()
    expect = {-3: 3, 0: 3, 3: 3, 4: 3, 5: 5, 6: 5, 7: 7, 8: 7, 9: 9, 12: 11}
    for fs, norm in expect.items():
        assert L.farms_normalize_filtersize(fs, ctypes.byref(r), ctypes.byref(p)) == norm
        assert r.value == norm // 2 and p.value == norm * norm


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import farms_b200
    with pytest.raises(farms_b200.FarmsError) as e:
        farms_b200.Farms(320, 320, 5, 5)
    assert e.value.code == farms_b200.ERR_CUDA
    out = subprocess.run([CLI, "--width", "16", "--height", "16", "--filename", "/nonexistent"], capture_output=True, text=True)
    assert out.returncode == 1 and "B200" in out.stderr


def test_create_rejects_bad_arguments():
    import farms_b200
    for w, h in [(0, 10), (10, 0), (-1, 5), (70000, 10)]:
        with pytest.raises(farms_b200.FarmsError) as e:
            farms_b200.Farms(w, h, 5, 5)
        assert e.value.code == farms_b200.ERR_ARG


def test_cli_flag_contract():
    out = subprocess.run([CLI, "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--help", "--filename", "--height", "--width", "--filtersize", "--inlierCheck", "--numEvents",
                 "--numevents", "--NUMEVENTS", "--SERIAL", "--v"):   # reference src/main.cpp:35-47
        assert flag in out.stdout
    bad = subprocess.run([CLI, "--bogus", "1"], capture_output=True, text=True)
    assert bad.returncode == 1 and bad.stderr.startswith("error: ")   # src/main.cpp:175-179
    bad = subprocess.run([CLI, "--width", "abc"], capture_output=True, text=True)
    assert bad.returncode == 1 and bad.stderr.startswith("error: ")
    bad = subprocess.run([CLI, "--width"], capture_output=True, text=True)
    assert bad.returncode == 1 and "missing" in bad.stderr


def test_synthetic_streams_are_deterministic_and_sorted():
    from farms_synth import Synth
    for cfg in (1, 2, 3, 4):
        s = Synth(cfg)
        x, y, t, p = s.time_range(2000, 9000, nthreads=3)
        x2, y2, t2, p2 = s.time_range(2000, 9000, nthreads=1)
        assert np.array_equal(x, x2) and np.array_equal(y, y2) and np.array_equal(t, t2) and np.array_equal(p, p2)
        assert len(x) > 100 and np.all(np.diff(t.astype(np.int64)) >= 0)
        assert x.max() < s.width and y.max() < s.height and set(np.unique(p)) <= {0, 1}
        # a range is the concatenation of its parts
        xa, _, ta, _ = s.time_range(2000, 5555)
        xb, _, tb, _ = s.time_range(5555, 9000)
        assert np.array_equal(np.concatenate([xa, xb]), x) and np.array_equal(np.concatenate([ta, tb]), t)


def test_slice_plan_tiles_the_time_axis():
    import slicing
    D, world = 10_000, 4
    plans = [slicing.slice_plan(r, world, D) for r in range(world)]
    assert plans[0].t_lo == 0 and plans[0].t_begin == 0
    for a, b in zip(plans[:-1], plans[1:]):
        assert a.t_end == b.t_begin                    # owned ranges tile
        assert b.t_lo == b.t_begin - slicing.HALO_US   # causal halo in front
        assert a.surf_end == b.t_lo                    # surface-exchange ranges tile up to the next halo start


_GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools"), os.path.join(ROOT, "aperture-robust-multiscale-optical-flow_b200")]
import slicing
from farms_synth import Synth
from helpers import Oracle
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
s = Synth(1)
D = 15000
plan = slicing.slice_plan(rank, world, D)
x, y, t, p = s.time_range(plan.t_lo, plan.t_end)
ts = t.astype(np.int64) - 1000
n_halo, n_surf = slicing.split_counts(ts, plan)
t0 = torch.tensor([int(t[0]) if rank == 0 else 0]); dist.broadcast(t0, 0); t0 = int(t0.item())
lt, hit = slicing.last_event_surface(x[:n_surf], y[:n_surf], (t[:n_surf] - t0).astype(np.uint32), s.width, s.height)
all_t = [torch.zeros(len(lt), dtype=torch.int64) for _ in range(world)]
all_h = [torch.zeros(len(lt), dtype=torch.uint8) for _ in range(world)]
dist.all_gather(all_t, torch.from_numpy(lt.astype(np.int64))); dist.all_gather(all_h, torch.from_numpy(hit))
acc_t, acc_h = np.zeros(len(lt), np.uint32), np.zeros(len(lt), np.uint8)
for r in range(rank):
    slicing.fold(acc_t, acc_h, all_t[r].numpy().astype(np.uint32), all_h[r].numpy())
# checker: the sequential oracle's surface after every event before this rank's halo start
xa, ya, ta, pa = s.time_range(0, plan.t_lo) if plan.t_lo > 0 else (x[:0], y[:0], t[:0], p[:0])
o = Oracle(s.width, s.height, 5, 5)
if len(xa): o.process(xa, ya, ta, pa)
ref_t, ref_h = o.state()
assert np.array_equal(ref_h, acc_h), "hit mask differs"
assert np.array_equal(ref_t[ref_h.astype(bool)].astype(np.uint32), acc_t[acc_h.astype(bool)]), "last times differ"
# halo covers every event that can still matter for pooling of the first owned event
assert n_halo == int(np.count_nonzero(ts < plan.t_begin)) and (rank == 0 or ts[0] >= plan.t_begin - slicing.HALO_US)
tot = torch.tensor([len(x) - n_halo]); dist.all_reduce(tot)
if rank == 0:
    xf, _, _, _ = s.time_range(0, world * D)
    assert int(tot.item()) == len(xf), "owned ranges must partition the stream"
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_time_slicing_with_gloo_world_size_2(tmp_path):
    """N > 1 host logic on CPU: two ranks, gloo all_gather of per-slice surfaces, fold, checked against the
    sequential oracle's surface at the slice boundary."""
    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + _GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "GLOO_OK" in out.stdout


def test_bench_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "30000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mevents/s" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] == 1 and line["e2e"]["h2d_bytes_per_step"] == 0


def test_viewer_reads_the_reference_output_format(capsys):
    """tools/farms_view.py on a golden reference output: Appendix B's plane has one flow direction,
    theta = 1.27934 rad = 73.3 deg, for all 184 valid events."""
    import farms_view
    import sys
    argv = sys.argv
    sys.argv = ["farms_view.py", os.path.join(ROOT, "tests", "golden", "kat_plane_16x12.ref.txt"), "--bins", "36"]
    try:
        assert farms_view.main() == 0
    finally:
        sys.argv = argv
    out = capsys.readouterr().out
    assert "192 events, 184 with flow" in out
    assert "      70.0      184" in out and "circular spread: local 0.0 deg, global 0.0 deg" in out


@pytest.mark.parametrize("w,h,kw", [(20, 24, {}), (36, 20, dict(slopes=((9, 4), (-6, 10), (7, -5)), gap=200))])
def test_summed_area_table_pooling_prototype_is_exact(w, h, kw):
    """tools/sat_pooling_prototype.py (the algorithm DESIGN.md section 9 sizes as the next step, not the product):
    per-slab summed-area tables corrected by the slab's own births and deaths give the oracle's scale for every event,
    including on a width > height sensor where rows alias the next column."""
    from helpers import run_oracle
    from kat_streams import sweeps
    from sat_pooling_prototype import pool_sat
    x, y, t, p = sweeps(w, h, **kw)
    o = run_oracle(w, h, 5, 5, x, y, t, p)
    gr, gth, sc, st = pool_sat(w, h, x, y, o["t_rel"], o["local_r"], o["local_theta"], o["valid"])
    v = o["valid"].astype(bool)
    assert v.sum() > 1000 and st["slabs"] > 5
    assert np.array_equal(sc[v], o["scale"][v])
    assert np.all(np.abs(gr[v] - o["global_r"][v]) <= 1e-10 * o["global_r"][v])
    assert np.all(np.abs(np.angle(np.exp(1j * (gth[v] - o["global_theta"][v])))) <= 1e-10)


@pytest.mark.parametrize("header", ["farms_b200.h", "farms_textio.h"])
def test_public_headers_are_plain_c(header, tmp_path):
    """The drop-in boundary is a C ABI: the headers must compile as strict C99 (and as C++11) on their own."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "%s"\nint main(void) { return 0; }\n' % os.path.join(ROOT, "include", header))
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", str(src)], check=True)
    subprocess.run(["g++", "-std=c++11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", str(src)],
                   check=True)


def test_integration_snippet_is_the_tested_example():
    """INTEGRATION.md section 2 shows examples/vFlowB200.cpp verbatim (the file `make -C oracle dropin` builds and the
    GPU suite runs behind the reference's own main)."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```cpp\n(// src/vFlowB200\.cpp.*?)```", text, re.S).group(1)
    assert code == open(os.path.join(ROOT, "examples", "vFlowB200.cpp")).read()


@pytest.mark.skipif(not os.path.isdir("/root/reference/include"), reason="needs the reference's headers")
def test_dropin_example_compiles_against_the_reference_headers():
    subprocess.run(["g++", "-std=c++11", "-fsyntax-only", "-w", "-I/root/reference/src", "-I/root/reference/include",
                    "-I" + os.path.join(ROOT, "oracle", "shim"), "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "vFlowB200.cpp")], check=True)
