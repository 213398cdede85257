"""Benchmark of the ISS + sieve hot path (BASELINE.json: time series / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: the throughput sweep of BASELINE.json (configs[4], SURVEY.md row
C5): i.i.d. N(0,1) float64 series of 3 dimensions and length 1,024, words
``of_weight(4, dim=3)`` in EXTENDED mode (445 iterated sums), sieves
NPI(q=(.5,1)) + PPV + MAX + MIN + END -> 2,225 features per series.  The 4 M
series of the sweep do not fit one GPU together with their 71 GB of features,
so every GPU holds a shard of ``--series-per-gpu`` series (default 524,288 =
4 M / 8: at N=8 the job is exactly the sweep; weak scaling below).

One step = one pass of ``Fruit.transform`` over the resident shard (one fused
CUDA launch).  ``value`` counts series of all ranks per second of the slowest
rank, inputs resident in HBM.  ``e2e`` is the same pipeline through the public
API on HOST buffers (pinned), host->device and device->host copies inside the
timed region, on ALL ranks at the same time (slowest rank counts).  At N > 1
the assembled feature matrix is verified on every rank (``gather_check``).
``configs`` (N = 1) carries fit / transform times, routes and parity numbers of
the other BASELINE.json configurations (C1-C4).  ``--impl reference`` times the
reference's CPU implementation of the path on the host cores: the unmodified
numba package (``baseline/_ref``, see baseline/install_ref.sh) when it imports,
else the oracle port (OpenMP), on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

T_LEN, N_DIMS, N_NODES, N_FEATS = 1024, 3, 445, 2225
FLOP_PER_SERIES = 2 * N_NODES * T_LEN            # one FMA per node and step
BYTES_PER_SERIES = 8 * T_LEN * N_DIMS + 8 * N_FEATS
NOMINAL_FP64_TFLOPS = 37.2                       # 148 SM x 64 DFMA/clk x 1.965 GHz
METRIC = "time series/sec (features)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--series-per-gpu", type=int, default=524288)
    ap.add_argument("--e2e-series", type=int, default=131072,
                    help="series per rank and end-to-end step (3.2 GB in + 2.3 GB out of pinned "
                         "host memory per rank; the same at every N)")
    ap.add_argument("--ref-seconds", type=float, default=200.0,
                    help="--impl reference: wall-clock target of the whole run (import + JIT, "
                         "calibration, warm-up and timed steps)")
    ap.add_argument("--ref-kind", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the C1-C4 block (fit / transform / parity of the other configurations)")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--no-multicast", action="store_true",
                    help="N > 1: copy-engine pushes instead of NVSwitch multicast stores")
    ap.add_argument("--nccl-gather", action="store_true",
                    help="assemble the features with NCCL instead of peer-memory pushes")
    return ap.parse_args()


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(names, r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def host_cores() -> int:
    return len(os.sched_getaffinity(0))


# ---------------------------------------------------------------------------
# CPU arms: the real reference (numba) and the oracle port (C + OpenMP)

def _c5_inputs(n: int, seed: int = 100):
    return np.random.default_rng(seed).standard_normal((n, N_DIMS, T_LEN))


class PortArm:
    """oracle/fruits_oracle.c: the reference's kernels restated in C, OpenMP
    over series (test infrastructure; timed here as the CPU baseline only)."""
    kind = "port"

    def __init__(self):
        import specs
        from oracle import pipeline as orc
        from oracle.build import load_oracle
        lib = load_oracle()
        # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
        lib.fo_set_num_threads(host_cores())
        self.cores = int(lib.fo_num_threads())
        self.fruit = orc.OracleFruit(specs.SPECS["C5_sweep"])
        np.random.seed(0)
        self.fruit.fit(specs.make_input("C5_sweep", 64))
        self.what = "oracle port of the reference's numba kernels (C, OpenMP over series)"

    def transform(self, X):
        return self.fruit.transform(X)


class ReferenceArm:
    """The unmodified reference package (irkri/fruits 1.0.0, numba) from
    ``$FRUITS_REF`` or ``baseline/_ref`` -- never /root/reference, which does
    not exist on the GPU box."""
    kind = "reference"

    def __init__(self):
        import importlib
        if not hasattr(np, "NINF"):
            np.NINF = -np.inf            # fruits/sieving/segment.py:72 needs it on numpy >= 2
        path = os.environ.get("FRUITS_REF") or os.path.join(ROOT, "baseline", "_ref")
        if not os.path.isdir(os.path.join(path, "fruits")):
            raise ImportError(f"no reference package under {path}")
        os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(ROOT, "baseline", "_numba_cache"))
        os.environ["NUMBA_NUM_THREADS"] = str(host_cores())
        # the repository's own `fruits` alias must not shadow the reference
        saved = {k: sys.modules.pop(k) for k in list(sys.modules)
                 if k == "fruits" or k.startswith("fruits.")}
        sys.path.insert(0, path)
        try:
            ref = importlib.import_module("fruits")
            assert os.path.realpath(ref.__file__).startswith(os.path.realpath(path)), ref.__file__
        finally:
            sys.path.remove(path)
        self.modules = {k: sys.modules.pop(k) for k in list(sys.modules)
                        if k == "fruits" or k.startswith("fruits.")}
        sys.modules.update(saved)
        import numba
        import specs
        numba.set_num_threads(min(host_cores(), numba.config.NUMBA_NUM_THREADS))
        self.cores = int(numba.get_num_threads())
        self.fruit = specs.build_fruit(ref, specs.SPECS["C5_sweep"])
        np.random.seed(0)
        self.fruit.fit(specs.make_input("C5_sweep", 64))
        self.what = (f"unmodified reference package (numba {numba.__version__}, "
                     f"{numba.config.THREADING_LAYER} threading layer) from {os.path.relpath(path, ROOT)}")

    def transform(self, X):
        return self.fruit.transform(X)


def make_cpu_arm(kind: str):
    if kind in ("auto", "reference"):
        try:
            return ReferenceArm()
        except Exception as exc:          # noqa: BLE001  (missing package, numba, JIT failure)
            if kind == "reference":
                raise
            print(f"# reference package unavailable ({type(exc).__name__}: {exc}); "
                  f"timing the oracle port", file=sys.stderr)
    return PortArm()


def time_cpu(arm, n: int, steps: int = 1, warmup: int = 0):
    X = _c5_inputs(n)
    for _ in range(warmup):
        arm.transform(X)
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.transform(X)
    dt = time.perf_counter() - t0
    return n * steps / dt, dt / steps


def calibrate(arm, seconds: float, calls: int, lo: int, hi: int):
    """Series per call so that ``calls`` calls take about ``seconds``: two
    calibration calls of different size separate the per-call overhead (numba
    launches one parallel region per word and sieve) from the per-series cost."""
    # (sizes large enough that every iterated sum [n, 1024] has left the caches:
    # the reference streams one such array per word and sieve)
    n1, n2 = max(128, 4 * arm.cores), max(512, 16 * arm.cores)
    arm.transform(_c5_inputs(n1, 98))               # JIT / thread pool warm-up
    _, t1 = time_cpu(arm, n1)
    _, t2 = time_cpu(arm, n2)
    per_series = max((t2 - t1) / (n2 - n1), 1e-6)
    fixed = max(t1 - per_series * n1, 0.0)
    n = int((seconds / max(calls, 1) - fixed) / per_series)
    return int(min(max(n, lo), hi)), {"per_call_overhead_s": fixed, "per_series_s": per_series,
                                      "calibration": [[n1, t1], [n2, t2]]}


def cpu_baseline_block():
    """cpu_baseline of the default run (N = 1, rank 0): the real reference when it
    imports, the port beside it; each on a bounded sample, two sizes to show that
    the cost is linear in the number of series."""
    out = {}
    for kind in ("reference", "port"):
        try:
            arm = ReferenceArm() if kind == "reference" else PortArm()
        except Exception as exc:          # noqa: BLE001
            out[kind] = {"unavailable": f"{type(exc).__name__}: {exc}"}
            continue
        # SURVEY.md 8(d) / BASELINE.md: C5 at N = 2,048 for the reference; the port is
        # ~3x faster, so it gets a larger sample
        n = 2048 if kind == "reference" else max(2048, 64 * arm.cores)
        arm.transform(_c5_inputs(max(64, 2 * arm.cores), 98))      # JIT / thread pool warm-up
        rate, step_s = time_cpu(arm, n)
        rate_half, _ = time_cpu(arm, n // 2)
        out[kind] = {"value": rate, "unit": "series/s", "cores": arm.cores, "kind": kind,
                     "sample": f"{n} series x {N_DIMS} x {T_LEN}, one transform call "
                               f"({step_s:.1f} s), scaled linearly in the number of series",
                     "half_sample_value": rate_half, "what": arm.what}
    best = out["reference"] if "value" in out.get("reference", {}) else out["port"]
    block = {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")}
    block["detail"] = out
    return block


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_start = time.perf_counter()
    arm = make_cpu_arm(args.ref_kind)            # (numba: ~1-2 min of JIT on a fresh box)
    calls = args.steps + args.warmup
    # per-series cost of the numba reference grows with the batch (every iterated sum
    # [n, 1024] is streamed through memory once per word and sieve), so its sample is
    # capped near the size BASELINE.md quotes (2,048); the port gets >= 64 series per core
    lo, hi = (max(2048, 64 * arm.cores), 16384) if arm.kind == "port" else (256, 1024)
    budget = max(args.ref_seconds - (time.perf_counter() - t_start), 30.0)
    n, cal = calibrate(arm, budget, calls, lo, hi)
    if arm.kind == "reference":
        # a fixed ladder of sample sizes (1,024 unless the budget forces less): the
        # reference's rate depends on the batch size (337 / 204 / 152 series/s at 1,024 /
        # 2,048 / 4,096 series on 16 cores), so runs on different boxes stay comparable
        n = max(s_ for s_ in (256, 512, 1024) if s_ <= max(n, 256))
    rate, step_s = time_cpu(arm, n, args.steps, args.warmup)
    sample = (f"{n} series x {N_DIMS} x {T_LEN} per step, {args.steps} step(s) after "
              f"{args.warmup} warm-up step(s), scaled linearly in the number of series")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate,
        "unit": "series/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": rate, "unit": "series/s", "cores": arm.cores, "kind": arm.kind,
                         "sample": sample, "what": arm.what, **cal},
        "e2e": {"value": rate, "unit": "series/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {
        "workload": ("C5 throughput sweep shard: series x 3 dims x length 1024, words "
                     "of_weight(4, dim=3) EXTENDED (445 iterated sums), sieves NPI(q=(.5,1)) + "
                     "PPV + MAX + MIN + END -> 2225 features"),
        "series_per_gpu": args.series_per_gpu, "n_dims": N_DIMS, "length": T_LEN,
        "n_features": N_FEATS,
        "l2": "inputs (12.9 GB per GPU) and outputs (9.3 GB) are far larger than the 126 MB L2",
    }


# ---------------------------------------------------------------------------
# the other BASELINE.json configurations (N = 1): times, routes, parity

# flop per series = c * nodes * T, c = 2 (unweighted Reals, Arctic) or 4
# (exponentially weighted Reals); slices 0-1 only (the CosWISS slices have no agreed count)
CONFIG_FLOPS = {"C1_readme": 7200, "C2_full": 428032, "C3_full": 6311936, "C4_twi": 4579328}


def config_block(fruits, peak_tflops: float):
    import torch

    import specs
    from helpers import parity_report

    gold = os.path.join(ROOT, "tests", "golden")
    block = {}
    for name in ("C1_readme", "C2_full", "C3_full", "C4_twi"):
        Xh = specs.make_input(name)
        X = torch.from_numpy(Xh).cuda()
        fruit = specs.build_fruit(fruits, specs.SPECS[name])
        np.random.seed(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fruit.fit(X)
        torch.cuda.synchronize()
        fit_s = time.perf_counter() - t0
        out = fruit.transform_device(X)                  # module load / JIT on first use
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 5
        ev[0].record()
        for _ in range(reps):
            fruit.transform_device(X, out=out)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        slices, col = [], 0
        for slc in fruit:
            k = slc.nfeatures()
            sub = out[:, col:col + k]
            ev[0].record()
            for _ in range(3):
                slc._transform_device(X, None, None, out, col, sanitize=True)
            ev[1].record()
            torch.cuda.synchronize()
            slices.append({"features": k, "ms": ev[0].elapsed_time(ev[1]) / 3,
                           "route": getattr(slc, "_last_launch", ("composed",))[0]})
            col += k
            del sub
        entry = {"series": int(X.shape[0]), "dims": int(X.shape[1]), "length": int(X.shape[2]),
                 "features": int(out.shape[1]), "fit_s": fit_s, "transform_ms": ms,
                 "series_per_s": X.shape[0] / (ms * 1e-3), "slices": slices}
        flops = CONFIG_FLOPS[name]
        ms01 = sum(s["ms"] for s in slices[:2])
        entry["roofline_slices_0_1"] = {
            "flop_per_series": flops, "achieved_tflops": flops * X.shape[0] / (ms01 * 1e-3) / 1e12,
            "frac": flops * X.shape[0] / (ms01 * 1e-3) / 1e12 / peak_tflops, "ms": ms01}
        # parity against vectors frozen from the real reference (tests/golden)
        res = out.cpu().numpy()
        if name == "C1_readme":
            g = np.load(os.path.join(gold, "pipeline_C1_readme.npz"))
            entry["parity"] = {"against": "reference golden, all rows",
                               "bit_identical": bool(np.array_equal(res, g["features"]))}
        elif name == "C4_twi":
            # its sieves need no fitting, so the first rows equal the 8-series golden
            g = np.load(os.path.join(gold, "pipeline_C4_twi.npz"))
            rep = parity_report(res[:8], g["features"])
            rep["arctic_slice_bit_identical"] = bool(np.array_equal(res[:8, 1533:],
                                                                    g["features"][:, 1533:]))
            entry["parity"] = {"against": "reference golden, first 8 rows", **rep}
        else:
            path = os.path.join(gold, f"full_{name}.npz")
            if os.path.exists(path):
                g = np.load(path)
                rep = parity_report(res[g["rows"]], g["features"])
                entry["parity"] = {"against": f"reference at full size, {len(g['rows'])} rows "
                                              f"(tests/golden/full_{name}.npz)", **rep}
        block[name] = entry
        del X, out
        torch.cuda.empty_cache()
    # the preparateur kernels either side of the path (csrc/prep_more.cu): HBM-bound
    # streaming kernels, each alone on 65,536 x 3 x 1,024 (1.6 GB in >> L2), algorithmic
    # bytes (input once + prepared copy once) over the measured copy bandwidth
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import prep_bandwidth
    head, rows = prep_bandwidth.measure(65536)
    block["preparateurs"] = {**head, "bound": "hbm", "kernels": [
        {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()} for r in rows]}
    torch.cuda.empty_cache()
    return block


# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import fruits_b200 as fruits
    import specs
    from fruits_b200 import _backend as be

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    be.lib()

    S = args.series_per_gpu
    spec = specs.SPECS["C5_sweep"]
    fruit = specs.build_fruit(fruits, spec)
    # fit on the host-generated parity subsample: one series is drawn
    fitX = specs.make_input("C5_sweep", 4096)
    np.random.seed(0)
    t0 = time.perf_counter()
    fruit.fit(fitX)
    torch.cuda.synchronize()
    fit_s = time.perf_counter() - t0

    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    X = torch.empty((S, N_DIMS, T_LEN), dtype=torch.float64, device=dev)
    chunk = 65536
    for i in range(0, S, chunk):
        X[i:i + chunk] = torch.randn((min(chunk, S - i), N_DIMS, T_LEN), dtype=torch.float64,
                                     device=dev, generator=gen)
    gather = world > 1 and not args.no_gather
    n_chunks = 8 if gather else 1
    from fruits_b200.parallel import transform_sharded
    # N > 1: every rank ends up with the assembled [N*S, F] feature matrix
    # (rank-major rows); the all-gather of row chunk c overlaps the kernel of c+1
    out = peer = None
    collective = "none"
    if gather and not args.nccl_gather:
        try:
            from fruits_b200.parallel import PeerGather
            peer = PeerGather(S, N_FEATS, multicast=not args.no_multicast)
            out = peer.out
            if peer.fused:
                n_chunks = 1
                collective = ("fused into the feature kernel: its epilogue stores with multimem.st "
                              "through the NVSwitch multicast mapping (NVLS) of the symmetric "
                              "[N*S, F] matrices, so every feature lands in the matrix of every "
                              "rank as it is written; one device-side barrier per step, no copy "
                              "or collective kernel")
            else:
                collective = ("every finished row chunk is pushed into the peers' feature "
                              "matrices (symmetric NVLink peer memory, copy engines, 8 chunks "
                              "overlapped with the kernels), one device-side barrier per step; "
                              "every rank holds the assembled [N*S, F] matrix")
        except Exception as exc:                      # no symmetric memory on this box
            if rank == 0:
                print(f"# PeerGather unavailable ({type(exc).__name__}: {exc}); NCCL all-gather",
                      file=sys.stderr)
            peer = None
    if out is None:
        out = torch.empty(((world if gather else 1) * S, N_FEATS), dtype=torch.float64,
                          device=dev)
        if gather:
            collective = ("nccl all_gather_into_tensor of the [S, F] feature blocks in 8 row "
                          "chunks on a side stream, overlapped with the kernels; every rank "
                          "holds the assembled [N*S, F] matrix")

    rows = S // n_chunks

    def step():
        if gather:
            transform_sharded(fruit, X, N_FEATS, chunks=n_chunks,
                              out=peer if peer is not None else out)
        else:
            fruit.transform_device(X, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches_per_call = fruit.get_slice(0)._last_launch[1]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * S * args.steps / (ms_total * 1e-3)

    # ---- N > 1: every rank verifies the assembled matrix ----
    gather_check = None
    if gather:
        # a hash (wrapping int64 sum of the bit patterns) of every rank block as this
        # rank sees it; rank r's own block must also equal a local recompute into
        # ordinary device memory, element for element
        blocks = out.view(world, S * N_FEATS).view(torch.int64)
        mine = blocks.sum(dim=1)                                       # [world]
        local_ref = torch.empty((S, N_FEATS), dtype=torch.float64, device=dev)
        fruit.transform_device(X, out=local_ref)
        own_equal = bool(torch.equal(out[rank * S:(rank + 1) * S], local_ref))
        own_hash = local_ref.view(torch.int64).sum().reshape(1)
        del local_ref
        seen = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(seen, mine)
        owner = [torch.empty_like(own_hash) for _ in range(world)]
        dist.all_gather(owner, own_hash)
        owner = torch.cat(owner)
        ok = own_equal and all(bool(torch.equal(s, owner)) for s in seen)
        flag = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_check = "ok" if int(flag.item()) == 1 else "MISMATCH"

    # ---- kernel-only timing of the dominant kernel + fp64 roof (rank 0) ----
    roofline = cpu = configs = None
    peak = None
    if rank == 0:
        kout = out[:S] if not gather else out[rank * S:rank * S + rows]
        kX = X if not gather else X[:rows]
        ks = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        ks[0].record()
        reps = 3
        for _ in range(reps):
            fruit.transform_device(kX, out=kout)
        ks[1].record()
        torch.cuda.synchronize()
        k_s = ks[0].elapsed_time(ks[1]) * 1e-3 / reps
        n_launch = kX.shape[0]
        kname, klaunches, kern = fruit.get_slice(0)._last_launch
        if kern is not None:
            em = kern.em
            kdesc = (f"{kname}: plan-specialised kernel, "
                     f"{len([p for p in em.p.parts if p.owned])} trie parts compiled separately "
                     f"and linked into one kernel, {32 * em.ppc * em.gpc} threads per CTA "
                     f"({em.gpc} groups of 32 series x {em.ppc} parts), tile {em.tt} steps")
        else:
            kdesc = f"{kname}<Reals, unweighted, PolP> (generic trie interpreter)"
        achieved = FLOP_PER_SERIES * n_launch / k_s / 1e12
        # measured fp64 FMA peak (same clocks / power state as the run)
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        grid, iters = sm_count * 8, 4096
        pbuf = torch.empty(grid * 256, dtype=torch.float64, device=dev)
        be.check(be.lib().fb_fp64_peak(pbuf.data_ptr(), grid, iters, be.stream_ptr()))
        torch.cuda.synchronize()
        ps = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ps[0].record()
        for _ in range(3):
            be.check(be.lib().fb_fp64_peak(pbuf.data_ptr(), grid, iters, be.stream_ptr()))
        ps[1].record()
        torch.cuda.synchronize()
        peak = 3 * grid * 256 * iters * 64 * 2 / (ps[0].elapsed_time(ps[1]) * 1e-3) / 1e12
        peaks, how = measured_peaks()
        hbm = BYTES_PER_SERIES * n_launch / k_s / 1e9
        # DRAM traffic of the kernel from the committed ncu capture, scaled to
        # this launch (traffic is linear in the number of series)
        traffic = tsrc = None
        for tname in ("r02_traffic.json", "r01_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if kern is not None and os.path.exists(tpath):
                with open(tpath) as f:
                    tr = json.load(f)
                traffic = ((tr["dram_bytes_read"] + tr["dram_bytes_write"])
                           / tr["series_per_launch"] * n_launch)
                tsrc = tname
                break
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic,
            "traffic_note": ("dram__bytes_read+write of one ncu --set full capture "
                             f"(profiles/{tsrc}), scaled to series_per_launch; "
                             f"algorithmic bytes per launch = {BYTES_PER_SERIES * n_launch}"),
            "kernel": kdesc, "launches_per_slice": klaunches,
            "kernel_ms": k_s * 1e3, "series_per_launch": n_launch,
            "flop_per_series": FLOP_PER_SERIES,
            "peak_source": "DFMA microbenchmark fb_fp64_peak measured in this run",
            "nominal_peak": NOMINAL_FP64_TFLOPS, "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS,
            "hbm": {"achieved": hbm, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": hbm / peaks["hbm_gbs"], "peak_source": f"MEASURED_PEAKS.json ({how})",
                    "bytes_per_series": BYTES_PER_SERIES},
        }

    # ---- end to end through the public API on pinned host buffers, ALL ranks at once ----
    # (the pinned buffers are first touched by a thread bound to the CPUs next to
    # this rank's GPU, so that the DMA does not cross the socket interconnect)
    affinity = os.sched_getaffinity(0)
    numa = "unbound"
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        near = os.sched_getaffinity(0)
        numa = f"{len(near)} of {len(affinity)} CPUs"
    except Exception as exc:              # noqa: BLE001  (no NVML / not permitted: unbound)
        numa = f"unbound ({type(exc).__name__})"
    E = min(args.e2e_series, S)
    while True:
        hx = hf = None
        try:
            hx = torch.empty((E, N_DIMS, T_LEN), dtype=torch.float64, pin_memory=True)
            hf = torch.empty((E, N_FEATS), dtype=torch.float64, pin_memory=True)
            got = 1
        except RuntimeError:            # not enough pinnable host memory
            hx = hf = None
            got = 0
        flag = torch.tensor([got], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            break
        if E <= 8192:
            raise RuntimeError("cannot pin the end-to-end host buffers")
        E //= 2                         # every rank retries with the same smaller batch
    hx.copy_(X[:E])
    hx_np, hf_np = hx.numpy(), hf.numpy()
    for _ in range(2):
        fruit.transform(hx_np, out=hf_np)
    e_steps = max(3, args.steps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        fruit.transform(hx_np, out=hf_np)
    torch.cuda.synchronize()
    e_s = time.perf_counter() - t0
    te = torch.tensor([e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_max = float(te.item())
    barrier()
    e2e = {"value": world * E * e_steps / e_max, "unit": "series/s",
           "h2d_bytes_per_step": int(hx.numel() * 8) * world,
           "d2h_bytes_per_step": int(hf.numel() * 8) * world,
           "series_per_rank_and_step": E, "steps": e_steps, "ms_per_step": e_max / e_steps * 1e3,
           "note": ("Fruit.transform(pinned numpy in, pinned numpy out) on every rank at the same "
                    "time, its own buffers per rank; barrier before, slowest rank counts; byte "
                    "counts are totals over all ranks")}
    if rank == 0 and world == 1:
        # the same call with ordinary (pageable) numpy arrays, as a caller of the
        # reference passes them: pinned staging ring + copy threads inside transform
        Ep = min(E, 65536)
        px = np.array(hx_np[:Ep])
        pf = np.empty((Ep, N_FEATS))
        fruit.transform(px, out=pf)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            fruit.transform(px, out=pf)
        torch.cuda.synchronize()
        e2e["pageable"] = {"value": Ep / ((time.perf_counter() - t0) / 2),
                           "unit": "series/s", "series_per_step": Ep,
                           "note": "plain numpy arrays in and out (host memcpy bound)"}
        del px, pf
    del hx, hf, hx_np, hf_np
    e2e["host_affinity"] = numa
    os.sched_setaffinity(0, affinity)

    if rank == 0 and world == 1:
        if not args.no_configs:
            del X, out
            torch.cuda.empty_cache()
            configs = config_block(fruits, peak)
        if not args.no_cpu_baseline:
            cpu = cpu_baseline_block()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "series/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": args.steps * n_chunks * launches_per_call,
            "roofline": roofline, "cpu_baseline": cpu, "fit_seconds": fit_s,
            "collective": collective,
        }
        if gather_check is not None:
            line["gather_check"] = gather_check
        if configs is not None:
            line["configs"] = configs
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
