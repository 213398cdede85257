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
timed region.  ``--impl reference`` times the CPU implementation of the path
(the oracle port of the reference's numba kernels, all host threads) on a
bounded sample of the same workload.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--series-per-gpu", type=int, default=524288)
    ap.add_argument("--e2e-series", type=int, default=0,
                    help="series per end-to-end step (default: 262144 at N=1 -- 6.4 GB in + 4.7 GB "
                         "out of pinned host memory -- and 65536 per rank at N>1)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


# ---------------------------------------------------------------------------
def cpu_reference(seconds: float, steps: int = 1, warmup: int = 0):
    """Time the CPU implementation (oracle port, OpenMP over series) of the C5
    transform on a bounded sample; returns (series/s, cores, sample text)."""
    import specs
    from oracle import pipeline as orc
    from oracle.build import load_oracle

    orc_lib = load_oracle()
    # all host cores, also under torchrun (which exports OMP_NUM_THREADS=1)
    orc_lib.fo_set_num_threads(len(os.sched_getaffinity(0)))
    cores = int(orc_lib.fo_num_threads())
    spec = specs.SPECS["C5_sweep"]
    fitX = specs.make_input("C5_sweep", 64)
    of = orc.OracleFruit(spec)
    np.random.seed(0)
    of.fit(fitX)
    # calibrate on a small sample, then size the timed sample for ~`seconds`
    n0 = max(cores, 16)
    X0 = np.random.default_rng(99).standard_normal((n0, N_DIMS, T_LEN))
    t0 = time.perf_counter()
    of.transform(X0)
    rate0 = n0 / (time.perf_counter() - t0)
    n = int(max(n0, min(16384, rate0 * seconds / max(steps + warmup, 1))))
    X = np.random.default_rng(100).standard_normal((n, N_DIMS, T_LEN))
    for _ in range(warmup):
        of.transform(X)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        of.transform(X)
        times.append(time.perf_counter() - t0)
    dt = sum(times)
    return n * steps / dt, cores, (f"{n} series x {N_DIMS} x {T_LEN} per step, {steps} step(s), "
                                   f"scaled linearly in the number of series"), dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, cores, sample, step_s = cpu_reference(args.cpu_seconds, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "time series/sec (features)", "value": rate,
        "unit": "series/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, sample_only=True),
        "cpu_baseline": {"value": rate, "unit": "series/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": rate, "unit": "series/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, sample_only=False):
    cfg = {
        "workload": ("C5 throughput sweep shard: series x 3 dims x length 1024, words "
                     "of_weight(4, dim=3) EXTENDED (445 iterated sums), sieves NPI(q=(.5,1)) + "
                     "PPV + MAX + MIN + END -> 2225 features"),
        "series_per_gpu": args.series_per_gpu, "n_dims": N_DIMS, "length": T_LEN,
        "n_features": N_FEATS,
        "l2": "inputs (12.9 GB per GPU) and outputs (9.3 GB) are far larger than the 126 MB L2",
    }
    if sample_only:
        cfg["note"] = "CPU arm: bounded sample of the same workload (see cpu_baseline.sample)"
    return cfg


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
    rows = S // n_chunks
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
                collective = ("fused into the feature kernel: its stores go through the NVSwitch "
                              "multicast mapping (NVLS) of the symmetric [N*S, F] matrices, so "
                              "every feature lands in the matrix of every rank as it is written; "
                              "one device-side barrier per step, no copy or collective kernel")
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

    def compute(x, o):
        fruit.transform_device(x, out=o)

    def step():
        if gather:
            transform_sharded(compute, X, N_FEATS, chunks=n_chunks,
                              out=peer if peer is not None else out)
        else:
            compute(X, out)

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

    # ---- kernel-only timing of the dominant kernel + fp64 roof (rank 0) ----
    roofline = e2e = cpu = None
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
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if kern is not None and os.path.exists(tpath):
            with open(tpath) as f:
                tr = json.load(f)
            traffic = ((tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["series_per_launch"]
                       * n_launch)
        roofline = {
            "bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic,
            "traffic_note": ("dram__bytes_read+write of one ncu --set full capture "
                             "(profiles/r01_traffic.json), scaled to series_per_launch; "
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

        # ---- end to end through the public API on pinned host buffers ----
        E = min(args.e2e_series or (262144 if world == 1 else 65536), S)
        while True:
            try:
                hx = torch.empty((E, N_DIMS, T_LEN), dtype=torch.float64, pin_memory=True)
                hf = torch.empty((E, N_FEATS), dtype=torch.float64, pin_memory=True)
                break
            except RuntimeError:            # not enough pinnable host memory: smaller batch
                if E <= 8192:
                    raise
                hx = hf = None
                E //= 2
        hx.copy_(X[:E])
        hx_np, hf_np = hx.numpy(), hf.numpy()
        for _ in range(2):
            fruit.transform(hx_np, out=hf_np)
        torch.cuda.synchronize()
        e_steps = max(2, args.steps)
        t0 = time.perf_counter()
        for _ in range(e_steps):
            fruit.transform(hx_np, out=hf_np)
        torch.cuda.synchronize()
        e_s = (time.perf_counter() - t0) / e_steps
        e2e = {"value": world * E / e_s, "unit": "series/s",
               "h2d_bytes_per_step": int(hx.numel() * 8), "d2h_bytes_per_step": int(hf.numel() * 8),
               "series_per_step": E, "ms_per_step": e_s * 1e3,
               "note": "Fruit.transform(numpy pinned in, numpy pinned out); value scaled by n_gpus"}
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
        e2e["pageable"] = {"value": world * Ep / ((time.perf_counter() - t0) / 2),
                           "unit": "series/s", "series_per_step": Ep,
                           "note": "plain numpy arrays in and out (host memcpy bound)"}
        del px, pf
        if not args.no_cpu_baseline and world == 1:
            rate, cores, sample, _ = cpu_reference(args.cpu_seconds)
            cpu = {"value": rate, "unit": "series/s", "cores": cores, "kind": "port",
                   "sample": sample}

    if rank == 0:
        line = {
            "metric": "time series/sec (features)", "value": value, "unit": "series/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": args.steps * n_chunks * launches_per_call,
            "roofline": roofline, "cpu_baseline": cpu, "fit_seconds": fit_s,
            "collective": collective,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
