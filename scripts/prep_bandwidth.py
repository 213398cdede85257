"""HBM bandwidth of the preparateur kernels of ``csrc/prep_more.cu``.

Every kernel is timed alone with CUDA events on a batch far larger than L2
(default 131,072 x 3 x 1,024 float64 = 3.2 GB in), after three warm-up
launches, best of five; the figure is the ALGORITHMIC traffic -- bytes of X
read once plus bytes of the prepared copy written once -- divided by the time,
against the measured copy bandwidth of MEASURED_PEAKS.json (6,530 GB/s).

    python scripts/prep_bandwidth.py [n_series] > gpurun_out/prep_bandwidth.log
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fruits_b200 as fruits  # noqa: E402

P = fruits.preparation
PEAK = 6530.0
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "MEASURED_PEAKS.json")
if os.path.exists(peaks):
    PEAK = float(json.load(open(peaks)).get("hbm_gbs", PEAK))


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    best = float("inf")
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, out


def measure(n: int = 131072, d: int = 3, t: int = 1024):
    """-> (header dict, [ {name, kernel, ms, gb_moved, gbs, frac, fit_ms} ])"""
    gen = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn((n, d, t), dtype=torch.float64, device="cuda", generator=gen).cumsum(dim=2)
    X2 = X[:, :2].contiguous()
    cases = [
        ("DOT(3)", "fb_time_mask", P.DOT(3), X),
        ("PDD()", "fb_time_mask", P.PDD(), X),
        ("WIN(0.1, 0.9)", "fb_time_mask (coquantiles cached)", P.WIN(0.1, 0.9), X),
        ("CTS(5)", "fb_time_shift", P.CTS(5), X),
        ("LAG()", "fb_lead_lag", P.LAG(), X),
        ("MAV(5)", "fb_moving_average", P.MAV(5), X),
        ("MAV(64)", "fb_moving_average", P.MAV(64), X),
        ("RIN(width=3)", "fb_random_increments", P.RIN(width=3), X),
        ("JLD(3)", "fb_dim_project", P.JLD(3), X),
        ("FFN()", "fb_ffn + fb_row_stats", P.FFN(), X),
        ("RDW('uniform')", "fb_dim_pow", P.RDW("uniform"), X.abs() + 0.5),
        ("SPE(0.5)", "fb_wave_embed", P.SPE(0.5), X),
        ("RPE(0.5)", "fb_rotate2", P.RPE(0.5), X2),
        ("QTC(0.9)", "fb_clip_where", P.QTC(0.9), X),
        ("NRM(True)", "fb_nrm_scale", P.NRM(True), X),
    ]
    head = {"series": n, "dims": d, "length": t, "gb_in": X.numel() * 8 / 1e9,
            "peak_gbs": PEAK, "peak_source": "MEASURED_PEAKS.json hbm_gbs",
            "bytes": "input read once + prepared copy written once",
            "device": torch.cuda.get_device_name()}
    rows = []
    state = np.random.get_state()
    np.random.seed(0)
    for name, kernel, prep, inp in cases:
        prep._cache = fruits.cache.SharedSeedCache(inp)
        fit_ms = 0.0
        if prep.requires_fitting:
            fit_ms, _ = timed(lambda: prep._fit_device(inp), reps=1, warm=0)
        if isinstance(prep, P.WIN):
            prep._transform_device(inp)          # (the coquantiles are cached per batch)
        ms, out = timed(lambda: prep._transform_device(inp))
        moved = (inp.numel() + out.numel()) * 8 / 1e9
        rows.append({"preparateur": name, "kernel": kernel, "ms": ms, "gb_moved": moved,
                     "gbs": moved / (ms * 1e-3), "frac": moved / (ms * 1e-3) / PEAK,
                     "fit_ms": fit_ms})
        del out, prep._cache
    np.random.set_state(state)
    return head, rows


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    head, rows = measure(n)
    print(f"# {head['series']} x {head['dims']} x {head['length']} float64 ({head['gb_in']:.2f} GB "
          f"in), peak {PEAK:.0f} GB/s (MEASURED_PEAKS.json), {head['device']}")
    print(f"{'preparateur':16s} {'kernel':36s} {'ms':>8s} {'GB moved':>9s} {'GB/s':>8s} "
          f"{'of peak':>8s}")
    for r in rows:
        print(f"{r['preparateur']:16s} {r['kernel']:36s} {r['ms']:8.3f} {r['gb_moved']:9.2f} "
              f"{r['gbs']:8.0f} {r['frac']:8.2f}"
              + (f"   (fit {r['fit_ms']:.1f} ms)" if r['fit_ms'] else ""))


if __name__ == "__main__":
    main()
