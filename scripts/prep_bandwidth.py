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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    d, t = 3, 1024
    gen = torch.Generator(device="cuda").manual_seed(1)
    X = torch.randn((n, d, t), dtype=torch.float64, device="cuda", generator=gen).cumsum(dim=2)
    X2 = X[:, :2].contiguous()
    cases = [
        ("DOT(3)            fb_time_mask", P.DOT(3), X),
        ("PDD()             fb_time_mask", P.PDD(), X),
        ("WIN(0.1, 0.9)     fb_time_mask + coquantiles", P.WIN(0.1, 0.9), X),
        ("CTS(5)            fb_time_shift", P.CTS(5), X),
        ("LAG()             fb_lead_lag", P.LAG(), X),
        ("MAV(5)            fb_moving_average", P.MAV(5), X),
        ("MAV(64)           fb_moving_average", P.MAV(64), X),
        ("RIN(width=3)      fb_random_increments", P.RIN(width=3), X),
        ("JLD(3)            fb_dim_project", P.JLD(3), X),
        ("FFN()             fb_ffn + fb_row_stats", P.FFN(), X),
        ("RDW('uniform')    fb_dim_pow", P.RDW("uniform"), X.abs() + 0.5),
        ("SPE(0.5)          fb_wave_embed", P.SPE(0.5), X),
        ("RPE(0.5)          fb_rotate2", P.RPE(0.5), X2),
        ("QTC(0.9)          fb_clip_where", P.QTC(0.9), X),
        ("NRM(True)         fb_nrm_scale", P.NRM(True), X),
    ]
    print(f"# {n} x {d} x {t} float64 ({X.numel() * 8 / 1e9:.2f} GB in), peak {PEAK:.0f} GB/s "
          f"(MEASURED_PEAKS.json), {torch.cuda.get_device_name()}")
    print(f"{'preparateur / kernel':48s} {'ms':>8s} {'GB moved':>9s} {'GB/s':>8s} {'of peak':>8s}")
    np.random.seed(0)
    for name, prep, inp in cases:
        cache = fruits.cache.SharedSeedCache(inp)
        prep._cache = cache
        fit_ms = 0.0
        if prep.requires_fitting:
            fit_ms, _ = timed(lambda: prep._fit_device(inp), reps=1, warm=0)
        if isinstance(prep, P.WIN):
            prep._transform_device(inp)          # (the coquantiles are cached per batch)
        ms, out = timed(lambda: prep._transform_device(inp))
        moved = (inp.numel() + out.numel()) * 8 / 1e9
        gbs = moved / (ms * 1e-3)
        print(f"{name:48s} {ms:8.3f} {moved:9.2f} {gbs:8.0f} {gbs / PEAK:8.2f}"
              + (f"   (fit {fit_ms:.1f} ms)" if fit_ms else ""))
        del out, prep._cache


if __name__ == "__main__":
    main()
