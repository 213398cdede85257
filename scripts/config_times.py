"""fit / transform wall time of the BASELINE.json configurations on one GPU,
with the plan-specialised kernel (default) and with the generic kernel.

    python scripts/config_times.py [C1_readme ...]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import fruits_b200 as fruits  # noqa: E402
import specs  # noqa: E402

SIZES = {"C1_readme": 200, "C2_reduced": 1000, "C2_cos": 1000, "C2_full": 1000, "C3_cos": 10000, "C3_full": 10000, "C3_general": 10000,
         "C4_twi": 100000, "C5_sweep": 65536}


def run(name: str, jit: bool):
    os.environ["FRUITS_B200_JIT"] = "1" if jit else "0"
    n = SIZES[name]
    Xh = specs.make_input(name, n)
    X = torch.from_numpy(Xh).cuda()
    fruit = specs.build_fruit(fruits, specs.SPECS[name])
    np.random.seed(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fruit.fit(X)
    torch.cuda.synchronize()
    fit_s = time.perf_counter() - t0
    out = fruit.transform(X)           # includes JIT compile / module load
    torch.cuda.synchronize()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fruit.transform(X)
    torch.cuda.synchronize()
    tr_s = (time.perf_counter() - t0) / reps
    return fit_s, tr_s, out


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(SIZES)):
        fit_j, tr_j, out_j = run(name, True)
        fit_g, tr_g, out_g = run(name, False)
        same = bool(torch.equal(out_j, out_g))
        err = float((out_j - out_g).abs().max())
        print(f"{name}: n={SIZES[name]} features={out_j.shape[1]}  "
              f"jit: fit {fit_j:.3f} s transform {tr_j * 1e3:.2f} ms ({SIZES[name] / tr_j:.0f} series/s)  "
              f"generic: fit {fit_g:.3f} s transform {tr_g * 1e3:.2f} ms ({SIZES[name] / tr_g:.0f} series/s)  "
              f"identical={same} maxabs={err:.3g}", flush=True)
