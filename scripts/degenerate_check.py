"""GPU vs oracle on degenerate inputs: zeros under negative exponents
(division by zero -> inf / nan), constant series, huge values (overflow).
Development aid; the stable cases are in tests/test_gpu_parity.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import fruits_b200 as fruits  # noqa: E402
import specs  # noqa: E402
from oracle import pipeline as orc  # noqa: E402


def run(tag, spec, X, force=None):
    if force is None:
        os.environ.pop("FRUITS_B200_JIT", None)
    else:
        os.environ["FRUITS_B200_JIT"] = force
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(1)
    fruit.fit(X)
    np.random.seed(1)
    with np.errstate(all="ignore"):
        of.fit(X)
        ref = of.transform(X)
    res = fruit.transform(X)
    same = (res == ref) | (np.isnan(res) & np.isnan(ref))
    # means: summation order is unspecified in the reference (numba fastmath)
    with np.errstate(all="ignore"):
        same |= np.abs(res - ref) <= 1e-12 * np.maximum(np.abs(ref), 1.0)
    routes = [getattr(s, "_last_launch", ("?",))[0] for s in fruit._slices]
    bad = np.argwhere(~same)
    print(f"{tag:34s} routes={routes} mismatches={len(bad)} of {same.size}", flush=True)
    for i, j in bad[:6]:
        print(f"      row {i} col {j} ({fruit.label(int(j))}): gpu {res[i, j]!r} ref {ref[i, j]!r}")


if __name__ == "__main__":
    rng = np.random.default_rng(2)
    sieves = [["NPI", {"q": [0.5, 1.0]}], ["MPI", {}], ["PPV", {}], ["MAX", {}], ["MIN", {}],
              ["END", {}]]
    base = {"preps": [], "iss": [{"words": ["[-1]", "[1][-1]", "[-11][2]", "[2][-2][1]", "[1][2]"],
                                   "mode": "extended"}], "sieves": sieves, "fit_sample_size": 1.0}
    for n in (40, 4100):
        X = rng.standard_normal((n, 2, 64))
        X[1, 0, 10] = 0.0          # 1/0
        X[2, 0, 0] = 0.0
        X[3, 1, 5:9] = 0.0
        X[4] = 1.5                 # constant series
        X[5, 0, 20] = 1e200        # overflow to inf
        X[6, 0, 20] = -1e200
        X[7, 1, 3] = 1e308
        run(f"reals div0/overflow n={n}", {"slices": [base]}, X)
    arctic = dict(base, iss=[{"words": ["[1][2]", "[-1][2][1]", "[11]"], "mode": "extended",
                              "semiring": "arctic"}])
    X = rng.standard_normal((4100, 2, 64))
    X[5, 0, 20] = 1e308
    X[6, 0, 21] = -1e308
    # (no +-inf in the INPUT: the reference adds 0 * inf = nan for the dimensions a letter
    #  does not use, fruits/iss/semiring.py:326-327 -- inputs are finite by contract)
    run("arctic huge / inf", {"slices": [arctic]}, X)
    std = dict(base, preps=[["STD", {}]], iss=[{"words": ["[1]", "[1][2]"], "mode": "extended"}])
    X = rng.standard_normal((50, 2, 64))
    X[4] = 2.0                     # constant series: std = 0 -> x / 1e-5
    run("STD of a constant series", {"slices": [std]}, X)
