"""One-off check in the build container (needs /root/reference): the numpy
restatement of the preparateurs (oracle/preps.py) and the host side of the
product's preparateurs (``_fit_device``: shapes and RNG draws only, no GPU)
against the REAL reference over a sweep of shapes -- short series, one
dimension, lengths around the parameters' corner cases (DIL with no room for
strips, DOT with n >= T, PDD with zero width, windows longer than the series).

    python oracle/check_preps_sweep.py

TEST INFRASTRUCTURE ONLY; prints one line per shape and exits non-zero on the
first mismatch."""
import os
import sys

import numpy as np

if not hasattr(np, "NINF"):
    np.NINF = -np.inf
REF = os.environ.get("FRUITS_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.append(ROOT)
sys.path.append(os.path.join(ROOT, "tests"))

import fruits as ref  # noqa: E402
import torch  # noqa: E402

import fruits_b200 as ours  # noqa: E402
import specs  # noqa: E402
from cases import PREP2_CASES, PREP2_EXACT  # noqa: E402
from oracle import pipeline as orc  # noqa: E402
from oracle import preps as more  # noqa: E402

STATE = {"w1": "_weights1", "b": "_biases", "w2": "_weights2", "kernel": "_kernel",
         "ndim": "_ndim_per_kernel", "dims": "_dims_per_kernel", "weights": "_weights",
         "bias": "_bias_weights", "indices": "_indices", "lengths": "_lengths", "n": "_n",
         "first": "_first", "width": "_width", "w": "_w", "quantile": "_quantile"}
GPU_FIT = ("QTC", "RDW")          # their fit computes on the device


def main():
    bad = 0
    for n, d, t in ((3, 1, 5), (2, 3, 9), (4, 2, 17), (3, 3, 100), (2, 2, 257), (5, 4, 12)):
        rng = np.random.default_rng(n * 1000 + t)
        X = rng.standard_normal((n, d, t)).cumsum(axis=2)
        checked = 0
        for name, desc in PREP2_CASES.items():
            kind, args = desc
            Xc = X
            if kind == "RPE":
                if d < 2:
                    continue
                Xc = np.ascontiguousarray(X[:, :2])
            if kind == "RDW":
                Xc = np.abs(X) + 0.5
            if kind == "RIN" and (args.get("kernel") is not None and d != 3):
                continue
            if kind == "RIN" and args.get("out_dim", -1) > d:
                continue
            if kind == "JLD" and args.get("distribute") and args.get("dim", 1) > d:
                continue
            try:
                p = specs._prep(ref, desc)
                np.random.seed(11)
                p.fit(Xc)
                after_ref = np.random.random()
                r = p.transform(Xc)
            except Exception as exc:            # the reference itself rejects the shape
                try:
                    np.random.seed(11)
                    st = more.fit_prep(desc, Xc)
                    more.transform_prep(desc, st, Xc, orc.RawCache(Xc))
                except Exception:
                    continue
                print(f"  {name} {Xc.shape}: reference raised {type(exc).__name__}, oracle did not")
                continue
            np.random.seed(11)
            st = more.fit_prep(desc, Xc)
            after_orc = np.random.random()
            with np.errstate(invalid="ignore"):
                o = more.transform_prep(desc, st, Xc, orc.RawCache(Xc))
            ok = after_ref == after_orc and o.shape == r.shape
            if ok:
                if kind in PREP2_EXACT:
                    ok = np.array_equal(o, r, equal_nan=True)
                else:
                    scale = np.maximum(np.abs(np.nan_to_num(r)).max(axis=-1, keepdims=True), 1e-300)
                    ok = bool(np.all((np.abs(np.nan_to_num(o) - np.nan_to_num(r)) <= 1e-12 * scale)
                                     | (np.isnan(o) & np.isnan(r))))
            # the product's host-side fit: same draws, same fitted state
            if ok and kind not in GPU_FIT:
                mine = specs._prep(ours, desc)
                np.random.seed(11)
                mine._fit_device(torch.from_numpy(Xc))
                ok = np.random.random() == after_ref
                for key, val in st.items():
                    ok = ok and np.array_equal(np.asarray(getattr(mine, STATE[key])), np.asarray(val))
            checked += 1
            if not ok:
                bad += 1
                print(f"  MISMATCH {name} on {Xc.shape}")
        print(f"shape {(n, d, t)}: {checked} cases checked")
    if bad:
        raise SystemExit(f"{bad} mismatches")
    print("reference, oracle and host-side fits agree on every shape")


if __name__ == "__main__":
    main()
