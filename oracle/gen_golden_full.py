"""Freeze BASELINE-size golden vectors from the REAL reference.

TEST INFRASTRUCTURE ONLY (see oracle/gen_golden.py).  Build container only:

    python oracle/gen_golden_full.py C2_full            # ~1 min on 8 cores
    python oracle/gen_golden_full.py C3_full [slices]   # ~1 h on 8 cores

What is frozen (tests/golden/full_<name>.npz):

* ``thresholds``  every fitted threshold of the reference's ``Fruit.fit`` on the
  FULL input of the configuration (C2: 1,000 x 1 x 512, C3: 10,000 x 6 x 1,024;
  ``np.random.seed(0)`` right before ``fit``), slice-major in the order of
  ``gen_golden.ref_thresholds``; ``thr_slices`` holds the slice boundaries.
* ``rows`` / ``features``  the reference's ``Fruit.transform`` of the sampled
  rows (transform is independent per series: fruits/iss/semiring.py:184,
  fruits/sieving/increment.py:121; STD is per series and dimension,
  fruits/preparation/transform.py:92-158).
* C2 only (the full transform finishes in seconds): ``counts`` = the
  integer-valued NPI columns of ALL rows as uint16, ``count_cols`` their
  column indices, and ``row_sha_slice<i>`` = sha256 per row of the bit-exact
  columns (``exact_cols_slice<i>``: NPI counts and END) of the arctic slice.
* ``source`` says which program produced each slice ("reference" = the numba
  package at /root/reference; "oracle" only if the reference did not finish).

The oracle (oracle/pipeline.py) is checked against the reference on the way.
"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden as gg  # noqa: E402  (imports the real reference, asserts its origin)

ref, orc, specs, GOLD = gg.ref, gg.orc, gg.specs, gg.GOLD


def slice_thresholds(slc):
    rows = []
    for sieves in slc._sieves_extended:
        for sv in sieves:
            sv = gg.unwrap(sv)
            q = getattr(sv, "_quantiles", None)
            if q is None:
                q = getattr(sv, "_q", [])
            rows.append(np.asarray(q, dtype=np.float64).ravel())
    return np.concatenate(rows) if rows else np.zeros(0)


def fit_by_slice(fruit, X, log, which=None, resume=None):
    """``Fruit.fit`` (fruits/fruit.py:121-136) slice by slice with timing; the
    global RNG is consumed in the same order as one ``fruit.fit(X)`` call."""
    cache = ref.cache.SharedSeedCache(X)
    out = []
    for i, slc in enumerate(fruit):
        t0 = time.perf_counter()
        slc.fit(X, cache=cache)
        dt = time.perf_counter() - t0
        thr = slice_thresholds(slc)
        log(f"  slice {i}: reference fit {dt:.1f} s, {thr.size} thresholds")
        if os.environ.get("GOLDEN_CHECKPOINT"):          # a long run keeps what it has
            np.save(os.path.join(os.environ["GOLDEN_CHECKPOINT"], f"thr_slice{i}.npy"), thr)
        out.append((thr, dt))
    fruit._fitted = True
    return out


def main(name, n_rows):
    spec = specs.SPECS[name]
    X = specs.make_input(name)
    log = lambda s: print(s, flush=True)   # noqa: E731
    log(f"[{name}] input {X.shape}")
    fruit = specs.build_fruit(ref, spec)
    np.random.seed(0)
    fits = fit_by_slice(fruit, X, log)
    thr = np.concatenate([f[0] for f in fits])
    bounds = np.cumsum([0] + [f[0].size for f in fits])
    rows = np.unique(np.linspace(0, X.shape[0] - 1, n_rows).astype(np.int64))
    out = {"thresholds": thr, "thr_slices": bounds, "rows": rows,
           "fit_seconds": np.array([f[1] for f in fits]),
           "xsha": np.array(gg.sha(X)), "n": np.array(X.shape[0]),
           "source": np.array("reference"), "nfeatures": np.array(fruit.nfeatures())}
    full = X.shape[0] <= 1000
    t0 = time.perf_counter()
    r = fruit.transform(X if full else np.ascontiguousarray(X[rows]))
    log(f"  reference transform of {r.shape[0]} rows: {time.perf_counter() - t0:.1f} s")
    out["features"] = r[rows] if full else r
    labels = [fruit.label(i) for i in range(fruit.nfeatures())]
    count_cols = np.array([i for i, s in enumerate(labels) if "| NPI" in s], dtype=np.int64)
    if full:
        cnt = r[:, count_cols]
        assert np.all(cnt == np.round(cnt)) and cnt.min() >= 0 and cnt.max() < 65536
        out["counts"] = cnt.astype(np.uint16)
        out["count_cols"] = count_cols
        # the bit-exact (arctic, unweighted) slices: one hash per row
        col = 0
        for si, slc in enumerate(spec["slices"]):
            k = fruit.get_slice(si).nfeatures()
            if slc["iss"][0].get("semiring") == "arctic":
                # (MPI means are sums whose order numba's fastmath leaves open: not hashed;
                # + 0.0 turns -0.0 into +0.0, the sign of zero is unspecified under fastmath)
                exact = np.array([c for c in range(col, col + k) if "| MPI" not in labels[c]])
                out[f"exact_cols_slice{si}"] = exact
                out[f"row_sha_slice{si}"] = np.array(
                    [hashlib.sha256(np.ascontiguousarray(r[j, exact] + 0.0).tobytes()).hexdigest()[:16]
                     for j in range(r.shape[0])])
            col += k
    path = os.path.join(GOLD, f"full_{name}.npz")
    np.savez_compressed(path, **out)
    log(f"written {path} ({os.path.getsize(path) / 1e6:.2f} MB)")

    if not full:
        return      # the oracle's numpy fit of 10 M values per quantile takes hours
    # the oracle against the reference at this size (thresholds; sampled rows)
    of = orc.OracleFruit(spec)
    np.random.seed(0)
    t0 = time.perf_counter()
    of.fit(X)
    log(f"  oracle fit {time.perf_counter() - t0:.1f} s")
    ot = gg.orc_thresholds(of)
    gg.check_close(ot, thr, f"thresholds {name} (oracle vs reference)", rtol=1e-11)
    o = of.transform(np.ascontiguousarray(X[rows]))
    want = out["features"]
    scale = np.maximum(np.abs(want), 1.0)
    bad = np.abs(o - want) > 1e-9 * scale
    log(f"  oracle features on {len(rows)} rows: {int(bad.sum())} of {bad.size} beyond 1e-9, "
        f"exact {np.array_equal(o, want)}")
    if bad.mean() > 1e-3:
        raise SystemExit("ORACLE MISMATCH")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "C2_full"
    main(which, int(sys.argv[2]) if len(sys.argv) > 2 else (64 if which.startswith("C2") else 16))
