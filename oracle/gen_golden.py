"""Pin the oracle against the REAL reference and freeze golden vectors.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where the reference
checkout is mounted at /root/reference (it does not exist on the GPU box):

    python oracle/gen_golden.py

The script
  1. imports the unmodified reference package (with the one-line
     ``np.NINF`` shim that numpy >= 2 needs),
  2. runs it on seeded inputs for every hot-path row of SURVEY.md section 8,
  3. checks the oracle (oracle/pipeline.py + fruits_oracle.c) against it --
     bit-exact where the reference is deterministic -- and aborts on any
     mismatch,
  4. writes the reference outputs to tests/golden/*.npz.
"""
import hashlib
import os
import sys

import numpy as np

if not hasattr(np, "NINF"):
    np.NINF = -np.inf  # the reference uses np.NINF (fruits/sieving/segment.py:72)

REF = os.environ.get("FRUITS_REF", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.append(ROOT)

import fruits as ref  # noqa: E402  (the real reference)

assert os.path.realpath(ref.__file__).startswith(os.path.realpath(REF)), ref.__file__

from oracle import pipeline as orc  # noqa: E402
sys.path.append(os.path.join(ROOT, "tests"))
import specs  # noqa: E402  (tests/specs.py)

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def check_equal(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape or not np.array_equal(a, b, equal_nan=True):
        bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
        raise SystemExit(f"ORACLE MISMATCH (exact) in {what}: {len(bad)} of "
                         f"{a.size} differ, first at {bad[:3].tolist()}")
    print(f"  ok (bit-exact)  {what}")


def check_close(a, b, what, rtol=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0 and b.size == 0:
        print(f"  ok (empty)  {what}")
        return
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    fin = np.where(np.isfinite(b), b, 0.0)
    scale = np.maximum(np.abs(fin), np.max(np.abs(fin), axis=-1, keepdims=True))
    with np.errstate(invalid="ignore"):
        err = np.where(same, 0.0, np.abs(a - b))
    if a.shape != b.shape or not np.all(err <= rtol * scale + 1e-300):
        raise SystemExit(f"ORACLE MISMATCH (rtol {rtol}) in {what}: max rel "
                         f"{np.max(err / (scale + 1e-300))}")
    exact = np.array_equal(a, b)
    print(f"  ok ({'bit-exact' if exact else f'rtol {rtol}'})  {what}")


# ---------------------------------------------------------------------------
def gen_words():
    print("[words]")
    out = {}
    for w, d in [(1, 1), (2, 1), (3, 1), (4, 1), (5, 1), (6, 1), (9, 1),
                 (2, 2), (3, 2), (4, 2), (5, 2), (6, 2), (2, 3), (3, 3),
                 (4, 3), (5, 3), (3, 4), (2, 12)]:
        r = [str(x) for x in ref.words.of_weight(w, d)]
        o = orc.of_weight(w, d)
        assert r == o, (w, d)
        out[f"of_weight_{w}_{d}"] = np.array("|".join(r))
        plan = ref.iss.CachePlan(ref.words.of_weight(w, d))._plan
        assert plan == orc.cache_plan(o), (w, d)
        out[f"plan_{w}_{d}"] = np.array(plan, dtype=np.int64)
    base = [24 * "[1]", 24 * "[2]", 12 * "[1][2]", 12 * "[2][1]", "[112][2][1]"]
    r = [str(x) for x in ref.words.alternate_sign(
        [ref.words.SimpleWord(b) for b in base])]
    assert r == orc.alternate_sign(base)
    out["alternate_sign"] = np.array("|".join(r))
    out["alternate_sign_plan"] = np.array(
        ref.iss.CachePlan([ref.words.SimpleWord(x) for x in r])._plan)
    assert list(out["alternate_sign_plan"]) == orc.cache_plan(r)
    # exponent matrices of awkward word strings
    for i, s in enumerate(["[12][122]", "[-1-12][(-11)3]", "[(10)(10)2][-2]",
                           "[1][-1][11-2]"]):
        m = np.array(list(ref.words.SimpleWord(s)), dtype=np.int32)
        assert np.array_equal(m, orc.parse_word(s)), s
        out[f"parse_{i}"] = m
        out[f"parse_{i}_str"] = np.array(s)
    np.savez_compressed(os.path.join(GOLD, "words.npz"), **out)
    print("  ok  word enumeration / cache plans / parser")


# ---------------------------------------------------------------------------
ISS_CASES = {
    # name: (iss description, input shape, input kind)
    "reals_w3d3_ext": ({"words": {"of_weight": [3, 3]}, "mode": "extended"},
                       (5, 3, 96), "normal"),
    "reals_w4d2_single": ({"words": {"of_weight": [4, 2]}, "mode": "single"},
                          (3, 2, 64), "walk"),
    "reals_neg": ({"words": ["[-1]", "[1][-1]", "[-11][2-2][1]", "[-1-1][22]"],
                   "mode": "extended"}, (4, 2, 50), "uniform1"),
    "reals_indices": ({"words": {"of_weight": [3, 2]}, "mode": "extended",
                       "weighting": ["Indices", {}]}, (4, 2, 80), "std"),
    "reals_indices_total": ({"words": {"of_weight": [3, 2]}, "mode": "extended",
                             "weighting": ["Indices", {"total": True}]},
                            (4, 2, 80), "std"),
    "reals_L1": ({"words": {"of_weight": [4, 1]}, "mode": "extended",
                  "weighting": ["L1", {}]}, (4, 1, 128), "walk"),
    "reals_L2_total_alpha": ({"words": ["[1][1][1]", "[1][11]", "[11][1]"],
                              "mode": "extended",
                              "weighting": ["L2", {"total": True, "scale": 5}],
                              "alphas": [[0.5, 1.0, 2.0], [0.5, 1.0],
                                         [0.25, 1.0]]},
                             (3, 1, 64), "walk"),
    "reals_plateaus": ({"words": ["[1][2]", "[12][1]"], "mode": "single",
                        "weighting": ["Plateaus", {"n": 4}]},
                       (3, 2, 64), "normal"),
    "arctic_alt24": ({"words": {"alternate_sign": [24 * "[1]", 12 * "[1][2]"]},
                      "mode": "extended", "semiring": "arctic"},
                     (4, 2, 100), "walk"),
    "arctic_exp": ({"words": ["[111]", "[1][112]", "[12][111][2]", "[-1-1-1][2]"],
                    "mode": "extended", "semiring": "arctic"},
                   (4, 2, 70), "normal"),
    "arctic_single": ({"words": ["[1][2][1]", "[2]", "[1][2]"],
                       "mode": "single", "semiring": "arctic"},
                      (4, 2, 70), "normal"),
    "arctic_indices": ({"words": ["[1][1]", "[1][2][1]"], "mode": "extended",
                        "semiring": "arctic", "weighting": ["Indices", {}]},
                       (3, 2, 60), "normal"),
    "arctic_L1_total": ({"words": ["[1][1]", "[1][2][1]"], "mode": "extended",
                         "semiring": "arctic",
                         "weighting": ["L1", {"total": True}]},
                        (3, 2, 60), "walk"),
}


def make_iss_input(shape, kind, seed=7):
    rng = np.random.default_rng(seed)
    if kind == "normal":
        return rng.standard_normal(shape)
    if kind == "walk":
        return rng.standard_normal(shape).cumsum(axis=2)
    if kind == "uniform1":
        return rng.random(shape) + 0.5
    if kind == "std":
        x = rng.standard_normal(shape).cumsum(axis=2)
        return (x - x.mean(axis=2, keepdims=True)) / x.std(axis=2, keepdims=True)
    raise ValueError(kind)


def gen_iss():
    print("[iss]")
    out = {}
    for name, (desc, shape, kind) in ISS_CASES.items():
        X = make_iss_input(shape, kind)
        r = specs.build_iss(ref, desc).transform(X)
        o = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
        weighted = desc.get("weighting") is not None
        if weighted:
            check_close(o, r, f"iss {name}")
        else:
            check_equal(o, r, f"iss {name}")
        out[name] = r
        out[name + "_xsha"] = np.array(sha(X))
    np.savez_compressed(os.path.join(GOLD, "iss.npz"), **out)


# ---------------------------------------------------------------------------
SIEVE_CASES = {
    "npi_default": ["NPI", {}],
    "npi_q": ["NPI", {"q": [0.25, 0.5, 1.0], "inc": 1}],
    "npi_inc0": ["NPI", {"q": [0.5, 1.0], "inc": 0}],
    "npi_inc2": ["NPI", {"q": [-1.0, 0.3, 1.0], "inc": 2}],
    "npi_incm1": ["NPI", {"q": [0.5, 1.0], "inc": -1}],
    "npi_cuts": ["NPI", {"cut": [10, -1, 30], "q": [0.0, 0.6, 1.0]}],
    "npi_cocuts": ["NPI", {"cut": [0.3, 0.7, -1]}],
    "mpi_default": ["MPI", {}],
    "mpi_q": ["MPI", {"q": [0.5, 1.0], "inc": 2}],
    "mpi_cuts": ["MPI", {"cut": [0.5, -1], "q": [0.2, 0.8],
                         "coquantile_norm": "L1"}],
    "xpi": ["XPI", {"q": [0.5, 1.0]}],
    "lpi": ["LPI", {"cut": [20, -1]}],
    "max_default": ["MAX", {}],
    "max_q": ["MAX", {"q": [-1.0, 0.5, 1.0]}],
    "max_cuts": ["MAX", {"cut": [15, 0.5, -1]}],
    "min_default": ["MIN", {}],
    "min_q": ["MIN", {"q": [-1.0, 0.5, 1.0]}],
    "min_cuts": ["MIN", {"cut": [15, -1], "coquantile_norm": "L1"}],
    "end_default": ["END", {}],
    "end_cuts": ["END", {"cut": [1, 17, 0.4, -1]}],
    "ppv_default": ["PPV", {}],
    "ppv_multi": ["PPV", {"quantile": [0.2, 0.0, 0.9],
                          "constant": [False, True, False]}],
    "ppv_segments": ["PPV", {"quantile": [0.2, 0.5, 0.9], "segments": True}],
}


def gen_sieves():
    print("[sieves]")
    out = {}
    rng = np.random.default_rng(11)
    raw = rng.standard_normal((12, 2, 48)).cumsum(axis=2)
    Y = rng.standard_normal((12, 48)).cumsum(axis=1)
    Y[3] = Y[3, 0] + 0.0 * Y[3]  # constant row
    Y -= np.median(Y, axis=1, keepdims=True) - 0.25  # every row straddles 0.25
    Y[5, 7:11] = Y[5, 6]  # ties
    out["raw_xsha"] = np.array(sha(raw))
    out["Y_xsha"] = np.array(sha(Y))
    for name, desc in SIEVE_CASES.items():
        sv = specs._sieve(ref, desc)
        sv._cache = ref.cache.SharedSeedCache(raw)
        np.random.seed(3)
        sv.fit(Y)
        r = sv.transform(Y)
        o = orc.OracleSieve(desc)
        np.random.seed(3)
        o.fit(Y)
        oo = o.transform(Y, orc.RawCache(raw))
        if desc[0] in ("MPI", "XPI"):
            check_close(oo, r, f"sieve {name}", rtol=1e-13)
        else:
            check_equal(oo, r, f"sieve {name}")
        out[name] = r
        if desc[0] == "PPV":
            out[name + "_thr"] = np.array(sv._q, dtype=np.float64)
        else:
            out[name + "_thr"] = np.array(sv._quantiles, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "sieves.npz"), **out)


# ---------------------------------------------------------------------------
PREP_CASES = {
    "inc": ["INC", {}],
    "inc_shift3_depth2": ["INC", {"shift": 3, "depth": 2}],
    "inc_nopad": ["INC", {"zero_padding": False}],
    "std": ["STD", {}],
    "std_novar": ["STD", {"var": False}],
    "nrm": ["NRM", {}],
    "new_inc": ["NEW", ["INC", {}]],
    "new_none": ["NEW", None],
}


def gen_preps():
    print("[preparateurs]")
    out = {}
    X = np.random.default_rng(5).standard_normal((6, 3, 40)).cumsum(axis=2)
    X[2, 1] = 4.0
    out["xsha"] = np.array(sha(X))
    for name, desc in PREP_CASES.items():
        p = specs._prep(ref, desc)
        p.fit(X)
        r = p.transform(X)
        o = orc.apply_prep(desc, X)
        check_equal(o, r, f"prep {name}")
        out[name] = r
    np.savez_compressed(os.path.join(GOLD, "preps.npz"), **out)


# ---------------------------------------------------------------------------
PIPE_CASES = {
    # name: (spec name, number of series)
    "C1_readme": ("C1_readme", 200),
    "C2_reduced": ("C2_reduced", 24),
    "C3_general": ("C3_general", 4),
    "C4_twi": ("C4_twi", 8),
    "C5_sweep": ("C5_sweep", 32),
}


def ref_thresholds(fruit):
    rows = []
    for slc in fruit:
        for sieves in slc._sieves_extended:
            for sv in sieves:
                q = getattr(sv, "_quantiles", None)
                if q is None:
                    q = getattr(sv, "_q", [])
                rows.append(np.asarray(q, dtype=np.float64).ravel())
    return np.concatenate(rows) if rows else np.zeros(0)


def orc_thresholds(of):
    rows = []
    for slc in of.slices:
        for sieves in slc.sieves_extended:
            for sv in sieves:
                q = sv.fitted_q if sv.name == "PPV" else sv.quantiles
                rows.append(np.asarray(q, dtype=np.float64).ravel())
    return np.concatenate(rows) if rows else np.zeros(0)


def gen_pipelines():
    print("[pipelines]")
    for name, (spec_name, n) in PIPE_CASES.items():
        spec = specs.SPECS[spec_name]
        X = specs.make_input(spec_name, n)
        fruit = specs.build_fruit(ref, spec)
        np.random.seed(0)
        fruit.fit(X)
        r = fruit.transform(X)
        of = orc.OracleFruit(spec)
        np.random.seed(0)
        of.fit(X)
        o = of.transform(X)
        assert fruit.nfeatures() == of.nfeatures() == r.shape[1]
        weighted = any(i.get("weighting") for s in spec["slices"] for i in s["iss"])
        has_mpi = any(sv[0] == "MPI" for s in spec["slices"] for sv in s["sieves"])
        rt, ot = ref_thresholds(fruit), orc_thresholds(of)
        if weighted or has_mpi:
            check_close(ot, rt, f"thresholds {name}", rtol=1e-11)
            # features: counts may flip when a value sits on a threshold
            scale = np.maximum(np.abs(r), 1.0)
            bad = np.abs(o - r) > 1e-9 * scale
            frac = bad.mean()
            print(f"  features {name}: {bad.sum()} of {bad.size} beyond 1e-9 "
                  f"(exact {np.array_equal(o, r)}, max rel "
                  f"{np.max(np.abs(o - r) / scale):.2e})")
            if frac > 1e-3:
                raise SystemExit(f"ORACLE MISMATCH in pipeline {name}")
        else:
            check_equal(ot, rt, f"thresholds {name}")
            check_equal(o, r, f"features {name}")
        labels = np.array("|".join(
            fruit.label(i) for i in
            sorted(set(np.linspace(0, r.shape[1] - 1, 23).astype(int)))))
        np.savez_compressed(
            os.path.join(GOLD, f"pipeline_{name}.npz"),
            features=r, thresholds=rt, xsha=np.array(sha(X)), n=np.array(n),
            nfeatures=np.array(fruit.nfeatures()), labels=labels,
            summary=np.array(fruit.summary()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["words", "iss", "sieves", "preps", "pipelines"]
    for w in which:
        globals()["gen_" + w]()
    print("golden vectors written to", GOLD)
