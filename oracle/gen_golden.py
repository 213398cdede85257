"""Pin the oracle against the REAL reference and freeze golden vectors.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where the reference
checkout is mounted at /root/reference (it does not exist on the GPU box):

    python oracle/gen_golden.py            # words iss sieves preps pipelines
    python oracle/gen_golden.py cos extra preps2 cos2 argmax corbeille   # the other groups

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
from cases import (COS_CASES, COS_PIPE_CASES, EXTRA_PIPE_CASES, ISS_CASES, PIPE_CASES, PREP_CASES, SIEVE_CASES, SUMMING_SIEVES, IMPLICIT_SIEVES, sieve_kind, unwrap,  # noqa: E402
                   make_iss_input, make_prep_input, make_sieve_input)

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
def gen_sieves():
    print("[sieves]")
    out = {}
    raw, Y = make_sieve_input()
    out["raw_xsha"] = np.array(sha(raw))
    out["Y_xsha"] = np.array(sha(Y))
    for name, desc in SIEVE_CASES.items():
        sv = specs._sieve(ref, desc)
        sv._cache = ref.cache.SharedSeedCache(raw)
        np.random.seed(3)
        sv.fit(Y)
        r = sv.transform(Y)
        o = orc.make_sieve(desc)
        np.random.seed(3)
        o.fit(Y)
        oo = o.transform(Y, orc.RawCache(raw))
        if sieve_kind(desc) in SUMMING_SIEVES:
            check_close(oo, r, f"sieve {name}", rtol=1e-13)
        else:
            check_equal(oo, r, f"sieve {name}")
        out[name] = r
        if sieve_kind(desc) in IMPLICIT_SIEVES:
            out[name + "_thr"] = np.array(unwrap(sv)._q, dtype=np.float64)
        else:
            out[name + "_thr"] = np.array(unwrap(sv)._quantiles, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "sieves.npz"), **out)


# ---------------------------------------------------------------------------
def gen_preps():
    print("[preparateurs]")
    out = {}
    X = make_prep_input()
    out["xsha"] = np.array(sha(X))
    for name, desc in PREP_CASES.items():
        p = specs._prep(ref, desc)
        p.fit(X)
        r = p.transform(X)
        o = orc.apply_prep(desc, X)
        check_equal(o, r, f"prep {name}")
        out[name] = r
    np.savez_compressed(os.path.join(GOLD, "preps.npz"), **out)


def gen_preps2():
    """The preparateurs beside INC / STD / NRM: fit under a seed on X, transform
    X and a second batch, the RNG state behind fit."""
    from oracle import preps as more
    from cases import PREP2_CASES, PREP2_EXACT, make_prep2_inputs
    print("[preparateurs 2]")
    out = {}
    for name, desc in PREP2_CASES.items():
        X, X2 = make_prep2_inputs(name)
        p = specs._prep(ref, desc)
        np.random.seed(7)
        p.fit(X)
        state_r = np.random.random()              # where the generator stands after fit
        r, r2 = p.transform(X), p.transform(X2)
        np.random.seed(7)
        st = more.fit_prep(desc, X)
        state_o = np.random.random()
        assert state_r == state_o, f"{name}: fit consumed the RNG differently"
        o = more.transform_prep(desc, st, X, orc.RawCache(X))
        o2 = more.transform_prep(desc, st, X2, orc.RawCache(X2))
        for a, b, what in ((o, r, name), (o2, r2, name + " (second batch)")):
            if desc[0] in PREP2_EXACT:
                check_equal(a, b, f"prep {what}")
            else:
                check_close(a, b, f"prep {what}", rtol=1e-12)
        out[name], out[name + "_2"], out[name + "_rng"] = r, r2, np.array(state_r)
    # L1 / L2 weightings with a Python transform of the lookup (fruits/iss/weighting.py:155-156)
    X = make_prep_input()
    for key, w in (("L1_sqrt_relative", ref.iss.weighting.L1(relative=True, transform=np.sqrt)),
                   ("L2_log1p", ref.iss.weighting.L2(transform=np.log1p, scale=5, total=True))):
        iss = ref.ISS([ref.words.SimpleWord("[1][2]"), ref.words.SimpleWord("[3][1][1]")],
                      mode=ref.ISSMode.EXTENDED, weighting=w)
        out["iss_" + key] = iss.transform(X)
        w._cache = ref.cache.SharedSeedCache(X)
        out["lookup_" + key] = w.get_lookup(X)
    np.savez_compressed(os.path.join(GOLD, "preps2.npz"), **out)


# ---------------------------------------------------------------------------
def ref_thresholds(fruit):
    rows = []
    for slc in fruit:
        for sieves in slc._sieves_extended:
            for sv in sieves:
                sv = unwrap(sv)                      # sieve wrappers: the wrapped sieve's
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
                sv = unwrap(sv)
                q = sv.fitted_q if sv.name in ("PPV", "CPV") else sv.quantiles
                rows.append(np.asarray(q, dtype=np.float64).ravel())
    return np.concatenate(rows) if rows else np.zeros(0)


def gen_cos():
    """Cosine weighted ISS: ISS cases and the
    CosWISS slices of experiments/fruit_reduced.py."""
    print("[coswiss]")
    out = {}
    for name, (desc, shape, kind) in COS_CASES.items():
        X = make_iss_input(shape, kind)
        iss = specs.build_iss(ref, desc)
        r = iss.transform(X)
        o = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
        check_close(o, r, f"coswiss {name}", rtol=1e-11)
        assert iss.n_iterated_sums() == orc.n_iterated_sums(desc) == r.shape[0]
        for i in (0, r.shape[0] - 1):
            assert iss.label(i) == orc.iss_label(desc, i), (iss.label(i), orc.iss_label(desc, i))
        out[name] = r
        out[name + "_xsha"] = np.array(sha(X))
        out[name + "_labels"] = np.array("|".join(iss.label(i) for i in range(r.shape[0])))
    np.savez_compressed(os.path.join(GOLD, "cos.npz"), **out)
    gen_pipelines(COS_PIPE_CASES)


def gen_argmax():
    """Arctic(argmax=True): ISS cases and one pipeline slice."""
    from cases import ARGMAX_CASES
    print("[arctic argmax]")
    out = {}
    for name, (desc, shape, kind) in ARGMAX_CASES.items():
        X = make_iss_input(shape, kind)
        iss = specs.build_iss(ref, desc)
        r = iss.transform(X)
        o = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
        assert iss.n_iterated_sums() == orc.n_iterated_sums(desc) == r.shape[0]
        if desc.get("weighting") is None:
            check_equal(o, r, f"argmax {name}")
        else:
            check_close(o, r, f"argmax {name}", rtol=1e-12)
        out[name] = r
        out[name + "_xsha"] = np.array(sha(X))
    np.savez_compressed(os.path.join(GOLD, "argmax.npz"), **out)
    gen_pipelines({"R_argmax": EXTRA_PIPE_CASES["R_argmax"]})


def gen_letters():
    """Generic words (Python letters) through ISS.transform in all three semirings."""
    from cases import LETTER_CASES
    print("[letters]")
    out = {}
    for name, (desc, shape, kind) in LETTER_CASES.items():
        X = make_iss_input(shape, kind)
        iss = specs.build_iss(ref, desc)
        r = iss.transform(X)
        o = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
        assert iss.n_iterated_sums() == orc.n_iterated_sums(desc) == r.shape[0]
        check_equal(o, r, f"letters {name}")
        out[name] = r
        out[name + "_xsha"] = np.array(sha(X))
    np.savez_compressed(os.path.join(GOLD, "letters.npz"), **out)


def gen_cos2():
    """The randomised CosWISS variants (ffn_size, dropout): fit under a seed,
    transform, where the generator stands after fit."""
    import copy
    from cases import COS_RANDOM_CASES
    print("[coswiss, randomised]")
    out = {}
    for name, (desc, shape, kind) in COS_RANDOM_CASES.items():
        X = make_iss_input(shape, kind)
        iss = specs.build_iss(ref, desc)
        assert iss.requires_fitting
        np.random.seed(3)
        iss.fit(X)
        state_r = np.random.random()
        r = iss.transform(X)
        odesc = copy.deepcopy(desc)
        np.random.seed(3)
        odesc["_state"] = orc.coswiss_fit(odesc, X)
        assert np.random.random() == state_r, f"{name}: fit consumed the RNG differently"
        o = np.stack(list(orc.iss_iter(X, odesc, orc.RawCache(X))))
        check_close(o, r, f"coswiss {name}", rtol=1e-11)
        out[name], out[name + "_rng"] = r, np.array(state_r)
        out[name + "_xsha"] = np.array(sha(X))
    np.savez_compressed(os.path.join(GOLD, "cos2.npz"), **out)
    gen_pipelines({"R_cosrand": EXTRA_PIPE_CASES["R_cosrand"]})


def gen_extra():
    """Pipelines of the rank 2-3 components (Bayesian semiring, CUR / CPV / XPI /
    LPI, sieve wrappers, chained ISS)."""
    gen_pipelines(EXTRA_PIPE_CASES)


def gen_pipelines(cases=None):
    print("[pipelines]")
    for name, (spec_name, n) in (PIPE_CASES if cases is None else cases).items():
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
        weighted = any(i.get("weighting") or i.get("coswiss")
                       for s in spec["slices"] for i in s["iss"])
        has_mpi = any(sieve_kind(sv) in SUMMING_SIEVES for s in spec["slices"]
                      for sv in s["sieves"])
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
        try:
            labels = np.array("|".join(
                fruit.label(i) for i in
                sorted(set(np.linspace(0, r.shape[1] - 1, 23).astype(int)))))
        except IndexError:
            # Arctic(argmax=True): ISS._label asks the cache plan for rows it does not hold
            labels = np.array("IndexError")
        np.savez_compressed(
            os.path.join(GOLD, f"pipeline_{name}.npz"),
            features=r, thresholds=rt, xsha=np.array(sha(X)), n=np.array(n),
            nfeatures=np.array(fruit.nfeatures()), labels=labels,
            summary=np.array(fruit.summary()))


def gen_corbeille():
    """Caller side of the path: two small datasets
    in the UCR .txt layout (committed under tests/golden/ucr/), what the
    REFERENCE's ``corbeille.data.load`` makes of them, and the features /
    accuracy of the steps of ``corbeille.fruitify``
    (experiments/corbeille/corbeille/fruitifier.py:50-72) run with the reference's
    Fruit.  (fruitifier.py itself imports matplotlib through fruitalyser, which
    this image lacks; data.py imports on its own.)"""
    import importlib.util
    print("[corbeille]")
    spec_ = importlib.util.spec_from_file_location(
        "ref_corbeille_data", os.path.join(REF, "experiments", "corbeille", "corbeille", "data.py"))
    rdata = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(rdata)
    root = os.path.join(GOLD, "ucr")
    rng = np.random.default_rng(21)
    for name, comma, nan in (("Delta", False, False), ("Eps", True, True)):
        os.makedirs(os.path.join(root, name), exist_ok=True)
        for split, n in (("TRAIN", 36), ("TEST", 24)):
            y = rng.integers(1, 3, size=n)
            X = (0.4 * rng.standard_normal((n, 48)).cumsum(axis=1)
                 + (y[:, None] == 2) * 2.5 * np.sin(np.linspace(0, 3 * np.pi, 48)))
            if nan:
                X[0, 0] = np.nan
                X[1, 5:8] = np.nan
                X[2, -1] = np.nan
            raw = np.concatenate([y[:, None].astype(float), np.round(X, 6)], axis=1)
            np.savetxt(os.path.join(root, name, f"{name}_{split}.txt"), raw, fmt="%.6f",
                       delimiter="," if comma else "  ")
    out = {}
    for name in ("Delta", "Eps"):
        Xtr, ytr, Xte, yte = rdata.load(os.path.join(root, name))
        out.update({f"{name}_X_train": Xtr, f"{name}_y_train": ytr, f"{name}_X_test": Xte,
                    f"{name}_y_test": yte})
        kept = rdata.load(os.path.join(root, name), keep_nan=True)[0]
        out[f"{name}_X_train_keep_nan"] = kept
        # the steps of fruitify with the reference's Fruit and default classifier
        from sklearn.linear_model import RidgeClassifierCV
        from sklearn.pipeline import Pipeline
        from sklearn.preprocessing import FunctionTransformer, StandardScaler
        fruit = specs.build_fruit(ref, specs.SPECS["C2_reduced"])
        A, B = np.nan_to_num(Xtr), np.nan_to_num(Xte)
        np.random.seed(0)
        fruit.fit(A)
        ftr, fte = fruit.transform(A), fruit.transform(B)
        clf = Pipeline(steps=[("scaler", StandardScaler()),
                              ("nantonum", FunctionTransformer(np.nan_to_num)),
                              ("ridge", RidgeClassifierCV(alphas=np.logspace(-3, 3, 10)))])
        clf.fit(ftr, ytr)
        out[f"{name}_features_train"], out[f"{name}_features_test"] = ftr, fte
        out[f"{name}_accuracy"] = np.array(clf.score(fte, yte))
        print(f"  {name}: train {Xtr.shape}, test {Xte.shape}, accuracy {float(out[name + '_accuracy']):.3f}")
    np.savez_compressed(os.path.join(GOLD, "corbeille.npz"), **out)


def _reference_corbeille():
    """The reference's whole ``corbeille`` package.  Its ``__init__`` pulls in the
    plotting class, i.e. matplotlib, which this image lacks: empty stand-ins for
    the matplotlib names it imports are enough (nothing here plots)."""
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.axes", "matplotlib.figure"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].cm = object()
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.axes"].Axes = object
    sys.modules["matplotlib.figure"].Figure = object
    sys.path.insert(0, os.path.join(REF, "experiments", "corbeille"))
    import corbeille as rc
    assert os.path.realpath(rc.__file__).startswith(os.path.realpath(REF)), rc.__file__
    return rc


def gen_corbeille2():
    """The rest of the harness: the multivariate .arff reader (fixture committed
    under tests/golden/ucr_mv/), multisine and the resampling helpers under seeds,
    tools.split_index, decide_which_fruit."""
    print("[corbeille 2]")
    rc = _reference_corbeille()
    out = {}
    root = os.path.join(GOLD, "ucr_mv", "Zeta")
    os.makedirs(root, exist_ok=True)
    rng = np.random.default_rng(3)

    def write(path, n, labels):
        d, t = 2, 6
        with open(path, "w") as f:
            f.write("@relation Zeta\n@attribute relationalAtt relational\n")
            for k in range(t):
                f.write(f"  @attribute att{k} numeric\n")
            f.write("@end relationalAtt\n@attribute classAttribute {up,down,flat}\n@data\n")
            for i in range(n):
                X = np.round(rng.standard_normal((d, t)), 3)
                if i == 1:
                    X[0, 2] = np.nan
                rows = "\\n".join(",".join("?" if np.isnan(v) else repr(float(v)) for v in row)
                                   for row in X)
                f.write(f"'{rows}',{labels[i % len(labels)]}\n")

    write(os.path.join(root, "Zeta_TRAIN.arff"), 5, ["down", "up", "down", "flat"])
    write(os.path.join(root, "Zeta_TEST.arff"), 4, ["flat", "up"])
    for keep in (False, True):
        got = rc.data.load(root, univariate=False, cache=False, keep_nan=keep)
        for key, a in zip(("X_train", "y_train", "X_test", "y_test"), got):
            out[f"arff_{key}" + ("_keep_nan" if keep else "")] = a
    X = np.random.default_rng(0).standard_normal((4, 2, 30)).cumsum(axis=2)
    out["util_X"] = X
    out["lengthen"] = rc.data.lengthen(X, 0.2)
    out["downsample"] = rc.data.downsample(X, 0.34)
    out["upsample"] = rc.data.upsample(X)
    out["upsample_1d"] = rc.data.upsample(X[:, :1])
    for tag, sl in (("a", 0.1), ("b", 0.5)):
        np.random.seed(5)
        out["stutter_" + tag] = rc.data.implant_stuttering(X, sl)
        out["stutter_" + tag + "_rng"] = np.array(np.random.random())
    np.random.seed(9)
    ms = rc.data.multisine(train_size=11, test_size=7, length=20, n_classes=3)
    out["multisine_rng"] = np.array(np.random.random())
    for key, a in zip(("X_train", "y_train", "X_test", "y_test"), ms):
        out["multisine_" + key] = a
    # tools.split_index over every index of a fruit with chained ISS and wide sieves
    fruit = specs.build_fruit(ref, specs.SPECS["R_mixed"])
    for level, count in (("prepared", len(fruit)),
                         ("iterated sums", sum(int(np.prod([i.n_iterated_sums() for i in s.get_iss()]))
                                               for s in fruit)),
                         ("features", fruit.nfeatures())):
        rows = [rc.tools.split_index(fruit, i, level) for i in range(count)]
        out["split_" + level.replace(" ", "_")] = np.array(rows, dtype=np.int64)
    # decide_which_fruit: which of three candidates the reference picks on a dataset
    Xtr, ytr, Xte, yte = (np.load(os.path.join(GOLD, "corbeille.npz"))[f"Delta_{k}"]
                          for k in ("X_train", "y_train", "X_test", "y_test"))
    choices = [specs.build_fruit(ref, specs.SPECS[n]) for n in ("R_decide_a", "R_decide_b")]
    choices.append((choices[0], specs.build_fruit(ref, specs.SPECS["C2_reduced"])))
    picked = []
    for seed in (0, 1, 2):
        np.random.seed(seed)
        chosen = rc.fruitifier.decide_which_fruit(choices, n_splits=2)(np.nan_to_num(Xtr), ytr)
        picked.append(chosen.nfeatures())
    out["decide_nfeatures"] = np.array(picked)
    lonely = ytr.copy()
    lonely[0] = 7                                   # a class with a single sample
    out["decide_lonely"] = np.array(rc.fruitifier.decide_which_fruit(choices)(Xtr, lonely).nfeatures())
    print("  decide_which_fruit picked fruits with", picked, "features")
    np.savez_compressed(os.path.join(GOLD, "corbeille2.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["words", "iss", "sieves", "preps", "pipelines"]
    for w in which:
        globals()["gen_" + w]()
    print("golden vectors written to", GOLD)
