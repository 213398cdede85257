"""CPU tests: the oracle against the frozen outputs of the real reference
(tests/golden/*.npz, written by oracle/gen_golden.py) and against the
reference's own hand-computed known-answer vectors."""
import hashlib
import os

import numpy as np
import pytest

import specs
from cases import (ISS_CASES, PIPE_CASES, PREP_CASES, SIEVE_CASES, SUMMING_SIEVES, IMPLICIT_SIEVES, sieve_kind, unwrap, KAT_X, SIEVE_KATS, make_iss_input,
                   make_prep_input, make_sieve_input)
from helpers import assert_close, assert_exact, oracle_thresholds
from oracle import pipeline as orc

# fixed array of the reference's tests (tests/signature/test_simple.py:5-8)
X_1 = np.array([
    [[-4, 0.8, 0, 5, -3], [2.0, 1, 0, 0, -7]],
    [[5.0, 8, 2, 6, 0], [-5, -1, -4, -0.5, -8]],
])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def test_words_and_plans(golden_dir):
    g = np.load(os.path.join(golden_dir, "words.npz"))
    for key in g.files:
        if key.startswith("of_weight_"):
            _, _, w, d = key.split("_")
            words = orc.of_weight(int(w), int(d))
            assert "|".join(words) == str(g[key])
            assert orc.cache_plan(words) == list(g[f"plan_{w}_{d}"])
    # reference golden: tests/signature/test_cache.py:11-26
    words = ["[1][11][3][11]", "[11][13][11][1][3]", "[1][13][1]",
             "[11][13][111][13][11]", "[3][11][111]", "[1][11][2]", "[11][2]",
             "[11][13][111][13][2]", "[3][11][1112][21]"]
    assert orc.cache_plan(words) == [4, 5, 2, 3, 3, 1, 1, 1, 2]


def test_reference_kat_reals():
    # tests/signature/test_simple.py:11-33
    words = ["[1]", "[2]", "[11]", "[12]", "[1][1]", "[1][2]"]
    correct = (
        np.array([[-4, -3.2, -3.2, 1.8, -1.2], [5, 13, 15, 21, 21]]),
        np.array([[2, 3, 3, 3, -4], [-5, -6, -10, -10.5, -18.5]]),
        np.array([[16, 16.64, 16.64, 41.64, 50.64], [25, 89, 93, 129, 129]]),
        np.array([[-8, -7.2, -7.2, -7.2, 13.8], [-25, -33, -41, -44, -44]]),
        np.array([[0, -3.2, -3.2, -19.2, -24.6], [0, 40, 66, 156, 156]]),
        np.array([[0., -4., -4., -4., -16.6], [0, -5, -57, -64.5, -232.5]]),
    )
    res = list(orc.iss_iter(X_1, {"words": words}, orc.RawCache(X_1)))
    for r, c in zip(res, correct):
        np.testing.assert_allclose(c, r)


def test_reference_kat_arctic():
    # tests/signature/test_semiring.py:10-33
    words = ["[1]", "[2]", "[11]", "[12]", "[1][1]", "[1][2]"]
    res = list(orc.iss_iter(X_1, {"words": words, "semiring": "arctic"}, orc.RawCache(X_1)))
    correct = (
        np.array([[-4, 0.8, 0.8, 5, 5], [5, 8, 8, 8, 8]]),
        np.array([[2, 2, 2, 2, 2], [-5, -1, -1, -0.5, -0.5]]),
        np.array([[-8, 1.6, 1.6, 10, 10], [10, 16, 16, 16, 16]]),
        np.array([[-2, 1.8, 1.8, 5, 5], [0, 7, 7, 7, 7]]),
        np.array([[-8, 1.6, 1.6, 10, 10], [10, 16, 16, 16, 16]]),
        np.array([[-2, 1.8, 1.8, 5., 5.], [0., 7., 7., 7.5, 7.5]]),
    )
    for r, c in zip(res, correct):
        np.testing.assert_allclose(c, r)


def test_reference_kat_features():
    # tests/core/test_branches.py:61-86
    spec = {"slices": [
        {"iss": [{"words": ["[1]", "[2]", "[11]"]}], "sieves": [["MAX", {}]]},
        {"iss": [{"words": ["[12]", "[1][1]", "[1][2]"]}], "sieves": [["MIN", {}]]},
    ]}
    of = orc.OracleFruit(spec)
    feats = of.fit_transform(X_1)
    np.testing.assert_allclose(np.array([
        [1.8, 3., 50.64, -8., -24.6, -16.6],
        [21, -5, 129, -44, 0, -232.5]]), feats)


@pytest.mark.parametrize("name", sorted(ISS_CASES))
def test_iss_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "iss.npz"))
    desc, shape, kind = ISS_CASES[name]
    X = make_iss_input(shape, kind)
    assert sha(X) == str(g[name + "_xsha"]), "seeded input changed"
    res = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
    if desc.get("weighting") is None:
        assert_exact(res, g[name], name)
    else:
        # exp() comes from the host libm: identical in the image that froze
        # the vectors, within an ulp elsewhere
        assert_close(res, g[name], 1e-12, name)


@pytest.mark.parametrize("name", sorted(SIEVE_CASES))
def test_sieve_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "sieves.npz"))
    raw, Y = make_sieve_input()
    assert sha(Y) == str(g["Y_xsha"])
    sv = orc.make_sieve(SIEVE_CASES[name])
    np.random.seed(3)
    sv.fit(Y)
    res = sv.transform(Y, orc.RawCache(raw))
    kind = sieve_kind(SIEVE_CASES[name])
    thr = unwrap(sv).fitted_q if kind in IMPLICIT_SIEVES else unwrap(sv).quantiles
    assert_exact(np.array(thr, dtype=np.float64), g[name + "_thr"], name + " thresholds")
    if kind in SUMMING_SIEVES:
        assert_close(res, g[name], 1e-12, name)
    else:
        assert_exact(res, g[name], name)


@pytest.mark.parametrize("name", sorted(PREP_CASES))
def test_prep_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "preps.npz"))
    X = make_prep_input()
    assert sha(X) == str(g["xsha"])
    assert_exact(orc.apply_prep(PREP_CASES[name], X), g[name], name)


@pytest.mark.parametrize("name", sorted(__import__("cases").LETTER_CASES))
def test_generic_words_golden(name, golden_dir):
    """Words over Python letters in the three semirings (oracle restatement of
    Semiring._iterated_sum, fruits/iss/semiring.py:54-75, :428-446) against the
    reference's frozen ISS output, bit for bit."""
    from cases import LETTER_CASES, make_iss_input
    g = np.load(os.path.join(golden_dir, "letters.npz"))
    desc, shape, kind = LETTER_CASES[name]
    X = make_iss_input(shape, kind)
    assert sha(X) == str(g[name + "_xsha"])
    assert_exact(np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X)))), g[name], name)


@pytest.mark.parametrize("name", sorted(__import__("cases").ARGMAX_CASES))
def test_arctic_argmax_golden(name, golden_dir):
    """Arctic(argmax=True): maxima and their positions (oracle restatement of
    fruits/iss/semiring.py:234-279) against the reference's frozen output."""
    from cases import ARGMAX_CASES, make_iss_input
    g = np.load(os.path.join(golden_dir, "argmax.npz"))
    desc, shape, kind = ARGMAX_CASES[name]
    X = make_iss_input(shape, kind)
    assert sha(X) == str(g[name + "_xsha"])
    res = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
    if desc.get("weighting") is None:
        assert_exact(res, g[name], name)
    else:
        assert_close(res, g[name], 1e-12, name)


@pytest.mark.parametrize("name", sorted(__import__("cases").COS_RANDOM_CASES))
def test_randomised_coswiss_golden(name, golden_dir):
    """CosWISS with a random network in front / random dropout (oracle
    restatement of fruits/iss/cos.py:52-160, :243-260) against the reference's
    frozen output; same draws, same generator state."""
    import copy
    from cases import COS_RANDOM_CASES, make_iss_input
    g = np.load(os.path.join(golden_dir, "cos2.npz"))
    desc, shape, kind = COS_RANDOM_CASES[name]
    X = make_iss_input(shape, kind)
    assert sha(X) == str(g[name + "_xsha"])
    desc = copy.deepcopy(desc)
    np.random.seed(3)
    desc["_state"] = orc.coswiss_fit(desc, X)
    assert np.random.random() == float(g[name + "_rng"])
    res = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
    assert_close(res, g[name], 1e-11, name)


@pytest.mark.parametrize("name", sorted(__import__("cases").PREP2_CASES))
def test_prep2_golden(name, golden_dir):
    """The preparateurs beside INC / STD / NRM: the numpy restatement
    (oracle/preps.py) against the outputs frozen from the reference, fit state
    reused on a second batch, same consumption of the global RNG."""
    from cases import PREP2_CASES, PREP2_EXACT, make_prep2_inputs
    from oracle import preps as more
    g = np.load(os.path.join(golden_dir, "preps2.npz"))
    desc = PREP2_CASES[name]
    X, X2 = make_prep2_inputs(name)
    np.random.seed(7)
    st = more.fit_prep(desc, X)
    assert np.random.random() == float(g[name + "_rng"])
    with np.errstate(invalid="ignore"):
        res = more.transform_prep(desc, st, X, orc.RawCache(X))
        res2 = more.transform_prep(desc, st, X2, orc.RawCache(X2))
    for got, key in ((res, name), (res2, name + "_2")):
        if desc[0] in PREP2_EXACT:
            assert_exact(got, g[key], key)
        else:
            assert_close(got, g[key], 1e-12, key)


@pytest.mark.parametrize("name", sorted(PIPE_CASES) + ["C2_cos", "R_mixed", "R_rng", "R_preps", "R_letters", "R_cosrand", "R_argmax"])
def test_pipeline_golden(name, golden_dir):
    from cases import COS_PIPE_CASES, EXTRA_PIPE_CASES
    g = np.load(os.path.join(golden_dir, f"pipeline_{name}.npz"))
    spec_name, n = {**PIPE_CASES, **COS_PIPE_CASES, **EXTRA_PIPE_CASES}[name]
    spec = specs.SPECS[spec_name]
    X = specs.make_input(spec_name, n)
    assert sha(X) == str(g["xsha"])
    of = orc.OracleFruit(spec)
    np.random.seed(0)
    of.fit(X)
    res = of.transform(X)
    assert of.nfeatures() == int(g["nfeatures"])
    if name in ("C1_readme", "C5_sweep"):
        assert_exact(oracle_thresholds(of), g["thresholds"], name + " thresholds")
        assert_exact(res, g["features"], name + " features")
    else:
        assert_close(oracle_thresholds(of), g["thresholds"], 1e-11, name + " thresholds")
        scale = np.maximum(np.abs(g["features"]), 1.0)
        bad = np.abs(res - g["features"]) > 1e-9 * scale
        assert bad.mean() < 1e-3, f"{bad.sum()} features beyond 1e-9"


@pytest.mark.parametrize("name", sorted(__import__("cases").COS_CASES))
def test_oracle_coswiss_matches_reference(name, golden_dir):
    """Cosine weighted ISS (fruits/iss/cos.py) of the oracle vs the frozen reference."""
    from cases import COS_CASES, make_iss_input
    g = np.load(os.path.join(golden_dir, "cos.npz"))
    desc, shape, kind = COS_CASES[name]
    X = make_iss_input(shape, kind)
    o = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
    ref = g[name]
    scale = np.maximum(np.abs(ref), np.abs(ref).max(axis=-1, keepdims=True))
    assert o.shape == ref.shape and np.all(np.abs(o - ref) <= 1e-11 * scale)
    labels = "|".join(orc.iss_label(desc, i) for i in range(ref.shape[0]))
    assert labels == str(g[name + "_labels"])


def test_coswiss_expansion_table():
    # cos(a-b)^2 = sin^2 a sin^2 b + 2 sin a cos a sin b cos b + cos^2 a cos^2 b
    w = orc.coswiss_weightings(2, 2, False)
    assert w.tolist() == [[1, 2, 0, 2, 0], [2, 1, 1, 1, 1], [1, 0, 2, 0, 2]]
    assert orc.coswiss_weightings(3, 2, True).shape == (27, 9)
    import fruits_b200 as fruits
    for n_letters, e, total in [(1, 1, True), (2, 3, False), (3, 2, True), (2, 4, True)]:
        word = fruits.words.SimpleWord("[1]" * n_letters)
        host = fruits.CosWISS([word], freqs=[0.1], exponent=e,
                              total_weighting=total)._get_weightings(word)
        assert np.array_equal(host, orc.coswiss_weightings(n_letters, e, total))


@pytest.mark.parametrize("kat", range(len(SIEVE_KATS)))
def test_reference_kat_sieve_table(kat):
    """The oracle reproduces the known answers of the reference's own sieve
    tests (tests/sieving/test_explicit.py, test_implicit.py; data in cases.py).
    A sieve applied stand-alone builds its cache from its own input
    (fruits/seed.py:26-51), so coquantile cuts refer to the sieved rows."""
    desc, block, want = SIEVE_KATS[kat]
    Y = KAT_X[block]
    sv = orc.make_sieve(desc)
    np.random.seed(0)
    sv.fit(Y)
    got = sv.transform(Y, orc.RawCache(Y[:, np.newaxis, :]))
    np.testing.assert_allclose(got, np.array(want, dtype=float), rtol=1e-12, atol=1e-15)


def _indices_lookup(t, scale=50):
    # fruits/iss/weighting.py:100-110: NRM(arange(1..T)/T) * scale
    base = np.arange(1, t + 1) / t
    return (base - base.min()) / (base.max() - base.min()) * scale


@pytest.mark.parametrize("semiring", ["reals", "arctic"])
def test_two_letter_word_against_brute_force(semiring):
    """Definition check independent of any implementation (the reference does
    the same in tests/signature/test_weighting.py): the iterated sum of
    ``[1][2]`` is a double sum / maximum over index pairs, with the weighting
    ``exp(alpha (g(i) - g(j)))`` (reals) or ``+ alpha (g(i) - g(j))`` (arctic)
    between the two indices."""
    rng = np.random.default_rng(8)
    n, t = 3, 23
    X = rng.standard_normal((n, 2, t))
    for weighting in (None, ["Indices", {"scale": 3}]):
        desc = {"words": ["[1][2]"], "mode": "single", "semiring": semiring,
                "weighting": weighting, "alphas": [[0.7, 1.0]] if weighting else None}
        got = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))[0]
        g = _indices_lookup(t, 3) if weighting else np.zeros(t)
        a0 = np.float64(np.float32(0.7)) if weighting else 0.0
        want = np.zeros((n, t))
        for s in range(n):
            for T in range(t):
                if semiring == "reals":
                    want[s, T] = sum(X[s, 0, i] * X[s, 1, j] * np.exp(a0 * (g[i] - g[j]))
                                     for j in range(T + 1) for i in range(j))
                else:
                    want[s, T] = max(X[s, 0, i] + X[s, 1, j] + a0 * (g[i] - g[j])
                                     for j in range(T + 1) for i in range(j + 1))
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-12)


def test_full_size_goldens_describe_the_baseline_inputs(golden_dir):
    """tests/golden/full_*.npz (oracle/gen_golden_full.py): frozen from the real
    reference at BASELINE size; the GPU tests compare against them.  Here: the
    files belong to the seeded inputs of tests/specs.py and are self-consistent
    (the generator checked the oracle against them when it wrote them)."""
    import hashlib
    import glob
    paths = sorted(glob.glob(os.path.join(golden_dir, "full_*.npz")))
    assert any(p.endswith("full_C2_full.npz") for p in paths)
    for path in paths:
        g = np.load(path)
        name = os.path.basename(path)[len("full_"):-len(".npz")]
        X = specs.make_input(name)
        sha = hashlib.sha256(np.ascontiguousarray(X).tobytes()).hexdigest()[:16]
        assert sha == str(g["xsha"]) and X.shape[0] == int(g["n"])
        assert str(g["source"]) == "reference"
        nslices = len(specs.SPECS[name]["slices"])
        assert g["thr_slices"].shape == (nslices + 1,) and g["thr_slices"][-1] == g["thresholds"].size
        assert g["features"].shape == (g["rows"].size, int(g["nfeatures"]))
        assert np.isfinite(g["features"]).all()
