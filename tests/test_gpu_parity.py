"""GPU parity tests (run on the B200 box): the CUDA path, called through the
C ABI of libfruits_b200.so via the Python mirror of the reference API, against
(a) the frozen outputs of the real reference (tests/golden), (b) the CPU
oracle on other seeded inputs, (c) the reference's hand-computed vectors and
(d) size-independent properties at larger sizes.

Bars: bit-exact for the arctic semiring, the unweighted real semiring, word
enumeration, thresholds and integer-valued sieves; |a-b| <= 1e-9 *
max(|b|, rowmax|b|) for exponentially weighted real iterated sums (exp() on
the device differs from the host libm in the last ulp); 1e-12 for MPI means
(the reference's own summation order is unspecified under numba fastmath).
"""
import os

import numpy as np
import pytest
import torch

import fruits_b200 as fruits
import specs
from cases import (ISS_CASES, PIPE_CASES, PREP_CASES, SIEVE_CASES, SUMMING_SIEVES, IMPLICIT_SIEVES, sieve_kind, unwrap, KAT_X, SIEVE_KATS, make_iss_input,
                   make_prep_input, make_sieve_input)
from helpers import (assert_close, assert_exact, fitted_thresholds, oracle_thresholds,
                     parity_report, record_report)

pytestmark = pytest.mark.gpu

X_1 = np.array([
    [[-4, 0.8, 0, 5, -3], [2.0, 1, 0, 0, -7]],
    [[5.0, 8, 2, 6, 0], [-5, -1, -4, -0.5, -8]],
])


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "these tests need the B200"
    from fruits_b200 import _backend as be
    be.lib()   # fails loudly if the CUDA library is missing


def _exact_iss(desc):
    # (weighted Bayesian sums multiply by exp(): device vs host libm)
    return desc.get("weighting") is None or desc.get("semiring") == "arctic"


# ---------------------------------------------------------------------------
# (a) frozen reference outputs

@pytest.mark.parametrize("name", sorted(ISS_CASES))
def test_iss_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "iss.npz"))
    desc, shape, kind = ISS_CASES[name]
    X = make_iss_input(shape, kind)
    res = specs.build_iss(fruits, desc).transform(X)
    if _exact_iss(desc):
        assert_exact(res, g[name], name)
    else:
        assert_close(res, g[name], 1e-9, name)


@pytest.mark.parametrize("name", sorted(n for n in ISS_CASES if n.startswith("arctic")))
@pytest.mark.parametrize("length", [None, 300, 1000])
def test_arctic_block_scan_over_time(name, length, golden_dir, monkeypatch):
    """The Arctic semiring through the block scan over T (``fb_arctic_word``: warp
    shuffles + warp totals in shared memory + carry between tiles; a running
    maximum is exactly associative): bit-identical to the time-serial
    lane-per-node kernel on every arctic case -- unweighted, Indices and L1
    weighted, total and not -- for lengths below, at and across the 256-step
    tile, and to the goldens frozen from the reference."""
    g = np.load(os.path.join(golden_dir, "iss.npz"))
    desc, shape, kind = ISS_CASES[name]
    if length is not None:
        shape = (shape[0], shape[1], length)
    X = make_iss_input(shape, kind)
    monkeypatch.setenv("FRUITS_B200_ARCTIC_SCAN", "0")
    serial = specs.build_iss(fruits, desc).transform(X)
    monkeypatch.setenv("FRUITS_B200_ARCTIC_SCAN", "1")
    scan = specs.build_iss(fruits, desc).transform(X)
    assert_exact(scan, serial, name + " block scan vs time-serial kernel")
    if length is None:
        if _exact_iss(desc):
            assert_exact(scan, g[name], name)
        else:
            assert_close(scan, g[name], 1e-9, name)


@pytest.mark.parametrize("name", sorted(SIEVE_CASES))
def test_sieve_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "sieves.npz"))
    raw, Y = make_sieve_input()
    sv = specs._sieve(fruits, SIEVE_CASES[name])
    sv._cache = fruits.cache.SharedSeedCache(raw)
    np.random.seed(3)
    sv.fit(Y)
    res = sv.transform(Y)
    kind = sieve_kind(SIEVE_CASES[name])
    thr = unwrap(sv)._q if kind in IMPLICIT_SIEVES else unwrap(sv)._quantiles
    assert_exact(np.array(thr, dtype=np.float64), g[name + "_thr"], name + " thresholds")
    if kind in SUMMING_SIEVES:
        assert_close(res, g[name], 1e-12, name)
    else:
        assert_exact(res, g[name], name)


@pytest.mark.parametrize("name", sorted(n for n in PREP_CASES if n != "nrm"))
def test_prep_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "preps.npz"))
    X = make_prep_input()
    p = specs._prep(fruits, PREP_CASES[name])
    p.fit(X)
    assert_exact(p.transform(X), g[name], name)


def test_nrm_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "preps.npz"))
    assert_exact(fruits.preparation.NRM().transform(make_prep_input()), g["nrm"], "nrm")


@pytest.mark.parametrize("name", sorted(PIPE_CASES))
def test_pipeline_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, f"pipeline_{name}.npz"))
    spec_name, n = PIPE_CASES[name]
    X = specs.make_input(spec_name, n)
    fruit = specs.build_fruit(fruits, specs.SPECS[spec_name])
    np.random.seed(0)
    fruit.fit(X)
    res = fruit.transform(X)
    assert res.shape == g["features"].shape and res.dtype == np.float64
    if name in ("C1_readme", "C5_sweep"):
        assert_exact(fitted_thresholds(fruit), g["thresholds"], name + " thresholds")
        assert_exact(res, g["features"], name + " features")
    else:
        assert_close(fitted_thresholds(fruit), g["thresholds"], 1e-9, name + " thresholds")
        _assert_features_close(res, g["features"], name)


def _assert_features_close(res, ref, what):
    """Weighted slices: values within 1e-9; an integer count may move by one
    when an increment sits within rounding of its threshold.  The numbers of
    rtol violations and count flips are reported (helpers.parity_report)."""
    rep = parity_report(res, ref)
    record_report(what, rep)
    scale = np.maximum(np.abs(ref), 1.0)
    bad = np.abs(res - ref) > 1e-9 * scale
    assert bad.mean() <= 2e-4, f"{what}: {bad.sum()} of {bad.size} features beyond 1e-9"
    assert np.all(np.abs(res - ref)[bad] <= 1.0 + 1e-9), what


# ---------------------------------------------------------------------------
# (c) the reference's hand-computed vectors

def test_reference_kat_reals():
    # reference: tests/signature/test_simple.py:11-41
    words = [fruits.words.SimpleWord(s) for s in
             ["[1]", "[2]", "[11]", "[12]", "[1][1]", "[1][2]"]]
    correct = (
        np.array([[-4, -3.2, -3.2, 1.8, -1.2], [5, 13, 15, 21, 21]]),
        np.array([[2, 3, 3, 3, -4], [-5, -6, -10, -10.5, -18.5]]),
        np.array([[16, 16.64, 16.64, 41.64, 50.64], [25, 89, 93, 129, 129]]),
        np.array([[-8, -7.2, -7.2, -7.2, 13.8], [-25, -33, -41, -44, -44]]),
        np.array([[0, -3.2, -3.2, -19.2, -24.6], [0, 40, 66, 156, 156]]),
        np.array([[0., -4., -4., -4., -16.6], [0, -5, -57, -64.5, -232.5]]),
    )
    results = fruits.ISS(words).batch_transform(X_1, batch_size=1)
    for i, result in enumerate(results):
        np.testing.assert_allclose(correct[i], result[0, :, :])
    np.testing.assert_allclose(correct[0],
                               fruits.ISS([words[0].copy()]).fit_transform(X_1)[0])


def test_reference_kat_arctic():
    # reference: tests/signature/test_semiring.py:10-33
    words = [fruits.words.SimpleWord(s) for s in
             ["[1]", "[2]", "[11]", "[12]", "[1][1]", "[1][2]"]]
    correct = (
        np.array([[-4, 0.8, 0.8, 5, 5], [5, 8, 8, 8, 8]]),
        np.array([[2, 2, 2, 2, 2], [-5, -1, -1, -0.5, -0.5]]),
        np.array([[-8, 1.6, 1.6, 10, 10], [10, 16, 16, 16, 16]]),
        np.array([[-2, 1.8, 1.8, 5, 5], [0, 7, 7, 7, 7]]),
        np.array([[-8, 1.6, 1.6, 10, 10], [10, 16, 16, 16, 16]]),
        np.array([[-2, 1.8, 1.8, 5., 5.], [0., 7., 7., 7.5, 7.5]]),
    )
    results = fruits.ISS(words, semiring=fruits.iss.semiring.Arctic()).batch_transform(
        X_1, batch_size=1)
    for i, result in enumerate(results):
        np.testing.assert_allclose(correct[i], result[0])


def test_reference_kat_two_slices():
    # reference: tests/core/test_branches.py:61-86
    fruit = fruits.Fruit()
    w = [fruits.words.SimpleWord(s) for s in
         ["[1]", "[2]", "[11]", "[12]", "[1][1]", "[1][2]"]]
    fruit.add(fruits.ISS(w[:3]))
    fruit.add(fruits.sieving.MAX)
    fruit.cut()
    fruit.add(fruits.ISS(w[3:]))
    fruit.add(fruits.sieving.MIN)
    assert fruit.nfeatures() == 6
    features = fruit.fit_transform(X_1)
    np.testing.assert_allclose(np.array([
        [1.8, 3., 50.64, -8., -24.6, -16.6],
        [21, -5, 129, -44, 0, -232.5]]), features)


def test_reference_kat_sieves():
    # reference: tests/sieving/test_explicit.py (MAX/MIN/END/NPI on X_1[:, 0, :])
    Y = X_1[:, 0, :]
    np.testing.assert_allclose(fruits.sieving.MAX().fit_transform(Y), [[5], [8]])
    np.testing.assert_allclose(fruits.sieving.MIN().fit_transform(Y), [[-4], [0]])
    np.testing.assert_allclose(fruits.sieving.END().fit_transform(Y), [[-3], [0]])
    np.testing.assert_allclose(fruits.sieving.NPI().fit_transform(Y), [[2], [2]])
    np.testing.assert_allclose(fruits.sieving.END(cut=[1, 3, -1]).fit_transform(Y),
                               [[-4, 0, -3], [5, 2, 0]])


def test_theoretical_identity():
    # reference: tests/signature/test_simple.py:44-51: <[1][1], ISS>_T = -T/2
    X = np.random.default_rng(3).random((25, 1, 100))
    X = (X - X.mean(axis=-1, keepdims=True)) / X.std(axis=-1, keepdims=True)
    res = fruits.ISS([fruits.words.SimpleWord("[1][1]")]).fit_transform(X)
    np.testing.assert_allclose(np.ones((25,)) * -50, res[0, :, -1])


# ---------------------------------------------------------------------------
# (b) oracle on other seeded inputs, edge cases

def _oracle_iss(X, desc):
    from oracle import pipeline as orc
    return np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))


@pytest.mark.parametrize("shape", [(1, 3, 1), (2, 3, 2), (3, 3, 31), (3, 3, 32), (2, 3, 33),
                                   (5, 3, 257), (1, 3, 1000), (130, 3, 64), (2, 3, 5000)])
@pytest.mark.parametrize("semiring", ["reals", "arctic"])
def test_iss_shapes_vs_oracle(shape, semiring):
    desc = {"words": {"of_weight": [3, 3]}, "mode": "extended", "semiring": semiring}
    X = np.random.default_rng(shape[2]).standard_normal(shape)
    assert_exact(specs.build_iss(fruits, desc).transform(X), _oracle_iss(X, desc), str(shape))


def test_iss_many_blocks_vs_oracle():
    # 1,351 nodes: several kernel blocks with duplicated ancestors
    desc = {"words": {"of_weight": [6, 2]}, "mode": "extended"}
    X = np.random.default_rng(1).standard_normal((3, 2, 96)) * 0.7
    assert_exact(specs.build_iss(fruits, desc).transform(X), _oracle_iss(X, desc), "w6d2")


def test_iss_long_letter_and_deep_word():
    desc = {"words": ["[111111111]", "[11111111111][2]", 60 * "[1]"], "mode": "extended"}
    X = 0.9 + 0.2 * np.random.default_rng(2).random((2, 2, 70))
    assert_exact(specs.build_iss(fruits, desc).transform(X), _oracle_iss(X, desc), "long")
    desc = {"words": {"alternate_sign": [100 * "[1]"]}, "mode": "extended",
            "semiring": "arctic"}
    assert_exact(specs.build_iss(fruits, desc).transform(X), _oracle_iss(X, desc), "deep")


def test_iss_custom_alpha_vs_oracle():
    desc = {"words": ["[1][2][1]", "[1][2][2]", "[2][11]"], "mode": "extended",
            "weighting": ["Indices", {"scale": 3}],
            "alphas": [[0.5, 1.0, 2.0], [0.5, 0.25, 1.0], [1.0, 1.0]]}
    X = np.random.default_rng(4).standard_normal((4, 2, 90))
    assert_close(specs.build_iss(fruits, desc).transform(X), _oracle_iss(X, desc), 1e-9, "alpha")
    desc["semiring"] = "arctic"
    assert_exact(specs.build_iss(fruits, desc).transform(X), _oracle_iss(X, desc), "alpha arctic")


def test_extended_equals_stacked_single():
    # reference: tests/signature/test_cache.py:29-80
    X = np.random.default_rng(5).random((10, 3, 100))
    ext = fruits.ISS([fruits.words.SimpleWord("[11][21][331][22]")],
                     mode=fruits.ISSMode.EXTENDED).fit_transform(X)
    single = fruits.ISS([fruits.words.SimpleWord(s) for s in
                         ["[11]", "[11][12]", "[11][12][133]", "[11][12][133][22]"]]
                        ).fit_transform(X)
    assert_exact(ext, single, "extended vs single")


def test_torch_input_stays_on_device():
    X = torch.randn(4, 3, 50, dtype=torch.float64, device="cuda")
    iss = fruits.ISS(fruits.words.of_weight(2, 3))
    res = iss.transform(X)
    assert isinstance(res, torch.Tensor) and res.is_cuda
    assert_exact(res.cpu().numpy(), iss.transform(X.cpu().numpy()), "torch vs numpy")


def test_input_validation():
    iss = fruits.ISS(fruits.words.of_weight(2, 3))
    with pytest.raises(TypeError):
        iss.transform(np.zeros((2, 3, 10), dtype=np.float32))
    with pytest.raises(IndexError):
        iss.transform(np.zeros((2, 2, 10)))
    fruit = specs.build_fruit(fruits, specs.SPECS["C1_readme"])
    with pytest.raises(RuntimeError, match="Missing call of self.fit"):
        fruit.transform(np.zeros((2, 3, 10)))


def test_empty_batch():
    fruit = specs.build_fruit(fruits, specs.SPECS["C5_sweep"])
    np.random.seed(0)
    fruit.fit(specs.make_input("C5_sweep", 4))
    assert fruit.transform(np.zeros((0, 3, 1024))).shape == (0, 2225)


@pytest.mark.parametrize("name,n,seed", [("C1_readme", 37, 1), ("C5_sweep", 70, 2),
                                         ("C2_reduced", 16, 3), ("C4_twi", 5, 4)])
def test_pipeline_vs_oracle_other_seeds(name, n, seed):
    from oracle import pipeline as orc
    spec = specs.SPECS[name]
    shape = {"C1_readme": (n, 3, 77), "C5_sweep": (n, 3, 300), "C2_reduced": (n, 1, 200),
             "C4_twi": (n, 3, 333)}[name]
    X = np.random.default_rng(seed).standard_normal(shape).cumsum(axis=2)
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(seed)
    fruit.fit(X)
    res = fruit.transform(X)
    of = orc.OracleFruit(spec)
    np.random.seed(seed)
    of.fit(X)
    ref = of.transform(X)
    if name in ("C1_readme", "C5_sweep"):
        assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
        assert_exact(res, ref, name)
    else:
        assert_close(fitted_thresholds(fruit), oracle_thresholds(of), 1e-9, "thresholds")
        _assert_features_close(res, ref, name)


def test_transform_on_unseen_data_and_fit_sample_fraction():
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": {"of_weight": [3, 2]}, "mode": "extended"}],
                        "sieves": [["NPI", {"q": [0.3, 1.0]}], ["PPV", {"sample_size": 0.5}],
                                   ["MAX", {"q": [-1.0, 0.7]}], ["END", {}]],
                        "fit_sample_size": 0.4}]}
    rng = np.random.default_rng(9)
    Xtr, Xte = rng.standard_normal((20, 2, 150)), rng.standard_normal((9, 2, 150))
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(11)
    fruit.fit(Xtr)
    after_gpu = np.random.random()
    np.random.seed(11)
    of.fit(Xtr)
    after_cpu = np.random.random()
    assert after_gpu == after_cpu, "global RNG consumed differently from the reference"
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
    # MAX on an empty selection raises in the reference (np.max of an empty
    # array); compare only where the oracle defines a value
    assert_exact(fruit.transform(Xte), of.transform(Xte), "unseen data")


def test_fused_equals_composed_route():
    """The fused kernel and the materialise+sieve route give identical bits."""
    X = specs.make_input("C5_sweep", 24)
    fruit = specs.build_fruit(fruits, specs.SPECS["C5_sweep"])
    np.random.seed(0)
    fruit.fit(X)
    fused = fruit.transform(X)
    composed = fruit.transform(X, callbacks=[fruits.callback.AbstractCallback()])
    assert_exact(fused, composed, "fused vs composed")


def test_general_sieves_composed_route():
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [["INC", {"depth": 2}]],
                        "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended",
                                 "semiring": "arctic"}],
                        "sieves": [["NPI", {"cut": [10, 0.5, -1], "q": [0.2, 0.6, 1.0]}],
                                   ["END", {"cut": [5, -1]}], ["LPI", {}], ["XPI", {}]],
                        "fit_sample_size": 1.0}]}
    X = np.random.default_rng(8).standard_normal((12, 2, 64)).cumsum(axis=2)
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(2)
    fruit.fit(X)
    np.random.seed(2)
    of.fit(X)
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
    assert_exact(fruit.transform(X), of.transform(X), "composed route")


def test_chained_iss():
    # reference: tests/signature/test_consecutive.py -- ISS of ISS in one slice
    from oracle import pipeline as orc
    spec = {"slices": [{"iss": [{"words": ["[1]", "[12]"]}, {"words": ["[1][1]", "[11]"]}],
                        "sieves": [["END", {}], ["NPI", {}]]}]}
    X = np.random.default_rng(6).standard_normal((5, 2, 40))
    fruit = specs.build_fruit(fruits, spec)
    assert fruit.nfeatures() == 8
    assert_exact(fruit.fit_transform(X), orc.OracleFruit(spec).fit_transform(X), "chained")


def test_callbacks_receive_host_arrays():
    seen = {"itsum": 0, "prep": 0, "end": 0}

    class CB(fruits.callback.AbstractCallback):
        def on_iterated_sum(self, X):
            assert isinstance(X, np.ndarray) and X.shape == (6, 100)
            seen["itsum"] += 1

        def on_preparateur(self, X):
            seen["prep"] += 1

        def on_sieving_end(self, X):
            seen["end"] += 1

    X = specs.make_input("C1_readme", 6)
    fruit = specs.build_fruit(fruits, specs.SPECS["C1_readme"])
    fruit.fit(X)
    fruit.transform(X, callbacks=[CB()])
    assert seen == {"itsum": 36, "prep": 1, "end": 2}


# ---------------------------------------------------------------------------
# (d) size-independent properties at larger sizes

def test_sweep_large_batch_properties():
    """C5 pipeline on 20,000 series: a subset equals the oracle bit for bit,
    rows are independent (permutation equivariance) and runs are idempotent."""
    from oracle import pipeline as orc
    spec = specs.SPECS["C5_sweep"]
    rng = np.random.default_rng(77)
    X = rng.standard_normal((20000, 3, 1024))
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(5)
    fruit.fit(X)
    res = fruit.transform(X)
    assert res.shape == (20000, 2225) and np.isfinite(res).all()
    assert_exact(fruit.transform(X), res, "idempotence")
    perm = rng.permutation(20000)[:4000]
    assert_exact(fruit.transform(X[perm]), res[perm], "row independence")
    pick = np.sort(rng.choice(20000, 48, replace=False))
    of = orc.OracleFruit(spec)
    np.random.seed(5)
    of.fit(X)          # same RNG draw -> same fit row
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
    assert_exact(res[pick], of.transform(X[pick]), "subset vs oracle")
    # NPI counts are integers in [0, T]; PPV in [0, 1]; MIN <= END <= MAX
    f = res.reshape(20000, 445, 5)
    assert np.all(f[..., 0] == np.round(f[..., 0])) and f[..., 0].min() >= 0
    assert f[..., 0].max() <= 1024 and 0 <= f[..., 1].min() and f[..., 1].max() <= 1
    assert np.all(f[..., 3] <= f[..., 4]) and np.all(f[..., 4] <= f[..., 2])


def test_twi_full_length_subset():
    from oracle import pipeline as orc
    spec = specs.SPECS["C4_twi"]
    X = specs.make_input("C4_twi", 600)
    fruit = specs.build_fruit(fruits, spec)
    fruit.fit(X)
    res = fruit.transform(X)
    pick = np.arange(0, 600, 100)
    of = orc.OracleFruit(spec)
    of.fit(X[pick])
    ref = of.transform(X[pick])
    # slice 1 (arctic, columns 1533:) is bit-exact, slice 0 is L1-weighted
    assert_exact(res[pick][:, 1533:], ref[:, 1533:], "arctic slice")
    _assert_features_close(res[pick][:, :1533], ref[:, :1533], "weighted slice")


# ---------------------------------------------------------------------------
# (e) the plan-specialised kernel (fruits_b200/_jit.py).  Small batches take
# the generic kernel by default, so these tests force the generated one.

@pytest.fixture
def force_jit(monkeypatch):
    monkeypatch.setenv("FRUITS_B200_JIT", "force")


def _routes(fruit):
    return [getattr(slc, "_last_launch", (None,))[0] for slc in fruit._slices]


@pytest.mark.parametrize("name", sorted(PIPE_CASES))
def test_jit_pipeline_golden(name, golden_dir, force_jit):
    g = np.load(os.path.join(golden_dir, f"pipeline_{name}.npz"))
    spec_name, n = PIPE_CASES[name]
    X = specs.make_input(spec_name, n)
    fruit = specs.build_fruit(fruits, specs.SPECS[spec_name])
    np.random.seed(0)
    fruit.fit(X)
    res = fruit.transform(X)
    routes = _routes(fruit)
    # every slice compiles to a generated kernel: the unweighted arctic slices to
    # the lane-per-node chain kernel, everything else to the thread-per-series one
    arctic = [slc["iss"][0].get("semiring") == "arctic" for slc in specs.SPECS[spec_name]["slices"]]
    want = ["fb_jit_chain" if a else "fb_jit_slice" for a in arctic]
    assert routes == want, routes
    if name in ("C1_readme", "C5_sweep"):
        assert_exact(res, g["features"], name + " features")
    else:
        _assert_features_close(res, g["features"], name)


@pytest.mark.parametrize("shape", [(1, 3, 1), (2, 3, 2), (33, 3, 15), (31, 3, 16), (70, 3, 17),
                                   (5, 3, 33), (129, 3, 64), (300, 3, 301)])
@pytest.mark.parametrize("semiring", ["reals", "arctic"])
def test_jit_shapes_vs_generic_and_oracle(shape, semiring, force_jit, monkeypatch):
    """Ragged batches (n not a multiple of the CTA, odd lengths, lengths below
    and across the tile) through the generated kernels: bit-identical to the
    generic kernel and to the oracle."""
    monkeypatch.setenv("FRUITS_B200_CHAIN", "force")     # (bushy arctic tries prefer fb_jit_slice)
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": {"of_weight": [3, 3]}, "mode": "extended",
                                 "semiring": semiring}],
                        "sieves": [["NPI", {"q": [0.4, 1.0]}], ["NPI", {"inc": 2}],
                                   ["MPI", {"inc": 0}], ["PPV", {}], ["MAX", {}], ["MIN", {}],
                                   ["END", {}]]}]}
    X = np.random.default_rng(shape[0] + shape[2]).standard_normal(shape)
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(1)
    fruit.fit(X)
    res = fruit.transform(X)
    assert _routes(fruit) == ["fb_jit_chain" if semiring == "arctic" else "fb_jit_slice"]
    if semiring == "arctic":
        # the same plan through the thread-per-series kernel
        monkeypatch.setenv("FRUITS_B200_CHAIN", "0")
        res2 = fruit.transform(X)
        assert _routes(fruit) == ["fb_jit_slice"]
        assert_exact(res, res2, "chain kernel vs thread-per-series kernel")
    monkeypatch.setenv("FRUITS_B200_JIT", "0")
    gen = fruit.transform(X)
    assert _routes(fruit) == ["fb::lns_kernel"]
    assert_exact(res, gen, "generated vs generic kernel")
    of = orc.OracleFruit(spec)
    np.random.seed(1)
    of.fit(X)
    ref = of.transform(X)
    mpi = np.zeros(ref.shape[1], dtype=bool)
    mpi[2::7] = True                       # MPI means: summation order of numba fastmath
    assert_exact(res[:, ~mpi], ref[:, ~mpi], "generated kernel vs oracle")
    assert_close(res[:, mpi], ref[:, mpi], 1e-12, "MPI")


def test_jit_bounded_quantile_intervals(force_jit):
    """(lo, hi] intervals with finite upper bounds and bounded MAX/MIN."""
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [],
                        "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended"}],
                        "sieves": [["NPI", {"q": [0.2, 0.7]}], ["MAX", {"q": [0.1, 0.9]}],
                                   ["MIN", {"q": [0.1, 0.9]}], ["END", {}]],
                        "fit_sample_size": 1.0}]}
    X = np.random.default_rng(3).standard_normal((40, 2, 90)).cumsum(axis=2)
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(2)
    fruit.fit(X)
    np.random.seed(2)
    of.fit(X)
    res = fruit.transform(X)
    assert _routes(fruit) == ["fb_jit_slice"]
    assert_exact(res, of.transform(X), "bounded intervals")


def test_jit_large_batch_default_route():
    """Large batches take the generated kernel without being forced."""
    X = np.random.default_rng(8).standard_normal((6000, 3, 128))
    fruit = specs.build_fruit(fruits, specs.SPECS["C5_sweep"])
    np.random.seed(0)
    fruit.fit(X)
    res = fruit.transform(X)
    assert _routes(fruit) == ["fb_jit_slice"]
    os.environ["FRUITS_B200_JIT"] = "0"
    try:
        gen = fruit.transform(X)
    finally:
        del os.environ["FRUITS_B200_JIT"]
    assert_exact(res, gen, "generated vs generic kernel at 6000 series")


# ---------------------------------------------------------------------------
# (f) cosine weighted ISS (reference: fruits/iss/cos.py)

@pytest.mark.parametrize("name", sorted(__import__("cases").COS_CASES))
def test_coswiss_golden(name, golden_dir):
    from cases import COS_CASES
    g = np.load(os.path.join(golden_dir, "cos.npz"))
    desc, shape, kind = COS_CASES[name]
    X = make_iss_input(shape, kind)
    iss = specs.build_iss(fruits, desc)
    res = iss.transform(X)
    assert res.shape == g[name].shape == (iss.n_iterated_sums(),) + (shape[0], shape[2])
    assert_close(res, g[name], 1e-9, name)          # sin/cos: device libm vs host libm
    assert "|".join(iss.label(i) for i in range(res.shape[0])) == str(g[name + "_labels"])
    # batches of words, like FruitSlice iterates them
    parts = list(iss.batch_transform(X, batch_size=1))
    assert len(parts) == len(iss.words)
    assert_exact(np.concatenate(parts), res, "batch_transform")
    # brute force (reference tests/signature/test_cosine.py): word [1][2], exponent 1
    if name == "cos_e2":
        f, T = 0.25, shape[2]
        w = np.cos(np.pi * (np.arange(T)[:, None] - np.arange(T)[None, :]) / (f * (T - 1))) ** 2
        want = np.zeros((shape[0], T))
        for t in range(T):
            for j in range(1, t + 1):
                want[:, t] += X[:, 1, j] * np.sum(X[:, 0, :j] * w[j, :j], axis=1)
        np.testing.assert_allclose(res[2], want, rtol=1e-6, atol=1e-8)


def test_coswiss_pipeline_golden(golden_dir):
    """Slices 2-3 of experiments/fruit_reduced.py (CosWISS exponents 1 and 2, total)."""
    g = np.load(os.path.join(golden_dir, "pipeline_C2_cos.npz"))
    X = specs.make_input("C2_cos", 16)
    fruit = specs.build_fruit(fruits, specs.SPECS["C2_cos"])
    assert fruit.nfeatures() == int(g["nfeatures"]) == 2310
    np.random.seed(0)
    fruit.fit(X)
    res = fruit.transform(X)
    assert_close(fitted_thresholds(fruit), g["thresholds"], 1e-9, "thresholds")
    _assert_features_close(res, g["features"], "C2_cos")
    if str(g["labels"]) == "IndexError":
        # Arctic(argmax=True): the reference's ISS._label asks the cache plan for rows it
        # does not hold (fruits/iss/iss.py:195-198); same here
        with pytest.raises(IndexError):
            fruit.label(res.shape[1] - 1)
    else:
        labels = "|".join(fruit.label(i) for i in
                          sorted(set(np.linspace(0, res.shape[1] - 1, 23).astype(int))))
        assert labels == str(g["labels"])
    assert fruit.summary() == str(g["summary"])
    # the expansion is compiled into the plan-specialised kernel ...
    assert _routes(fruit) == ["fb_jit_slice", "fb_jit_slice"]
    # ... and agrees with the materialise + stand-alone sieves route
    os.environ["FRUITS_B200_JIT"] = "0"
    try:
        composed = fruit.transform(X)
    finally:
        del os.environ["FRUITS_B200_JIT"]
    _assert_features_close(res, composed, "generated kernel vs composed route")


@pytest.mark.parametrize("name", sorted(__import__("cases").COS_RANDOM_CASES))
def test_randomised_coswiss_golden(name, golden_dir):
    """The randomised CosWISS variants (a random two-layer network per word and
    frequency in front, random dropout ahead of every cumulative sum) against
    outputs frozen from the reference: same draws in fit, same generator state,
    iterated sums within the CosWISS bound."""
    from cases import COS_RANDOM_CASES
    g = np.load(os.path.join(golden_dir, "cos2.npz"))
    desc, shape, kind = COS_RANDOM_CASES[name]
    X = make_iss_input(shape, kind)
    iss = specs.build_iss(fruits, desc)
    assert iss.requires_fitting and iss.copy().requires_fitting
    with pytest.raises(RuntimeError):
        iss.transform(X)
    np.random.seed(3)
    iss.fit(X)
    assert np.random.random() == float(g[name + "_rng"]), "fit consumed the RNG differently"
    res = iss.transform(X)
    assert_close(res, g[name], 1e-9, name)
    batches = np.concatenate(list(iss.batch_transform(X, batch_size=2)))
    assert_exact(batches, res, "batch_transform")


def test_coswiss_of_one_letter_words_only():
    """No junction anywhere (one-letter words, no total weighting): the cosine
    weights drop out and the sums are plain cumulative sums."""
    X = np.random.default_rng(2).standard_normal((5, 2, 33))
    iss = fruits.CosWISS([fruits.words.SimpleWord("[1]"), fruits.words.SimpleWord("[12]")],
                         freqs=[0.1, 0.3])
    res = iss.transform(X)
    assert_exact(res[0], np.cumsum(X[:, 0], axis=1), "[1]")
    assert_exact(res[1], res[0], "[1], second frequency")
    assert_exact(res[2], np.cumsum(X[:, 0] * X[:, 1], axis=1), "[12]")


def test_coswiss_four_letter_words_generated_kernel():
    """Four-letter words with exponent 2 and the total weighting (81 expansion
    terms, 120 running sums per word and frequency -- part of them live in
    local memory): generated kernel vs the composed route vs the oracle."""
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": ["[1][2][1][2]", "[1][1][2]", "[2]", "[12][1][1][-2]"],
                                 "coswiss": {"freqs": [0.05, 0.35], "exponent": 2, "total": True}}],
                        "sieves": [["NPI", {"q": [0.5, 1.0]}], ["MPI", {"inc": 0}], ["MAX", {}],
                                   ["END", {}]],
                        "fit_sample_size": 1.0}]}
    X = np.random.default_rng(21).random((37, 2, 53)) + 0.5
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(4)
    fruit.fit(X)
    res = fruit.transform(X)
    assert _routes(fruit) == ["fb_jit_slice"]
    os.environ["FRUITS_B200_JIT"] = "0"
    try:
        composed = fruit.transform(X)
    finally:
        del os.environ["FRUITS_B200_JIT"]
    _assert_features_close(res, composed, "generated kernel vs composed route")
    of = orc.OracleFruit(spec)
    np.random.seed(4)
    of.fit(X)
    assert_close(fitted_thresholds(fruit), oracle_thresholds(of), 1e-9, "thresholds")
    _assert_features_close(res, of.transform(X), "generated kernel vs oracle")


# ---------------------------------------------------------------------------
# Bayesian semiring, CUR / AVG / STD, CPV

def test_bayesian_slice_with_curvature_and_components():
    """A slice over the Bayesian semiring (scan kernel, sieved on materialised
    sums) with the rank-2 sieves, against the oracle; thresholds bit-exact,
    counts bit-exact, curvature sums within 1e-12."""
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [],
                        "iss": [{"words": {"of_weight": [3, 2]}, "mode": "extended",
                                 "semiring": "bayesian"}],
                        "sieves": [["NPI", {"q": [0.4, 1.0]}], ["CPV", {"quantile": [0.3, 0.8]}],
                                   ["CUR", {"cut": [0.5, -1], "q": [-1.0, 0.5, 1.0]}],
                                   ["MAX", {}], ["END", {}]],
                        "fit_sample_size": 1.0}]}
    X = np.random.default_rng(17).random((23, 2, 333))
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(5)
    fruit.fit(X)
    np.random.seed(5)
    of.fit(X)
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
    res, ref = fruit.transform(X), of.transform(X)
    assert res.shape == ref.shape == (23, fruit.nfeatures())
    nf = sum(s.nfeatures() for s in fruit.get_slice().get_sieves())
    cur_cols = np.zeros(res.shape[1], dtype=bool)
    for e in range(res.shape[1] // nf):
        cur_cols[e * nf + 3:e * nf + 7] = True       # NPI 1, CPV 2, CUR 4, MAX 1, END 1
    assert_exact(res[:, ~cur_cols], ref[:, ~cur_cols], "counts / max / end")
    assert_close(res[:, cur_cols], ref[:, cur_cols], 1e-12, "curvature")


def test_bayesian_batch_transform_and_ragged_lengths():
    """Partial emission ranges and lengths around the scan tile (256)."""
    from oracle import pipeline as orc
    desc = {"words": ["[1][2][1]", "[1][2]", "[2][2][1][1]", "[1]"], "mode": "extended",
            "semiring": "bayesian"}
    for T in (1, 2, 255, 256, 257, 700):
        X = np.random.default_rng(T).random((3, 2, T)) + 0.25
        iss = specs.build_iss(fruits, desc)
        ref = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
        assert_exact(iss.transform(X), ref, f"T={T}")
        parts = list(iss.batch_transform(X, batch_size=3))
        assert_exact(np.concatenate(parts, axis=0), ref, f"batches T={T}")


def test_rank2_sieves_large_rows_property():
    """CUR == sum of squared second differences, CPV == rising edges, on
    20,000 rows (no oracle needed: closed forms in numpy)."""
    Y = np.random.default_rng(3).standard_normal((20000, 96)).cumsum(axis=1)
    cur = fruits.sieving.CUR()
    cur.fit(Y)
    d1 = np.diff(Y, axis=1, prepend=Y[:, :1]); d1[:, 0] = 0.0
    d2 = np.diff(d1, axis=1, prepend=d1[:, :1]); d2[:, 0] = 0.0
    assert_close(cur.transform(Y)[:, 0], (d2 ** 2).sum(axis=1), 1e-12, "CUR")
    cpv = fruits.sieving.CPV(quantile=0.0, constant=True)
    cpv.fit(Y)
    ind = (Y >= 0.0).astype(np.int8)
    edges = ((ind[:, 1:] - ind[:, :-1]) == 1).sum(axis=1)
    assert_exact(cpv.transform(Y)[:, 0], 2 * edges / 96, "CPV")


def test_corbeille_fruitify_on_ucr_layout(tmp_path):
    """The reference's experiment harness (fruitify / fruitify_all) against the
    GPU path on a synthetic dataset in the UCR .txt layout: two well separated
    classes must be classified, the CSV must be written."""
    import corbeille
    from test_host_api import _write_ucr
    rng = np.random.default_rng(1)
    _write_ucr(str(tmp_path), "Gamma", rng, n_train=40, n_test=30, length=64)
    fruit = specs.build_fruit(fruits, specs.SPECS["C2_reduced"])
    data = corbeille.data.load(str(tmp_path / "Gamma"))
    np.random.seed(0)
    seconds, acc = corbeille.fruitify(data, fruit)
    assert seconds > 0 and 0.5 <= acc <= 1.0
    df = corbeille.fruitify_all(str(tmp_path), fruit, output_csv=str(tmp_path / "res.csv"))
    assert list(df.columns) == ["Dataset", "Accuracy", "Time"] and df["Dataset"][0] == "Gamma"
    assert (tmp_path / "res.csv").exists()


def test_corbeille_decide_which_fruit_equals_the_reference(golden_dir):
    """decide_which_fruit on the committed dataset under three seeds picks the
    candidates the reference's harness picks with the reference's Fruit (the
    validation splits come from the global numpy RNG through sklearn)."""
    import corbeille
    g = np.load(os.path.join(golden_dir, "corbeille2.npz"))
    d = np.load(os.path.join(golden_dir, "corbeille.npz"))
    Xtr, ytr = np.nan_to_num(d["Delta_X_train"]), d["Delta_y_train"]
    choices = [specs.build_fruit(fruits, specs.SPECS[n]) for n in ("R_decide_a", "R_decide_b")]
    choices.append((choices[0], specs.build_fruit(fruits, specs.SPECS["C2_reduced"])))
    picked = []
    for seed in (0, 1, 2):
        np.random.seed(seed)
        chosen = corbeille.decide_which_fruit(choices, n_splits=2)(Xtr, ytr)
        assert chosen is not choices[0] and chosen is not choices[1]        # a deep copy
        picked.append(chosen.nfeatures())
    assert picked == list(g["decide_nfeatures"])
    lonely = ytr.copy()
    lonely[0] = 7
    assert corbeille.decide_which_fruit(choices)(Xtr, lonely).nfeatures() == int(g["decide_lonely"])


def test_corbeille_fruitify_equals_the_reference(golden_dir):
    """``corbeille.fruitify`` on the committed UCR-layout datasets: the features it
    classifies are ``Fruit.transform`` of the loaded arrays, they agree with the
    features of the reference's Fruit on the same files (corbeille.npz), and the
    default classifier reaches the reference's accuracy."""
    import corbeille
    g = np.load(os.path.join(golden_dir, "corbeille.npz"))
    for name in ("Delta", "Eps"):
        data = corbeille.data.load(os.path.join(golden_dir, "ucr", name))
        seen = []

        class Spy:                                     # a classifier that keeps what it is given
            def __init__(self):
                self.inner = corbeille.fruitifier._default_classifier()

            def fit(self, F, y):
                seen.append(F.copy())
                self.inner.fit(F, y)

            def score(self, F, y):
                seen.append(F.copy())
                return self.inner.score(F, y)

        fruit = specs.build_fruit(fruits, specs.SPECS["C2_reduced"])
        np.random.seed(0)
        seconds, acc = corbeille.fruitify(data, fruit, classifier=Spy())
        assert seconds > 0
        assert_exact(seen[0], fruit.transform(np.nan_to_num(data[0])), "fruitify train features")
        assert_exact(seen[1], fruit.transform(np.nan_to_num(data[2])), "fruitify test features")
        _assert_features_close(seen[0], g[f"{name}_features_train"], f"corbeille {name} train")
        _assert_features_close(seen[1], g[f"{name}_features_test"], f"corbeille {name} test")
        assert abs(acc - float(g[f"{name}_accuracy"])) <= 1.0 / len(data[3]) + 1e-12


@pytest.mark.parametrize("n", [40, 4100])          # generic kernel / generated kernel
def test_degenerate_values_match_the_oracle(n):
    """Division by zero under negative exponents (inf, nan), overflow to +-inf,
    constant series, STD of a constant series: nan / inf propagate through the
    sums, the sieves and np.nan_to_num exactly as in the reference (means
    within the summation-order tolerance)."""
    from oracle import pipeline as orc
    sieves = [["NPI", {"q": [0.5, 1.0]}], ["MPI", {}], ["PPV", {}], ["MAX", {}], ["MIN", {}],
              ["END", {}]]
    spec = {"slices": [
        {"preps": [], "iss": [{"words": ["[-1]", "[1][-1]", "[-11][2]", "[2][-2][1]", "[1][2]"],
                               "mode": "extended"}], "sieves": sieves, "fit_sample_size": 1.0},
        {"preps": [], "iss": [{"words": ["[1][2]", "[-1][2][1]", "[11]"], "mode": "extended",
                               "semiring": "arctic"}], "sieves": sieves, "fit_sample_size": 1.0},
        {"preps": [["STD", {}]], "iss": [{"words": ["[1]", "[1][2]"], "mode": "extended"}],
         "sieves": sieves, "fit_sample_size": 1.0}]}
    X = np.random.default_rng(2).standard_normal((n, 2, 64))
    X[1, 0, 10] = 0.0
    X[2, 0, 0] = 0.0
    X[3, 1, 5:9] = 0.0
    X[4] = 1.5
    X[5, 0, 20] = 1e200
    X[6, 0, 20] = -1e200
    X[7, 1, 3] = 1e308
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(1)
    fruit.fit(X)
    np.random.seed(1)
    with np.errstate(all="ignore"):
        of.fit(X)
        ref = of.transform(X)
    res = fruit.transform(X)
    assert not np.isnan(res).any()                   # fruit.py:172
    with np.errstate(all="ignore"):
        same = (res == ref) | (np.abs(res - ref) <= 1e-12 * np.maximum(np.abs(ref), 1.0))
    assert same.all(), f"{(~same).sum()} of {same.size} differ, first {np.argwhere(~same)[:3]}"


def test_sieve_wrappers_in_a_slice():
    """INC / INT sieve wrappers (fruits/sieving/wrapper.py) inside a fruit:
    composed route vs the oracle, thresholds and features bit-exact."""
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended"}],
                        "sieves": [["INC", {"sieve": ["NPI", {"q": [0.3, 1.0], "inc": 0}],
                                            "shift": 2}],
                                   ["INT", {"sieve": ["MAX", {"q": [-1.0, 0.6]}]}],
                                   ["INC", {"sieve": ["END", {"cut": [9, -1]}], "depth": 0}],
                                   ["NPI", {}]],
                        "fit_sample_size": 1.0}]}
    X = np.random.default_rng(12).standard_normal((19, 2, 77)).cumsum(axis=2)
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(6)
    fruit.fit(X)
    np.random.seed(6)
    of.fit(X)
    assert fruit.nfeatures() == of.nfeatures()
    assert_exact(fruit.transform(X), of.transform(X), "wrapped sieves")
    assert fruit.get_slice().get_sieves()[0].label(0).startswith("INC of NPI")


# ---------------------------------------------------------------------------
# BASELINE.json full sizes (one GPU's share): size-independent properties

def test_sweep_full_shard_size():
    """C5 at the size bench.py times (524,288 series x 3 x 1024, generated on
    the device like the bench does): sampled rows equal the oracle bit for
    bit, the run is deterministic, a chunk transformed on its own gives the
    same rows (row independence), and the feature ranges hold everywhere."""
    from oracle import pipeline as orc
    n = 524288
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs 40 GB of free device memory")
    spec = specs.SPECS["C5_sweep"]
    gen = torch.Generator("cuda").manual_seed(1234)
    X = torch.randn((n, 3, 1024), dtype=torch.float64, device="cuda", generator=gen)
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(0)
    fruit.fit(X)
    out = fruit.transform_device(X)
    assert out.shape == (n, 2225) and bool(torch.isfinite(out).all())
    checksum = out.sum(dim=0)
    again = fruit.transform_device(X)
    assert bool(torch.equal(again, out)), "two runs differ"
    del again
    lo = 300000
    part = fruit.transform_device(X[lo:lo + 70000])
    assert bool(torch.equal(part, out[lo:lo + 70000])), "rows depend on their batch"
    del part
    pick = torch.tensor(sorted(np.random.default_rng(1).choice(n, 40, replace=False).tolist())
                        + [0, n - 1], device="cuda")
    Xs = X.index_select(0, pick).cpu().numpy()
    np.random.seed(0)
    fit_row = np.random.randint(0, n)                 # the draw fit() made (fruit.py:434)
    of = orc.OracleFruit(spec)
    np.random.seed(0)
    of.fit(X[fit_row:fit_row + 1].cpu().numpy())      # randint(0, 1) == 0: the same row
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
    assert_exact(out.index_select(0, pick).cpu().numpy(), of.transform(Xs), "sampled rows")
    f = out.view(n, 445, 5)
    assert bool((f[..., 0] == f[..., 0].round()).all()) and float(f[..., 0].min()) >= 0
    assert float(f[..., 0].max()) <= 1024 and float(f[..., 1].min()) >= 0
    assert float(f[..., 1].max()) <= 1
    assert bool((f[..., 3] <= f[..., 4]).all()) and bool((f[..., 4] <= f[..., 2]).all())
    assert bool(torch.isfinite(checksum).all())


def test_twi_full_size():
    """C4 at full size (100,000 x 3 x 2048): sampled rows against the oracle
    (arctic slice bit-exact, L1-weighted slice within 1e-9), determinism."""
    from oracle import pipeline as orc
    spec = specs.SPECS["C4_twi"]
    X = specs.make_input("C4_twi")
    assert X.shape == (100000, 3, 2048)
    Xd = torch.from_numpy(X).cuda()
    fruit = specs.build_fruit(fruits, spec)
    fruit.fit(Xd)
    out = fruit.transform_device(Xd)
    assert out.shape == (100000, 1725)
    assert bool(torch.equal(fruit.transform_device(Xd), out))
    pick = np.array([0, 1, 31, 32, 4097, 50000, 99999])
    of = orc.OracleFruit(spec)
    of.fit(X[pick])
    ref = of.transform(X[pick])
    res = out[torch.from_numpy(pick).cuda()].cpu().numpy()
    assert_exact(res[:, 1533:], ref[:, 1533:], "arctic slice")
    _assert_features_close(res[:, :1533], ref[:, :1533], "weighted slice")


def _np_increments(V, t, inc):
    rows = V.reshape(V.shape[0], -1, t).copy()
    for _ in range(inc):
        d = np.zeros_like(rows)
        d[..., 1:] = rows[..., 1:] - rows[..., :-1]
        rows = d
    return rows.reshape(V.shape[0], -1)


@pytest.mark.parametrize("t,n", [(1, 700), (2, 513), (37, 300), (256, 40), (1000, 9), (50, 4000)])
def test_order_stats_multi_matches_numpy(t, n):
    """fb_order_stats_multi (three reads for up to four selections, increments
    formed on the fly) == np.quantile of the materialised increments, bit for
    bit: ties, zeros, negative values, huge values, a NaN problem."""
    from fruits_b200.sieving.abstract import quantile_multi
    rng = np.random.default_rng(t * 1000 + n)
    P = 5
    V = rng.standard_normal((P, n * t)).cumsum(axis=1)
    V[1] = np.round(V[1])                      # many ties
    V[2, ::7] = 0.0                            # (fewer than the candidate list holds)
    V[3] *= 1e150
    V[4, (n * t) // 2] = np.nan
    pairs = [(0, 0.5), (1, 0.5), (2, 0.3), (1, 0.999), (0, 0.0), (2, 1.0), (0, 0.123)]
    got = quantile_multi(torch.from_numpy(V).cuda(), t, pairs)
    for inc, q in pairs:
        ref = np.quantile(_np_increments(V, t, inc), q, axis=1)
        res = got[(inc, q)]
        if res is None:
            # legitimate only if some problem has more keys in the quantile's 24-bit
            # bucket (sign, exponent, 12 mantissa bits) than the candidate list holds
            A = _np_increments(V, t, inc)
            lower = np.quantile(A, q, axis=1, method="lower")
            top = A.view(np.uint64) >> np.uint64(40)
            big = [(top[p] == (np.float64(lower[p]).view(np.uint64) >> np.uint64(40))).sum()
                   for p in range(P) if not np.isnan(lower[p])]
            assert max(big) > 65536, (inc, q, big)
            continue
        assert_exact(res, ref, f"inc={inc} q={q} t={t}")


def test_order_stats_multi_reports_large_buckets():
    """More equal values at the quantile than the candidate list holds: the
    selection is reported as not done and fit takes the eight-pass path."""
    from fruits_b200.sieving.abstract import quantile_multi, quantile_rows
    rng = np.random.default_rng(0)
    V = rng.standard_normal((2, 400 * 500))
    V[0, :150000] = 0.25                       # the median bucket holds 150,000 keys
    Vd = torch.from_numpy(V).cuda()
    got = quantile_multi(Vd, 500, [(0, 0.5), (0, 0.999)])
    assert got[(0, 0.5)] is None
    assert_exact(got[(0, 0.999)], np.quantile(V, 0.999, axis=1), "unaffected selection")
    assert_exact(quantile_rows(Vd, 0.5), np.quantile(V, 0.5, axis=1), "eight-pass path")
    # through fit: constant rows make the increments' median bucket huge
    spec = {"slices": [{"iss": [{"words": ["[1]"], "mode": "single"}],
                        "sieves": [["NPI", {"q": [0.5, 1.0]}], ["MPI", {"q": [0.2, 0.8], "inc": 0}]],
                        "fit_sample_size": 1.0}]}
    from oracle import pipeline as orc
    X = rng.standard_normal((700, 1, 200))
    X[:500] = 0.0
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    fruit.fit(X)
    of.fit(X)
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")


def test_series_longer_than_the_16_bit_counters():
    """T >= 65536: the fused kernels pack their counters in 16 bits, such
    series take the materialise + stand-alone sieve route; same numbers."""
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": ["[1]", "[1][2]", "[2][1][1]"], "mode": "extended"}],
                        "sieves": [["NPI", {"q": [0.5, 1.0]}], ["PPV", {}], ["MAX", {}], ["MIN", {}],
                                   ["END", {}]],
                        "fit_sample_size": 1}]}
    X = np.random.default_rng(3).standard_normal((3, 2, 70001)) * 0.01
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(2)
    fruit.fit(X)
    np.random.seed(2)
    of.fit(X)
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
    res = fruit.transform(X)
    assert res[:, 0].max() > 0                      # NPI counts beyond 16 bits are possible
    assert_exact(res, of.transform(X), "long series")


def test_words_over_many_dimensions():
    """of_weight(2, 12): 90 words over 12 input dimensions -- more distinct
    dimensions than one kernel block stages; ISS.transform and a fruit must
    still equal the oracle bit for bit."""
    from oracle import pipeline as orc
    desc = {"words": {"of_weight": [2, 12]}, "mode": "extended"}
    X = np.random.default_rng(4).standard_normal((6, 12, 40))
    ref = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
    assert_exact(specs.build_iss(fruits, desc).transform(X), ref, "iss over 12 dimensions")
    spec = {"slices": [{"preps": [], "iss": [desc],
                        "sieves": [["NPI", {"q": [0.5, 1.0]}], ["MAX", {}], ["END", {}]],
                        "fit_sample_size": 1.0}]}
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    fruit.fit(X)
    of.fit(X)
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
    assert_exact(fruit.transform(X), of.transform(X), "features over 12 dimensions")


def test_mid_size_batches_use_a_compiled_kernel_only_if_it_exists(monkeypatch):
    """1,000 <= n < 4,096: the generated kernel is used when its cubin is in
    memory or on disk, never compiled for such a batch; identical numbers on
    either route."""
    from fruits_b200 import _jit
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": ["[1]", "[1][2]", "[2][2][1]", "[1][1][1]"],
                                 "mode": "extended"}],
                        "sieves": [["NPI", {"q": [0.37, 1.0]}], ["MAX", {}], ["END", {}]],
                        "fit_sample_size": 1}]}
    X = np.random.default_rng(31).standard_normal((5000, 2, 48))
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(1)
    fruit.fit(X)
    compiled = []
    real_init = _jit.JitSlice.__init__

    def spy(self, gen):
        compiled.append(gen.digest())
        real_init(self, gen)

    monkeypatch.setattr(_jit.JitSlice, "__init__", spy)
    monkeypatch.setattr(_jit.JitSlice, "_loaded", {})
    monkeypatch.setattr(_jit, "CACHE_DIR", "/nonexistent-jit-cache")
    small = fruit.transform(X[:2000])                 # nothing compiled yet: generic kernel
    assert _routes(fruit) == ["fb::lns_kernel"] and not compiled
    big = fruit.transform(X)                          # 5,000 series: compiled now
    assert _routes(fruit) == ["fb_jit_slice"] and len(compiled) == 1
    again = fruit.transform(X[:2000])                 # ... and reused for the mid-size batch
    assert _routes(fruit) == ["fb_jit_slice"] and len(compiled) == 1
    assert_exact(again, small, "generated vs generic kernel")
    assert_exact(big[:2000], small, "row independence")
    tiny = fruit.transform(X[:500])                   # below the crossover: generic kernel
    assert _routes(fruit) == ["fb::lns_kernel"]
    assert_exact(tiny, small[:500], "small batch")


@pytest.mark.parametrize("kat", range(len(SIEVE_KATS)))
def test_reference_kat_sieve_table(kat):
    """Known answers of the reference's own sieve tests (data in cases.py)
    through the stand-alone seed API of the product."""
    desc, block, want = SIEVE_KATS[kat]
    sv = specs._sieve(fruits, desc)
    np.random.seed(0)
    got = sv.fit_transform(KAT_X[block])
    np.testing.assert_allclose(got, np.array(want, dtype=float), rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("name", sorted(__import__("cases").PREP2_CASES))
def test_prep2_golden(name, golden_dir):
    """MAV, LAG, FFN, RIN, RDW, JLD, SPE, RPE, CTS, QTC, FUN, DIL, WIN, DOT, PDD and
    NRM(scale_dim=True) (csrc/prep_more.cu) against outputs frozen from the
    reference: fit under a seed, transform of the fit batch and of a second
    batch, numpy arrays and device tensors, the RNG left where the reference
    leaves it.  Copies / masks / np.where are bit-identical; the reference's
    numba fastmath loops, BLAS and libm calls within 1e-12 of the row maximum."""
    from cases import PREP2_CASES, PREP2_EXACT, make_prep2_inputs
    g = np.load(os.path.join(golden_dir, "preps2.npz"))
    desc = PREP2_CASES[name]
    X, X2 = make_prep2_inputs(name)
    prep = specs._prep(fruits, desc)
    np.random.seed(7)
    prep.fit(X)
    assert np.random.random() == float(g[name + "_rng"]), "fit consumed the RNG differently"
    res, res2 = prep.transform(X), prep.transform(X2)
    on_device = prep.transform(torch.from_numpy(X2).cuda())
    assert isinstance(on_device, torch.Tensor) and on_device.is_cuda
    for got, key in ((res, name), (res2, name + "_2"), (on_device.cpu().numpy(), name + "_2")):
        assert got.dtype == np.float64 and got.flags.c_contiguous
        if desc[0] in PREP2_EXACT:
            assert_exact(got, g[key], key)
        else:
            assert_close(got, g[key], 1e-12, key)
    again = prep.copy()
    assert not hasattr(again, "_kernel") and not hasattr(again, "_weights")    # copies are unfitted


@pytest.mark.parametrize("key", ["L1_sqrt_relative", "L2_log1p"])
def test_increment_sum_weighting_with_python_transform(key, golden_dir):
    """L1 / L2 lookups whose sums go through the caller's scalar function
    (np.vectorize in the reference, fruits/iss/weighting.py:155-156), frozen from
    the reference: the lookup and the weighted iterated sums."""
    g = np.load(os.path.join(golden_dir, "preps2.npz"))
    W = fruits.iss.weighting
    w = (W.L1(relative=True, transform=np.sqrt) if key.startswith("L1")
         else W.L2(transform=np.log1p, scale=5, total=True))
    X = make_prep_input()
    iss = fruits.ISS([fruits.words.SimpleWord("[1][2]"), fruits.words.SimpleWord("[3][1][1]")],
                     mode=fruits.ISSMode.EXTENDED, weighting=w)
    assert_close(iss.transform(X), g["iss_" + key], 1e-9, "weighted iterated sums")
    w._cache = fruits.cache.SharedSeedCache(X)
    try:
        assert_close(w.get_lookup(X), g["lookup_" + key], 1e-14, "lookup")
    finally:
        del w._cache


def test_reference_preparateur_known_answers():
    """The hand-computed vectors of the reference's own preparateur tests
    (tests/preparation/test_filter.py:13-66, test_transform.py:12-160) through
    the GPU path: WIN, DOT, NRM(scale_dim), MAV, LAG, RIN with a planted kernel,
    JLD shapes, FFN against its own weights."""
    P = fruits.preparation
    Xw = np.array([[[1, 2, 4, 5, 6], [11, 22, 33, 44, 55]],
                   [[10, 20, 30, 40, 50], [111, 222, 333, 444, 555]]], dtype=float)
    np.testing.assert_allclose(P.WIN(0.0, 0.7).fit_transform(Xw), [
        [[1, 2, 0, 0, 0], [11, 22, 0, 0, 0]], [[10, 20, 30, 0, 0], [111, 222, 333, 0, 0]]])
    np.testing.assert_allclose(P.WIN(0.7, 1.0).fit_transform(Xw), [
        [[0, 2, 4, 5, 6], [0, 22, 33, 44, 55]], [[0, 0, 30, 40, 50], [0, 0, 333, 444, 555]]])
    np.testing.assert_allclose(P.DOT(0.4).fit_transform(X_1), [
        [[0., 0.8, 0., 5., 0.], [0., 1., 0., 0., 0.]], [[0., 8., 0., 6., 0.], [0., -1., 0., -0.5, 0.]]])
    np.testing.assert_allclose(P.DOT(0.9).fit_transform(X_1), [
        [[0, 0, 0, 5, 0], [0, 0, 0, 0, 0]], [[0, 0, 0, 6, 0], [0, 0, 0, -0.5, 0]]])
    ramp = np.arange(100, dtype=float)[np.newaxis, np.newaxis, :]
    want = np.zeros(ramp.shape)
    want[:, :, 9::10] = ramp[:, :, 9::10]
    np.testing.assert_allclose(P.DOT(0.1).fit_transform(ramp), want)
    np.testing.assert_allclose(P.NRM(scale_dim=True).fit_transform(X_1), [
        [[3/12, 7.8/12, 7/12, 1., 4/12], [9/12, 8/12, 7/12, 7/12, 0.]],
        [[13/16, 1., 10/16, 14/16, 8/16], [3/16, 7/16, 4/16, 7.5/16, 0.]]])
    np.testing.assert_allclose(P.MAV(2).fit_transform(X_1), [
        [[0, -1.6, 0.4, 2.5, 1], [0, 1.5, 0.5, 0, -3.5]],
        [[0, 6.5, 5, 4, 3], [0, -3, -2.5, -2.25, -4.25]]])
    np.testing.assert_allclose(P.MAV(0.6).fit_transform(X_1), np.array([
        [[0, 0, -3.2, 5.8, 2.], [0, 0, 3., 1., -7.]],
        [[0, 0, 15., 16., 8.], [0, 0, -10., -5.5, -12.5]]]) / 3)
    np.testing.assert_allclose(P.LAG().fit_transform(X_1), [
        [[-4., 0.8, 0.8, 0., 0., 5., 5., -3., -3.], [-4., -4., 0.8, 0.8, 0., 0., 5., 5., -3.],
         [2., 1., 1., 0., 0., 0., 0., -7., -7.], [2., 2., 1., 1., 0., 0., 0., 0., -7.]],
        [[5., 8., 8., 2., 2., 6., 6., 0., 0.], [5., 5., 8., 8., 2., 2., 6., 6., 0.],
         [-5., -1., -1., -4., -4., -0.5, -0.5, -8., -8.],
         [-5., -5., -1., -1., -4., -4., -0.5, -0.5, -8.]]])
    other = np.random.default_rng(0).random((46, 2, 189))
    rin = P.RIN(2, adaptive_width=True)
    rin.fit(other)
    rin._kernel = np.array([[4., 1.], [4., 1.]])
    rin._ndim_per_kernel = np.array([1, 1], dtype=np.int32)
    rin._dims_per_kernel = np.array([0, 1], dtype=np.int32)
    np.testing.assert_allclose(rin.transform(X_1), [
        [[-4., 4.8, 15.2, 1.8, -8.], [2., -1., -9., -4., -7.]],
        [[5., 3., -26., -28., -14.], [-5., 4., 17., 7.5, 8.5]]])
    rin = P.RIN(width=2, adaptive_width=False, out_dim=1)
    rin.fit(other)
    assert rin._kernel.shape == (2, 2)
    rin._kernel = np.array([[4., 1.], [2., 3.]])
    rin._ndim_per_kernel = np.array([2], dtype=np.int32)
    rin._dims_per_kernel = np.array([0, 1], dtype=np.int32)
    np.testing.assert_allclose(rin.transform(X_1), [[[0., 0, 8.2, -.2, -15.]],
                                                    [[0., 0., -17., -14.5, -12.5]]])
    wide = np.random.default_rng(1).random((46, 100, 189))
    assert P.JLD(25).fit_transform(wide).shape == (46, 25, 189)
    ffn = P.FFN(d_hidden=3, center=False, relu_out=False)
    ffn.fit(X_1)
    hidden = np.stack([ffn._weights1 @ X_1[i] + ffn._biases[:, np.newaxis] for i in range(2)])
    hidden = hidden * (hidden > 0)
    np.testing.assert_allclose(ffn.transform(X_1),
                               np.stack([ffn._weights2 @ hidden[i] for i in range(2)]))


@pytest.mark.parametrize("shape", [(3, 1, 5), (2, 3, 9), (4, 2, 17), (3, 3, 100), (2, 2, 257),
                                   (5, 4, 12), (7, 3, 1030)])
def test_preparateurs_equal_the_oracle_over_shapes(shape):
    """Every preparateur case on other shapes -- short series, one dimension,
    lengths across the 256-column tile of the kernels, corner cases of the fitted
    parameters (no room for DIL strips, DOT with n >= T, PDD of width zero, windows
    longer than the series) -- against the numpy restatement, which
    ``oracle/check_preps_sweep.py`` holds against the real reference on the same
    shapes: same draws in fit, same outputs."""
    from cases import PREP2_CASES, PREP2_EXACT
    from oracle import pipeline as orc
    from oracle import preps as more
    n, d, t = shape
    X = np.random.default_rng(n * 1000 + t).standard_normal(shape).cumsum(axis=2)
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
        if kind == "RIN" and ((args.get("kernel") is not None and d != 3)
                              or args.get("out_dim", -1) > d):
            continue
        if kind == "JLD" and args.get("distribute") and args.get("dim", 1) > d:
            continue
        np.random.seed(11)
        try:
            st = more.fit_prep(desc, Xc)
            after = np.random.random()
            with np.errstate(invalid="ignore"):
                want = more.transform_prep(desc, st, Xc, orc.RawCache(Xc))
        except Exception:
            continue                   # (the reference rejects this shape, too)
        prep = specs._prep(fruits, desc)
        np.random.seed(11)
        prep.fit(Xc)
        assert np.random.random() == after, f"{name}: fit consumed the RNG differently"
        got = prep.transform(Xc)
        if kind in PREP2_EXACT:
            assert_exact(got, want, f"{name} {shape}")
        else:
            assert_close(got, want, 1e-12, f"{name} {shape}")
        checked += 1
    assert checked >= 30


def test_preparateur_properties_at_scale():
    """Size-independent properties on 20,000 x 3 x 1,024 series (device tensors
    in and out): masks are idempotent and only ever zero values, LAG interleaves
    the series with itself, CTS shifts, QTC clips at a value of the batch, SPE is
    linear in the series, a prepared slice gives the same features for the same
    series whatever batch they arrive in."""
    P = fruits.preparation
    gen = torch.Generator(device="cuda").manual_seed(7)
    X = torch.randn((20000, 3, 1024), dtype=torch.float64, device="cuda", generator=gen).cumsum(2)
    np.random.seed(1)
    for prep in (P.DOT(7, 2), P.DIL(0.05), P.PDD(0.9, 0.5), P.WIN(0.2, 0.8), P.CTS(9, True)):
        prep.fit(X)
        once = prep.transform(X)
        assert torch.equal(prep.transform(once) if not isinstance(prep, P.WIN) else once, once)
        kept = once != 0
        assert torch.equal(once[kept], X[kept]) and 0 < int(kept.sum()) < X.numel()
    lag = P.LAG().transform(X)
    assert lag.shape == (20000, 6, 2047)
    assert torch.equal(lag[:, 0::2, 0::2], X) and torch.equal(lag[:, 1::2, 0::2], X)
    assert torch.equal(lag[:, 0::2, 1::2], X[:, :, 1:]) and torch.equal(lag[:, 1::2, 1::2], X[:, :, :-1])
    cts = P.CTS(0.25).fit_transform(X)
    assert torch.equal(cts[:, :, :768], X[:, :, 256:])
    assert torch.equal(cts[:, :, 768:], X[:, :, -1:].expand(-1, -1, 256))
    qtc = P.QTC(0.9)
    qtc.fit(X)
    cut = qtc.transform(X)
    assert float(cut.max()) == float(qtc._quantile) and torch.equal(cut[X <= cut.max()], X[X <= cut.max()])
    assert 0.099 < float((X > qtc._quantile).double().mean()) < 0.101
    spe = P.SPE(0.5)
    assert torch.allclose(spe.transform(2.0 * X), 2.0 * spe.transform(X), rtol=1e-15, atol=0.0)
    mav = P.MAV(8)
    mav.fit(X)
    assert torch.equal(mav.transform(torch.ones_like(X))[:, :, 7:], torch.ones_like(X)[:, :, 7:])
    # a slice behind a prepared copy: features of a series do not depend on its batch
    fruit = fruits.Fruit()
    fruit.add(P.LAG(), P.INC(), fruits.ISS(fruits.words.of_weight(2, 2), mode=fruits.ISSMode.EXTENDED),
              fruits.sieving.NPI(q=(0.5, 1.0)), fruits.sieving.MAX(), fruits.sieving.END())
    np.random.seed(0)
    fruit.fit(X[:64, :1])
    whole = fruit.transform(X[:, :1].contiguous())
    part = fruit.transform(X[4096:4096 + 512, :1].contiguous())
    assert torch.equal(whole[4096:4096 + 512], part)


def test_preparateur_edge_shapes():
    """Shapes at the edges: one time step, windows longer than the series, empty
    batches, the cache-row quirk of WIN / SPE on a one-series fit sample."""
    P = fruits.preparation
    one = np.array([[[2.0], [3.0]]])
    np.testing.assert_array_equal(P.LAG().transform(one), [[[2.0], [2.0], [3.0], [3.0]]])
    np.testing.assert_array_equal(P.CTS(1).fit_transform(one), one)
    np.testing.assert_array_equal(P.MAV(5).fit_transform(one), np.zeros_like(one))
    np.testing.assert_array_equal(P.MAV(1).fit_transform(one), one)
    X = np.random.default_rng(3).standard_normal((5, 2, 17)).cumsum(axis=2)
    empty = X[:0]
    for prep in (P.LAG(), P.CTS(2), P.DOT(), P.MAV(3), P.RPE(0.5), P.JLD(3), P.FFN()):
        prep.fit(X)
        assert prep.transform(empty).shape[0] == 0
    # numpy broadcasting of one series against the cached sums of a batch (what a
    # fit_sample_size=1 fit does, fruits/cache.py:97-112): the batch size comes back
    cache = fruits.cache.SharedSeedCache(X)
    spe = P.SPE(0.5, step_transform="L1")
    spe._cache = cache
    try:
        got = spe._transform_device(torch.from_numpy(X[2:3]).cuda()).cpu().numpy()
    finally:
        del spe._cache
    l1 = np.cumsum(np.abs(np.diff(X[:, 0, :], prepend=X[:, :1, 0])), axis=1)
    want = X[2:3] * np.sin(l1 / l1[:, -1:] ** 0.5)[:, np.newaxis, :]
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    win = P.WIN(0.2, 0.7)
    win._cache = cache
    try:
        got = win._transform_device(torch.from_numpy(X[3:5]).cuda()).cpu().numpy()
    finally:
        del win._cache
    l2 = np.cumsum(np.diff(X[:, 0, :], prepend=X[:, :1, 0]) ** 2, axis=1)
    want = np.zeros((2, 2, 17))
    for i in range(2):          # rows 0 and 1 of the cache serve rows 3 and 4 of X
        a, b = np.sum(l2[i] <= 0.2 * l2[i, -1]) - 1, np.sum(l2[i] <= 0.7 * l2[i, -1])
        want[i, :, a:b] = X[3 + i, :, a:b]
    np.testing.assert_array_equal(got, want)
    # the cached keep mask follows the series length and the fitted state
    dot = P.DOT(3)
    dot.fit(X)
    a = dot.transform(X)
    longer = np.concatenate((X, X), axis=2)
    b = dot.transform(longer)
    assert b.shape == longer.shape and np.array_equal(b[:, :, :17], a)
    assert np.array_equal(b[:, :, 2::3], longer[:, :, 2::3]) and np.count_nonzero(b) == b[:, :, 2::3].size
    dot._n, dot._first = 2, 0                  # (what a second fit would change)
    assert np.array_equal(dot.transform(X)[:, :, 0::2], X[:, :, 0::2])
    with pytest.raises(ValueError):
        P.RPE(0.5).transform(np.zeros((1, 3, 4)))
    with pytest.raises(RuntimeError):
        P.MAV(-1).fit_transform(X)           # (like the reference: fit sets no width)
    with pytest.raises(RuntimeError):
        P.DIL().transform(X)


def test_prepared_copy_then_fused_kernels(golden_dir, monkeypatch):
    """A slice whose preparateurs the kernels cannot apply while loading
    (LAG, DOT, RIN, MAV ...) writes a prepared copy and still takes a fused
    kernel -- the iterated sums are not materialised -- with the features of the
    composed route."""
    X = specs.make_input("R_preps")
    fruit = specs.build_fruit(fruits, specs.SPECS["R_preps"])
    np.random.seed(0)
    fruit.fit(X)
    routes = []
    FS = fruits.fruit.FruitSlice
    fused, composed = FS._transform_fused, FS._transform_composed
    monkeypatch.setattr(FS, "_transform_fused",
                        lambda self, *a, **k: (routes.append("fused"), fused(self, *a, **k))[1])
    monkeypatch.setattr(FS, "_transform_composed",
                        lambda self, *a, **k: (routes.append("composed"), composed(self, *a, **k))[1])
    res = fruit.transform(X)
    assert routes == ["fused", "fused", "fused", "composed"]
    monkeypatch.setattr(FS, "_transform_prepared_fused", lambda self, *a, **k: False)
    routes.clear()
    ref = fruit.transform(X)
    assert routes == ["composed"] * 4
    _assert_features_close(res, ref, "prepared copy + fused kernel vs composed route")
    # more series than the thread-per-series kernels ask for: same pipeline, generated kernels
    big = np.random.default_rng(9).standard_normal((4200, 2, 48)).cumsum(axis=2)
    monkeypatch.undo()
    a = fruit.transform(big)
    monkeypatch.setattr(FS, "_transform_prepared_fused", lambda self, *a, **k: False)
    b = fruit.transform(big)
    _assert_features_close(a, b, "generated kernels on the prepared copy vs composed route")


@pytest.mark.parametrize("name", sorted(__import__("cases").ARGMAX_CASES))
def test_arctic_argmax_golden(name, golden_dir):
    """Arctic(argmax=True) -- per level the running maximum and the positions
    that produced it, as a block scan over T of (maximum, first position) --
    against outputs frozen from the reference: bit-identical unweighted, position
    rows identical and values within 1e-12 with a weighting (device FMA order)."""
    from cases import ARGMAX_CASES
    g = np.load(os.path.join(golden_dir, "argmax.npz"))
    desc, shape, kind = ARGMAX_CASES[name]
    X = make_iss_input(shape, kind)
    iss = specs.build_iss(fruits, desc)
    res = iss.transform(X)
    assert res.shape == g[name].shape == (iss.n_iterated_sums(), shape[0], shape[2])
    if desc.get("weighting") is None:
        assert_exact(res, g[name], name)
    else:
        assert_close(res, g[name], 1e-12, name)
        first = 0
        for w in iss.words:          # the position rows are integers: no tolerance
            for k in range(len(w)):
                base = first + k + k * (k + 1) // 2
                assert_exact(res[base + 1:base + k + 2], g[name][base + 1:base + k + 2],
                             f"{name}: positions of level {k}")
            first += len(w) + len(w) * (len(w) + 1) // 2
    batches = np.concatenate(list(iss.batch_transform(X, batch_size=min(2, len(iss.words)))))
    assert_exact(batches, res, "batch_transform")
    single = fruits.ISS(iss.words, semiring=fruits.semiring.Arctic(argmax=True))
    with pytest.raises(NotImplementedError):
        single.transform(X)


def test_arctic_argmax_across_tiles_equals_the_oracle():
    """Series longer than one tile of the scan (256 steps), lengths at and around
    the tile boundary, plateaus (ties keep the first position)."""
    from oracle import pipeline as orc
    desc = {"words": ["[1][2][1]", "[2][-1]"], "mode": "extended", "semiring": "arctic_argmax"}
    iss = specs.build_iss(fruits, desc)
    for t in (1, 2, 255, 256, 257, 700):
        X = np.round(np.random.default_rng(t).standard_normal((3, 2, t)).cumsum(axis=2))
        want = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
        assert_exact(iss.transform(X), want, f"length {t}")


@pytest.mark.parametrize("semiring", ["reals", "arctic"])
def test_more_distinct_alphas_than_one_launch_holds(semiring):
    """Seven words with seven different alpha vectors: one launch holds four
    distinct alpha values, so the emissions are materialised in consecutive ranges
    (ISS._dim_pieces) and a slice with such an ISS takes the composed route; same
    numbers as the oracle."""
    from oracle import pipeline as orc
    words = ["[1][2]", "[2][1]", "[1][1]", "[2][2]", "[12][1]", "[1][12]", "[2][1][2]"]
    alphas = [[0.1 * (i + 1), 0.05 * (i + 2), 0.3][:w.count("[")] for i, w in enumerate(words)]
    desc = {"words": words, "mode": "extended", "semiring": semiring, "alphas": alphas,
            "weighting": ["Indices", {"scale": 3}]}
    X = np.random.default_rng(21).standard_normal((11, 2, 50)).cumsum(axis=2) / 4
    iss = specs.build_iss(fruits, desc)
    assert len(iss._dim_pieces(None)) > 1
    want = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
    got = iss.transform(X)
    if semiring == "arctic":
        assert_exact(got, want, "arctic, seven alpha vectors")
    else:
        assert_close(got, want, 1e-9, "reals, seven alpha vectors")
    spec = {"slices": [{"preps": [], "iss": [desc],
                        "sieves": [["NPI", {"q": [0.5, 1.0]}], ["END", {}]],
                        "fit_sample_size": 1.0}]}
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(0)
    fruit.fit(X)
    np.random.seed(0)
    of.fit(X)
    _assert_features_close(fruit.transform(X), of.transform(X), "slice with seven alpha vectors")


@pytest.mark.parametrize("semiring", ["reals", "arctic"])
@pytest.mark.parametrize("mode", ["single", "extended"])
def test_generic_words_equal_the_oracle(semiring, mode):
    """Words over Python letters: the letters are evaluated on the host, their
    rows go through the kernels as extra dimensions -- bit for bit the
    reference's Semiring._iterated_sum (oracle restatement, pinned by the
    R_letters golden), through transform, batch_transform and device tensors."""
    from oracle import pipeline as orc
    words = ["[ABS(1)DIM(2)][DIM(1)]", "[ABS(1)DIM(2)][RELU(2)][DIM(1)DIM(1)]", "[LAGDIFF(1)]",
             "[RELU(2)][DIM(1)DIM(1)]", "[ABS(1)DIM(2)]"]
    desc = {"words": words, "mode": mode, "semiring": semiring}
    X = np.random.default_rng(11).standard_normal((9, 2, 37)).cumsum(axis=2) / 3
    iss = specs.build_iss(fruits, desc)
    want = np.stack(list(orc.iss_iter(X, desc, orc.RawCache(X))))
    got = iss.transform(X)
    assert got.shape == want.shape == (iss.n_iterated_sums(), 9, 37)
    assert_exact(got, want, "generic words")
    assert_exact(iss.transform(torch.from_numpy(X).cuda()).cpu().numpy(), want, "device input")
    batches = list(iss.batch_transform(X, batch_size=2))
    assert len(batches) == 3
    assert_exact(np.concatenate(batches), want, "batch_transform")
    with pytest.raises(IndexError):
        iss.transform(X[:, :1])


@pytest.mark.parametrize("name", sorted(__import__("cases").LETTER_CASES))
def test_generic_words_golden(name, golden_dir):
    """Words over Python letters in the three semirings against ISS outputs frozen
    from the reference, bit for bit -- the Bayesian semiring with the shift of the
    reference's general recursion (letter rows moved in time around the unshifted
    kernels), mixed with SimpleWords on its fast path."""
    from cases import LETTER_CASES
    g = np.load(os.path.join(golden_dir, "letters.npz"))
    desc, shape, kind = LETTER_CASES[name]
    X = make_iss_input(shape, kind)
    iss = specs.build_iss(fruits, desc)
    res = iss.transform(X)
    if desc.get("weighting") is None or desc.get("semiring") == "arctic":
        assert_exact(res, g[name], name)
    else:
        # weighted SimpleWords (device exp: 1e-9) beside generic words, whose weighting the
        # reference ignores: those rows stay bit-identical (alpha = 0 in the twin)
        from oracle import pipeline as orc
        assert_close(res, g[name], 1e-9, name)
        words = desc["words"]
        plan = orc.cache_plan(words) if desc["mode"] == "extended" else [1] * len(words)
        first = 0
        for w, ext in zip(words, plan):
            if orc.is_generic_word(w):
                assert_exact(res[first:first + ext], g[name][first:first + ext], f"{name}: {w}")
            first += ext
    assert_exact(np.concatenate(list(iss.batch_transform(X, batch_size=2))), res, "batch_transform")
    chunks = [c for _, c in iss.iter_chunks(torch.from_numpy(X).cuda(),
                                           max_bytes=X.shape[0] * X.shape[2] * 8)]
    assert len(chunks) > 1
    assert_exact(torch.cat(chunks).cpu().numpy(), res, "iter_chunks")


def test_generic_words_in_a_bayesian_slice_equal_the_oracle():
    """A slice over generic words in the Bayesian semiring (sieved on materialised
    rows) against the oracle pipeline."""
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [["NRM", {}]],
                        "iss": [{"words": ["[ABS(1)][DIM(2)DIM(2)][ABS(1)DIM(2)]", "[ABS(1)][ABS(2)]",
                                           "[1][2]"], "mode": "extended", "semiring": "bayesian"}],
                        "sieves": [["NPI", {"q": [0.5, 1.0]}], ["MAX", {}], ["END", {}]],
                        "fit_sample_size": 1.0}]}
    X = np.random.default_rng(31).random((25, 2, 41)) + 0.05
    fruit = specs.build_fruit(fruits, spec)
    of = orc.OracleFruit(spec)
    np.random.seed(0)
    fruit.fit(X)
    np.random.seed(0)
    of.fit(X)
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
    assert_exact(fruit.transform(X), of.transform(X), "features")


def test_generic_words_take_the_fused_kernels(monkeypatch):
    """A slice whose ISS holds generic words runs its SimpleWord twin through
    the fused kernels (iterated sums never materialised) with the features of
    the composed route, for small batches and for the generated kernels."""
    spec = {"slices": [specs.SPECS["R_letters"]["slices"][0], specs.SPECS["R_letters"]["slices"][1]]}
    fruit = specs.build_fruit(fruits, spec)
    X = specs.make_input("R_letters")
    np.random.seed(0)
    fruit.fit(X)
    FS = fruits.fruit.FruitSlice
    routes = []
    fused = FS._transform_fused
    monkeypatch.setattr(FS, "_transform_fused",
                        lambda self, *a, **k: (routes.append("fused"), fused(self, *a, **k))[1])
    big = np.random.default_rng(12).standard_normal((4200, 2, 45)).cumsum(axis=2) / 4
    res, res_big = fruit.transform(X), fruit.transform(big)
    assert routes == ["fused"] * 4
    monkeypatch.setattr(FS, "_transform_prepared_fused", lambda self, *a, **k: False)
    _assert_features_close(res, fruit.transform(X), "generic words: fused vs composed")
    _assert_features_close(res_big, fruit.transform(big), "generic words: generated vs composed")
    assert routes == ["fused"] * 4


@pytest.mark.parametrize("name", ["R_mixed", "R_rng", "R_preps", "R_letters", "R_cosrand",
                                  "R_argmax"])
def test_extra_pipeline_golden(name, golden_dir):
    """Frozen outputs of the real reference: ``R_mixed`` -- a Bayesian slice
    (rank-2 sieves, sieve wrappers) and a slice of two chained ISS; ``R_rng``
    -- fractional fit samples and PPV subsamples, i.e. the order in which fit
    consumes the global numpy RNG."""
    g = np.load(os.path.join(golden_dir, f"pipeline_{name}.npz"))
    X = specs.make_input(name)
    fruit = specs.build_fruit(fruits, specs.SPECS[name])
    assert fruit.nfeatures() == int(g["nfeatures"])
    np.random.seed(0)
    fruit.fit(X)
    res = fruit.transform(X)
    if name in ("R_preps", "R_cosrand"):
        # preparateurs in front (fastmath loops of the reference: 1e-12) and a weighted slice
        assert_close(fitted_thresholds(fruit), g["thresholds"], 1e-9, "thresholds")
        _assert_features_close(res, g["features"], name)
    else:
        assert_exact(fitted_thresholds(fruit), g["thresholds"], "thresholds")
        if name == "R_letters":
            assert_close(res, g["features"], 1e-12, "features")      # MPI: summation order
        elif name == "R_rng":
            assert_exact(res, g["features"], "features")
        else:
            assert_close(res, g["features"], 1e-12, "features")      # CUR: summation order
    if str(g["labels"]) == "IndexError":
        # Arctic(argmax=True): the reference's ISS._label asks the cache plan for rows it
        # does not hold (fruits/iss/iss.py:195-198); same here
        with pytest.raises(IndexError):
            fruit.label(res.shape[1] - 1)
    else:
        labels = "|".join(fruit.label(i) for i in
                          sorted(set(np.linspace(0, res.shape[1] - 1, 23).astype(int))))
        assert labels == str(g["labels"])
    assert fruit.summary() == str(g["summary"])


@pytest.mark.parametrize("semiring", ["reals", "arctic"])
@pytest.mark.parametrize("shape", [(37, 2, 1), (70, 2, 2), (64, 2, 33), (300, 2, 129)])
def test_rank2_sieves_fused_equal_composed_and_oracle(semiring, shape, monkeypatch):
    """XPI / LPI / CUR / CPV as accumulators of the generated kernel (the ISS tensor
    stays in registers) against the composed route (materialise + stand-alone sieve
    kernels, the only route they had before) and against the oracle."""
    from oracle import pipeline as orc
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": {"of_weight": [3, 2]}, "mode": "extended",
                                 "semiring": semiring}],
                        "sieves": [["NPI", {"q": [0.4, 1.0]}], ["XPI", {"q": [0.4, 1.0]}],
                                   ["LPI", {"inc": 2, "q": [0.3, 1.0]}], ["XPI", {"inc": 0, "q": [0.2, 1.0]}],
                                   ["LPI", {"inc": 0, "q": [0.2, 1.0]}], ["CUR", {"q": [-1.0, 0.7]}],
                                   ["CPV", {}], ["PPV", {}], ["MPI", {"inc": 2, "q": [0.3, 1.0]}],
                                   ["END", {}]],
                        "fit_sample_size": 1.0}]}
    X = np.random.default_rng(shape[0] + shape[2]).standard_normal(shape).cumsum(axis=2)
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(5)
    fruit.fit(X)
    monkeypatch.setenv("FRUITS_B200_JIT", "force")
    monkeypatch.setenv("FRUITS_B200_CHAIN", "0")
    res = fruit.transform(X)
    assert _routes(fruit) == ["fb_jit_slice"]
    monkeypatch.setenv("FRUITS_B200_JIT", "0")
    comp = fruit.transform(X)
    assert _routes(fruit) == ["composed"]
    of = orc.OracleFruit(spec)
    np.random.seed(5)
    of.fit(X)
    ref = of.transform(X)
    nf = 10
    summed = np.zeros(ref.shape[1], dtype=bool)
    for f in (5, 8):                           # CUR, MPI: sums (order of the additions)
        summed[f::nf] = True
    assert_exact(res[:, ~summed], comp[:, ~summed], "fused vs composed route")
    assert_exact(res[:, ~summed], ref[:, ~summed], "fused vs oracle")
    assert_close(res[:, summed], comp[:, summed], 1e-12, "CUR / MPI vs composed")
    assert_close(res[:, summed], ref[:, summed], 1e-12, "CUR / MPI vs oracle")


@pytest.mark.parametrize("shape", [(70, 2, 96), (33, 3, 41), (1, 2, 2)])
def test_bayesian_fused_equals_composed_and_oracle(shape, monkeypatch):
    """Unweighted Bayesian (max, times) sums compiled into the generated kernel (the
    ISS tensor stays in registers: fruits/iss/semiring.py:461-494, one multiplication /
    division per letter occurrence, then the running maximum) against the composed
    route (fb_bayes_word block scan + stand-alone sieves) and the oracle."""
    from oracle import pipeline as orc
    spec = {"slices": [{"iss": [{"words": ["[1][-2][2]", "[11][2]", "[2][22][1][12]", "[2][1]"],
                                 "mode": "extended", "semiring": "bayesian"}],
                        "sieves": [["NPI", {"q": [0.4, 1.0]}], ["PPV", {}], ["MAX", {}], ["MIN", {}],
                                   ["LPI", {"inc": 0, "q": [0.2, 1.0]}], ["CPV", {}], ["END", {}]],
                        "fit_sample_size": 1.0}]}
    X = 0.25 + 1.5 * np.random.default_rng(shape[0]).random(shape)
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(5)
    fruit.fit(X)
    monkeypatch.setenv("FRUITS_B200_JIT", "force")
    res = fruit.transform(X)
    assert _routes(fruit) == ["fb_jit_slice"]
    monkeypatch.setenv("FRUITS_B200_JIT", "0")
    comp = fruit.transform(X)
    assert _routes(fruit) == ["composed"]
    of = orc.OracleFruit(spec)
    np.random.seed(5)
    of.fit(X)
    assert_exact(fitted_thresholds(fruit), oracle_thresholds(of), "thresholds")
    assert_exact(res, comp, "fused vs composed route")
    assert_exact(res, of.transform(X), "fused vs oracle")


@pytest.mark.parametrize("semiring", ["reals", "arctic"])
@pytest.mark.parametrize("cut", [0.4, 0.999, 0.001, 30, 1, -3])
def test_single_cut_fused_equals_composed_and_oracle(cut, semiring, monkeypatch):
    """One int or float ("coquantile") cut shared by the segment sieves of a slice
    (fruits/sieving/segment.py:51-64, fruits/cache.py:16-22): the generated kernel
    gets the end of the segment per series as a table and sieves [0, cut) only --
    no materialised ISS tensor -- with the numbers of the composed route and of
    the oracle, END included (the value at cut - 1)."""
    from oracle import pipeline as orc
    sieves = [["NPI", {"cut": cut, "q": [0.4, 1.0]}], ["MPI", {"cut": cut, "q": [0.4, 1.0]}],
              ["XPI", {"cut": cut, "inc": 2, "q": [0.5, 1.0]}], ["LPI", {"cut": cut, "inc": 0, "q": [0.3, 1.0]}],
              ["MAX", {"cut": cut}], ["MIN", {"cut": cut, "q": [0.2, 0.9]}], ["CUR", {"cut": cut}],
              ["PPV", {}], ["CPV", {}], ["END", {"cut": cut}]]
    spec = {"slices": [{"preps": [["INC", {}]],
                        "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended",
                                 "semiring": semiring}],
                        "sieves": sieves, "fit_sample_size": 1.0}]}
    X = np.random.default_rng(17).standard_normal((150, 2, 97)).cumsum(axis=2)
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(5)
    fruit.fit(X)
    monkeypatch.setenv("FRUITS_B200_JIT", "force")
    res = fruit.transform(X)
    assert _routes(fruit) == ["fb_jit_slice"]
    monkeypatch.setenv("FRUITS_B200_JIT", "0")
    comp = fruit.transform(X)
    assert _routes(fruit) == ["composed"]
    of = orc.OracleFruit(spec)
    np.random.seed(5)
    of.fit(X)
    ref = of.transform(X)
    nf = len(sieves)
    summed = np.zeros(ref.shape[1], dtype=bool)
    for f in (1, 6):                           # MPI, CUR: sums (order of the additions)
        summed[f::nf] = True
    assert_exact(res[:, ~summed], comp[:, ~summed], "fused vs composed route")
    assert_exact(res[:, ~summed], ref[:, ~summed], "fused vs oracle")
    assert_close(res[:, summed], comp[:, summed], 1e-12, "MPI / CUR vs composed")
    assert_close(res[:, summed], ref[:, summed], 1e-12, "MPI / CUR vs oracle")
    # sieves with different cuts do not share a kernel: composed route
    mixed = specs.build_fruit(fruits, {"slices": [dict(spec["slices"][0], sieves=sieves[:1] + [["END", {}]])]})
    assert mixed.get_slice(0)._fused_sieves() is None


def test_chain_kernel_long_series_fewer_series_per_cta(monkeypatch):
    """The chain kernel stages whole series in shared memory; for long series it is
    regenerated with fewer series per CTA before the generic kernel has to take
    over: same numbers as the generic kernel."""
    spec = {"slices": [dict(specs.SPECS["C4_twi"]["slices"][1], fit_sample_size=1)]}
    X = np.random.default_rng(4).standard_normal((40, 1, 9000)).cumsum(axis=2)
    fruit = specs.build_fruit(fruits, spec)
    np.random.seed(0)
    fruit.fit(X)
    monkeypatch.setenv("FRUITS_B200_JIT", "force")
    res = fruit.transform(X)
    route, _, kern = fruit.get_slice(0)._last_launch
    assert route == "fb_jit_chain" and kern.em.spc < 4 and kern.fits(9000)
    monkeypatch.setenv("FRUITS_B200_JIT", "0")
    gen = fruit.transform(X)
    assert _routes(fruit) == ["fb::lns_kernel"]
    assert_exact(res, gen, "chain kernel (fewer series per CTA) vs generic kernel")


# ---------------------------------------------------------------------------
# (k) BASELINE sizes against golden vectors frozen from the REAL reference
# (oracle/gen_golden_full.py: the reference's own fit on the full input, its
# transform of all / of sampled rows)

def _slice_feature_bounds(fruit):
    b = [0]
    for slc in fruit:
        b.append(b[-1] + slc.nfeatures())
    return b


def test_c2_full_size_against_the_reference(golden_dir):
    """experiments/fruit_reduced.py, all four slices, at BASELINE size
    (1,000 x 1 x 512): every threshold of the full fit, the features of 64 rows,
    the NPI counts of ALL rows and the arctic slice of all rows bit for bit."""
    import hashlib
    g = np.load(os.path.join(golden_dir, "full_C2_full.npz"))
    X = specs.make_input("C2_full")
    assert X.shape[0] == int(g["n"]) == 1000
    fruit = specs.build_fruit(fruits, specs.SPECS["C2_full"])
    np.random.seed(0)
    fruit.fit(X)
    thr = fitted_thresholds(fruit)
    tb = g["thr_slices"]
    assert thr.shape == g["thresholds"].shape
    # slice 1 (arctic, unweighted): bit-identical thresholds; the others: device exp / sin / cos
    assert_exact(thr[tb[1]:tb[2]], g["thresholds"][tb[1]:tb[2]], "C2 full arctic thresholds")
    assert_close(thr, g["thresholds"], 1e-9, "C2 full thresholds")
    res = fruit.transform(X)
    assert res.shape == (1000, int(g["nfeatures"])) == (1000, 4431)
    fb = _slice_feature_bounds(fruit)
    rows = g["rows"]
    exact = g["exact_cols_slice1"]         # arctic slice without the MPI means (summation order)
    assert fb[1] <= exact.min() and exact.max() < fb[2] and exact.size == 4 * 188
    assert_exact(res[rows][:, exact], g["features"][:, exact], "C2 full arctic rows")
    assert_close(res[rows][:, fb[1]:fb[2]], g["features"][:, fb[1]:fb[2]], 1e-12, "C2 arctic MPI")
    _assert_features_close(res[rows], g["features"], "C2_full 64 rows vs reference")
    # all 1,000 rows: integer counts, and one hash per row of the bit-exact slice
    cnt = res[:, g["count_cols"]]
    rep = parity_report(cnt, g["counts"].astype(np.float64))
    record_report("C2_full NPI counts of all rows vs reference", rep)
    assert rep["rtol_violations"] == rep["count_flips"] <= 2e-4 * cnt.size, rep
    sha = [hashlib.sha256(np.ascontiguousarray(res[j, exact] + 0.0).tobytes()).hexdigest()[:16]
           for j in range(res.shape[0])]
    assert sha == [str(x) for x in g["row_sha_slice1"]]


def test_c3_full_size_against_the_reference(golden_dir):
    """experiments/fruit_general.py, all four slices, at BASELINE size
    (10,000 x 6 x 1,024): all 40,334 thresholds of the reference's own full fit
    (102 minutes of numba on 8 cores; 1,731 + 1,150 iterated sums, 10.2 M values
    per quantile, the arctic slice with its buckets of equal increments through
    the eight-pass fallback of the select) and the features of 16 sampled rows."""
    g = np.load(os.path.join(golden_dir, "full_C3_full.npz"))
    X = specs.make_input("C3_full")
    assert X.shape == (10000, 6, 1024) and int(g["n"]) == 10000
    fruit = specs.build_fruit(fruits, specs.SPECS["C3_full"])
    np.random.seed(0)
    fruit.fit(torch.from_numpy(X).cuda())
    thr = fitted_thresholds(fruit)
    tb = g["thr_slices"]
    assert thr.shape == g["thresholds"].shape == (40334,)
    assert_exact(thr[tb[1]:tb[2]], g["thresholds"][tb[1]:tb[2]], "C3 full arctic thresholds")
    for si in range(4):
        a, b = thr[tb[si]:tb[si + 1]], g["thresholds"][tb[si]:tb[si + 1]]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        rel = np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)
        record_report(f"C3_full thresholds slice {si}",
                      {"thresholds": int(fin.sum()), "bit_identical": int((a[fin] == b[fin]).sum()),
                       "max_rel_err": float(rel.max(initial=0.0))})
    assert_close(thr, g["thresholds"], 1e-9, "C3 full thresholds")
    rows = g["rows"]
    res = fruit.transform(np.ascontiguousarray(X[rows]))
    assert res.shape == g["features"].shape == (16, 20167)
    fb = _slice_feature_bounds(fruit)
    exact = np.array([c for c in range(fb[1], fb[2]) if (c - fb[1]) % 7 not in (3, 4, 5)])
    assert_exact(res[:, exact], g["features"][:, exact], "C3 full arctic rows (counts, END)")
    assert_close(res[:, fb[1]:fb[2]], g["features"][:, fb[1]:fb[2]], 1e-12, "C3 arctic MPI")
    _assert_features_close(res, g["features"], "C3_full 16 rows vs reference")


# ---------------------------------------------------------------------------
# (l) row-sharded fit on one GPU: the phases of the distributed selection with
# the ranks emulated one after the other (their workspaces reduced by hand), and
# parallel.fit_sharded in a process group of one

def _emulated_reduce(works, off, count, dtype, op):
    width = 4 if dtype == torch.int32 else 8
    regions = [w[off:off + count * width].view(dtype) for w in works]
    total = torch.stack(regions).sum(0) if op == "sum" else torch.stack(regions).min(0).values
    for r in regions:
        r.copy_(total)


@pytest.mark.parametrize("counts", [(700, 300, 0), (1, 0, 0), (512, 512, 513)])
def test_row_sharded_selection_phases(counts):
    """fb_order_stats_dist / fb_order_stats_dist8 with emulated ranks (one of them
    without rows) equal np.quantile over the concatenated rows; ties and a
    bucket of equal values larger than the candidate list included."""
    import ctypes
    from fruits_b200 import _backend as be
    from fruits_b200.sieving.abstract import _lerp, _virtual_index
    L = be.lib()
    t, P = 64, 5
    rng = np.random.default_rng(sum(counts))
    n = sum(counts)
    Y = rng.standard_normal((P, n, t)).cumsum(axis=2)
    Y[1] = np.round(Y[1])                        # many ties
    Y[2, :, 10:] = Y[2, :, 9:10]                 # increments mostly exactly 0
    bounds = np.cumsum((0,) + counts)
    shards = [torch.from_numpy(np.ascontiguousarray(Y[:, bounds[r]:bounds[r + 1]])).cuda()
              .reshape(P, -1) for r in range(len(counts))]
    M = n * t
    pairs = [(0, 0.5), (1, 0.3), (2, 0.9), (1, 0.5)]
    S = len(pairs)
    kg = [_virtual_index(M, q) for _, q in pairs]
    incs = (ctypes.c_int32 * S)(*[i for i, _ in pairs])
    ks = (ctypes.c_int64 * S)(*[k for k, _ in kg])
    lay = (ctypes.c_int64 * 5)()
    be.check(L.fb_order_stats_dist_layout(P, S, lay))
    works = [torch.zeros((lay[4],), dtype=torch.uint8, device="cuda") for _ in counts]
    outs = [(be.empty((P, S)), be.empty((P, S)), be.empty((P, S), dtype=torch.int32)) for _ in counts]
    ns = P * S
    for phase in range(10):
        for V, w, (lo, hi, done) in zip(shards, works, outs):
            be.check(L.fb_order_stats_dist(phase, V.data_ptr() if V.numel() else 0, V.shape[1], P,
                                           V.shape[1], M, t, S, incs, ks, lo.data_ptr(),
                                           hi.data_ptr(), done.data_ptr(), w.data_ptr(),
                                           be.stream_ptr()))
        if phase in (0, 1):
            _emulated_reduce(works, lay[0], ns * 4096, torch.int32, "sum")
        elif phase in (2, 8):
            _emulated_reduce(works, lay[2], ns * 2, torch.int64, "sum")
            _emulated_reduce(works, lay[3], ns * 4, torch.int64, "min")
        elif 3 <= phase <= 7:
            _emulated_reduce(works, lay[1], ns * 256, torch.int32, "sum")

    def increments(a, depth):
        for _ in range(depth):
            a = np.concatenate([np.zeros_like(a[..., :1]), np.diff(a, axis=-1)], axis=-1)
        return a

    lo0, hi0, done0 = (x.cpu().numpy() for x in outs[0])
    for r in range(1, len(counts)):              # every rank ends with the same answer
        for a, b in zip(outs[0], outs[r]):
            assert_exact(a.cpu().numpy(), b.cpu().numpy(), "ranks disagree")
    for s, ((inc, q), (k, gamma)) in enumerate(zip(pairs, kg)):
        want = np.array([np.quantile(increments(Y[p], inc).ravel(), q) for p in range(P)])
        got = _lerp(lo0[:, s], hi0[:, s] if k < M - 1 else lo0[:, s], gamma)
        ok = done0[:, s].astype(bool)
        assert_exact(got[ok], want[ok], f"dist select inc={inc} q={q}")
        assert ok[0] and ok[3] and ok[4]         # (rows 1-2 may exceed the candidate list;
        #                                          one value many times is resolved anyway)
    # the eight-pass form on materialised values
    for inc, q in ((1, 0.5), (0, 0.25)):
        vals = [torch.from_numpy(np.ascontiguousarray(
            increments(Y[:, bounds[r]:bounds[r + 1]], inc))).cuda().reshape(P, -1)
            for r in range(len(counts))]
        k, gamma = _virtual_index(M, q)
        lay8 = (ctypes.c_int64 * 4)()
        be.check(L.fb_order_stats_dist8_layout(P, lay8))
        works8 = [torch.zeros((lay8[3],), dtype=torch.uint8, device="cuda") for _ in counts]
        outs8 = [(be.empty((P,)), be.empty((P,))) for _ in counts]
        for phase in range(10):
            for V, w, (lo, hi) in zip(vals, works8, outs8):
                be.check(L.fb_order_stats_dist8(phase, V.data_ptr() if V.numel() else 0,
                                                V.shape[1], P, V.shape[1], M, k, lo.data_ptr(),
                                                hi.data_ptr(), w.data_ptr(), be.stream_ptr()))
            if phase <= 7:
                _emulated_reduce(works8, lay8[0], P * 256, torch.int32, "sum")
            elif phase == 8:
                _emulated_reduce(works8, lay8[1], P * 2, torch.int64, "sum")
                _emulated_reduce(works8, lay8[2], P, torch.int64, "min")
        a, b = outs8[0][0].cpu().numpy(), outs8[0][1].cpu().numpy()
        want = np.array([np.quantile(increments(Y[p], inc).ravel(), q) for p in range(P)])
        assert_exact(_lerp(a, b if k < M - 1 else a, gamma), want, f"dist8 inc={inc} q={q}")


def test_fit_sharded_rows_in_a_group_of_one():
    """parallel.fit_sharded(shard="rows") with world size 1: the whole row-sharded
    path (sample positions, sharded raw-input cache, phased selections) on one
    GPU equals ``Fruit.fit``."""
    import torch.distributed as dist
    from fruits_b200.parallel import fit_sharded
    spec = {"slices": [
        {"preps": [["INC", {}]],
         "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended", "weighting": ["L1", {}]}],
         "sieves": [["NPI", {"q": [0.4, 1.0]}], ["MPI", {"q": [0.1, 0.9], "inc": 0}],
                    ["PPV", {"sample_size": 0.5}], ["END", {}]], "fit_sample_size": 0.5},
        {"iss": [{"words": {"alternate_sign": [6 * "[1]", 3 * "[1][2]"]}, "mode": "extended",
                  "semiring": "arctic"}],
         "sieves": [["NPI", {"q": [0.5, 1.0], "inc": i}] for i in range(3)] + [["END", {}]],
         "fit_sample_size": 1.0},
        {"iss": [{"words": ["[1]", "[12]"], "mode": "extended"}],
         "sieves": [["NPI", {"q": [0.5, 1.0]}], ["PPV", {}]], "fit_sample_size": 1}]}
    X = torch.from_numpy(np.random.default_rng(8).standard_normal((301, 2, 120)).cumsum(axis=2)).cuda()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1)
    try:
        fruit = specs.build_fruit(fruits, spec)
        np.random.seed(40)
        fit_sharded(fruit, X, shard="rows")
        after = np.random.random()
    finally:
        dist.destroy_process_group()
    single = specs.build_fruit(fruits, spec)
    np.random.seed(40)
    single.fit(X)
    assert after == np.random.random()
    assert_exact(fitted_thresholds(fruit), fitted_thresholds(single), "row-sharded fit thresholds")
    assert torch.equal(fruit.transform_device(X), single.transform_device(X))
