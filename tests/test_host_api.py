"""CPU tests of the host-side mirror of the reference API: word algebra, cache
plan, trie compilation, labels / summary, error behaviour, and that the C ABI
library loads and exports every symbol declared in include/fruits_b200.h
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

import fruits_b200 as fruits
import specs
from fruits_b200 import _backend as be
from fruits_b200._plan import Trie

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_word_enumeration_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "words.npz"))
    for key in g.files:
        if key.startswith("of_weight_"):
            _, _, w, d = key.split("_")
            words = fruits.words.of_weight(int(w), int(d))
            assert "|".join(str(x) for x in words) == str(g[key])
            assert fruits.iss.CachePlan(words)._plan == list(g[f"plan_{w}_{d}"])
    base = [24 * "[1]", 24 * "[2]", 12 * "[1][2]", 12 * "[2][1]", "[112][2][1]"]
    alt = fruits.words.alternate_sign([fruits.words.SimpleWord(b) for b in base])
    assert "|".join(str(x) for x in alt) == str(g["alternate_sign"])
    assert fruits.iss.CachePlan(alt)._plan == list(g["alternate_sign_plan"])
    for i in range(4):
        s = str(g[f"parse_{i}_str"])
        assert np.array_equal(fruits.words.SimpleWord(s).exponents(), g[f"parse_{i}"])


def test_word_counts():
    # reference: tests/signature/test_simple.py:54-57
    for n in range(1, 7):
        assert len(fruits.words.of_weight(n, dim=1)) == 2 ** (n - 1)
    assert len(fruits.words.of_weight(4, dim=2)) == 82


def test_cache_plan_reference_golden():
    # reference: tests/signature/test_cache.py:11-26
    words = [fruits.words.SimpleWord(s) for s in [
        "[1][11][3][11]", "[11][13][11][1][3]", "[1][13][1]", "[11][13][111][13][11]",
        "[3][11][111]", "[1][11][2]", "[11][2]", "[11][13][111][13][2]",
        "[3][11][1112][21]"]]
    plan = fruits.iss.CachePlan(words)
    assert plan._plan == [4, 5, 2, 3, 3, 1, 1, 1, 2]
    assert plan.n_iterated_sums() == 22
    assert plan.get_word_string(0) == "[1]"
    assert plan.get_word_string(4) == "[11]"
    assert plan.get_word_index(4) == 1


def test_simpleword_behaviour():
    w = fruits.words.SimpleWord("[12][122]")
    assert list(w) == [[1, 1], [1, 2]]
    assert w == fruits.words.SimpleWord("[21][212]")
    assert str(w) == "[12][122]" and len(w) == 2
    assert np.array_equal(w.alpha, np.ones(2, dtype=np.float32))
    w.alpha = [0.5, 2]
    assert w.alpha.dtype == np.float32
    with pytest.raises(ValueError):
        w.alpha = [1.0]
    with pytest.raises(ValueError):
        fruits.words.SimpleWord("[1][a]")
    w.multiply("[3]")
    assert list(w) == [[1, 1, 0], [1, 2, 0], [0, 0, 1]]
    c = w.copy()
    assert c == w and c is not w
    neg = fruits.words.SimpleWord("[-1-12][(-11)3]")
    assert neg.exponents()[0, 0] == -2 and neg.exponents()[1, 10] == -1


@pytest.mark.parametrize("name,slice_idx,n_nodes", [
    ("C1_readme", 0, 18), ("C2_reduced", 0, 115), ("C2_reduced", 1, 188),
    ("C3_general", 0, 1351), ("C3_general", 1, 380), ("C4_twi", 0, 511),
    ("C4_twi", 1, 96), ("C5_sweep", 0, 445)])
def test_trie_sizes(name, slice_idx, n_nodes):
    # emitted iterated sums per config
    fruit = specs.build_fruit(fruits, specs.SPECS[name])
    iss = fruit.get_slice(slice_idx).get_iss()[0]
    assert iss.n_iterated_sums() == n_nodes
    trie = Trie(iss.words, iss._cache_plan._plan, iss.weighting is not None)
    assert len(trie.emits) == n_nodes
    # every prefix is computed once: no more nodes than emissions here
    assert len(trie.nodes) == n_nodes
    # emission order = words in order, each its new prefixes shortest first
    seen, e = set(), 0
    for w in iss.words:
        letters = str(w).split("]")[:-1]
        for k in range(len(letters)):
            prefix = "]".join(letters[:k + 1]) + "]"
            if prefix not in seen:
                seen.add(prefix)
                node = trie.nodes[trie.emits[e]]
                assert node.depth == k + 1
                e += 1
    assert e == n_nodes


def test_trie_duplicate_emission_and_single_mode():
    words = [fruits.words.SimpleWord(s) for s in ["[12]", "[21]", "[1][2]", "[1]"]]
    trie = Trie(words, None, False)           # SINGLE: one emission per word
    assert len(trie.emits) == 4
    assert len({id(trie.nodes[e]) for e in trie.emits}) == 4
    sub = trie.subset(2, 4)
    assert len(sub.emits) == 2 and all(e is not None for e in sub.emits)


def test_feature_counts_and_labels(golden_dir):
    for name in ("C1_readme", "C2_reduced", "C3_general", "C4_twi", "C5_sweep"):
        g = np.load(os.path.join(golden_dir, f"pipeline_{name}.npz"))
        fruit = specs.build_fruit(fruits, specs.SPECS[name])
        nf = int(g["nfeatures"])
        assert fruit.nfeatures() == nf
        idx = sorted(set(np.linspace(0, nf - 1, 23).astype(int)))
        assert "|".join(fruit.label(i) for i in idx) == str(g["labels"])
        assert fruit.summary() == str(g["summary"])
    # SURVEY.md appendix B
    fruit = specs.build_fruit(fruits, specs.SPECS["C1_readme"])
    assert fruit.label(0) == "INC | [11] | NPI[inc=1]!-1![0.5, 1.0]"
    assert fruit.label(36) == "[11] | NPI[inc=1]!-1![0.0, 1.0]"


def test_reference_feature_counts():
    # reference: tests/core/test_fruit.py:11-41 (738 features)
    fruit = fruits.Fruit()
    fruit.add(fruits.preparation.INC())
    fruit.add(fruits.ISS(fruits.words.of_weight(4, dim=2), mode=fruits.ISSMode.EXTENDED))
    fruit.add(fruits.sieving.NPI(q=(0.5, 1.0)), fruits.sieving.MPI(q=(0.5, 1.0)))
    fruit.add(fruits.sieving.MAX, fruits.sieving.MIN, fruits.sieving.END,
              fruits.sieving.PPV(quantile=[0.2, 0.5], constant=False))
    assert fruit.nfeatures() == 115 * 7


def test_error_behaviour():
    fruit = fruits.Fruit()
    with pytest.raises(RuntimeError, match="Missing call of self.fit"):
        fruit.transform(np.zeros((2, 1, 8)))
    fruit.cut()
    with pytest.raises(TypeError):
        fruit.add(3)
    with pytest.raises(IndexError):
        fruit.switch_slice(5)
    slc = fruits.FruitSlice()
    with pytest.raises(RuntimeError, match="No ISS given"):
        slc._compile()
    slc.add(fruits.ISS([fruits.words.SimpleWord("[1]")]))
    with pytest.raises(RuntimeError, match="No feature sieves given"):
        slc._compile()
    iss = fruits.ISS(fruits.words.of_weight(2, 1))
    with pytest.raises(ValueError):
        next(iss.batch_transform(np.zeros((1, 1, 4)), batch_size=10))
    argmax = fruits.ISS([fruits.words.SimpleWord("[1][2]")],
                        semiring=fruits.semiring.Arctic(argmax=True))
    with pytest.raises(NotImplementedError, match="ISSMode.SINGLE"):
        argmax.n_iterated_sums()
    assert fruits.ISS([fruits.words.SimpleWord("[1][2]"), fruits.words.SimpleWord("[1][2][1]")],
                      mode=fruits.ISSMode.EXTENDED,
                      semiring=fruits.semiring.Arctic(argmax=True)).n_iterated_sums() == 5 + 9
    with pytest.raises(NotImplementedError):
        fruits.ISS([fruits.words.Word()])                      # no extended letter
    with pytest.raises(TypeError):
        fruits.ISS(["[1]"])
    with pytest.raises(NotImplementedError):
        fruits.ISS([fruits.words.Word("[DIM(1)]")], mode=fruits.ISSMode.EXTENDED,
                   semiring=fruits.semiring.Arctic(argmax=True))
    with pytest.raises(ValueError):
        fruits.preparation.MAV(width=1.5)
    with pytest.raises(ValueError):
        fruits.preparation.JLD(dim=1.5)
    with pytest.raises(ValueError):
        fruits.preparation.PDD(density=2.0)
    with pytest.raises(TypeError):
        fruits.preparation.DOT(n="3")
    with pytest.raises(ValueError):
        fruits.preparation.INC(depth=0)
    with pytest.raises(ValueError):
        fruits.iss.weighting.Plateaus(1)


def test_fusability_rules():
    for name in ("C1_readme", "C2_reduced", "C3_general", "C4_twi", "C5_sweep"):
        fruit = specs.build_fruit(fruits, specs.SPECS[name])
        for slc in fruit:
            assert slc._is_fusable(6, None), name
    slc = fruits.FruitSlice()
    slc.add(fruits.ISS(fruits.words.of_weight(2, 1)), fruits.sieving.NPI(cut=(10, -1)))
    assert not slc._is_fusable(1, None)      # several cuts -> composed route
    slc = fruits.FruitSlice()
    slc.add(fruits.preparation.INC(depth=2), fruits.ISS(fruits.words.of_weight(2, 1)),
            fruits.sieving.END)
    assert not slc._is_fusable(1, None)      # INC depth 2 is not fused on load
    slc = fruits.FruitSlice()
    slc.add(fruits.preparation.NEW(fruits.preparation.INC()), fruits.preparation.STD)
    assert slc._fused_dims(2) == [(0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 1)]


def test_drop_in_alias():
    import fruits as alias
    assert alias.Fruit is fruits.Fruit
    assert alias.iss.weighting.Indices is fruits.iss.weighting.Indices
    assert alias.semiring.Arctic is fruits.semiring.Arctic
    assert alias.words.of_weight is fruits.words.of_weight
    assert alias.sieving.NPI is fruits.sieving.NPI


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fruits_b200.h")).read()
    declared = set(re.findall(r"FB_API\s+[\w\s\*]+?\b(fb_\w+)\s*\(", header))
    assert declared == set(be.EXPORTED), declared ^ set(be.EXPORTED)
    lib = ctypes.CDLL(be.LIB_PATH)       # loads without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert be.lib().fb_abi_version() == 1
    assert be.lib().fb_slice_rows(be.POLICY_MAT) >= 1


def test_struct_layouts_match_header():
    assert be.SLOT_DTYPE.itemsize == 16
    assert ctypes.sizeof(be.FbDim) == 16
    assert ctypes.sizeof(be.FbBatch) == 56
    assert ctypes.sizeof(be.FbSievePlan) == 4 + 4 * 16 * 2 + 4 + 8
    assert ctypes.sizeof(be.FbIssPlan) == 8 * 4 + 4 * 4 + 7 * 16 + 3 * 8


# ---------------------------------------------------------------------------
# caller side: UCR .txt loader of the harness

def _write_ucr(root, name, rng, n_train=12, n_test=7, length=40, comma=False, nan=False):
    import os
    os.makedirs(os.path.join(root, name))
    out = {}
    for split, n in (("TRAIN", n_train), ("TEST", n_test)):
        y = rng.integers(1, 3, size=n)
        X = rng.standard_normal((n, length)).cumsum(axis=1) + y[:, None]
        if nan:
            X[0, 0] = np.nan
            X[1, 5:8] = np.nan
            X[2, -1] = np.nan
        raw = np.concatenate([y[:, None].astype(float), X], axis=1)
        np.savetxt(os.path.join(root, name, f"{name}_{split}.txt"), raw,
                   delimiter="," if comma else "  ")
        out[split] = (X, y)
    return out


def test_corbeille_loader(tmp_path):
    import corbeille
    rng = np.random.default_rng(0)
    a = _write_ucr(str(tmp_path), "Alpha", rng)
    b = _write_ucr(str(tmp_path), "Beta", rng, comma=True, nan=True)
    Xtr, ytr, Xte, yte = corbeille.data.load(str(tmp_path / "Alpha") + "/")
    assert Xtr.shape == (12, 1, 40) and Xte.shape == (7, 1, 40)
    assert Xtr.dtype == np.float64 and ytr.dtype == np.int32
    np.testing.assert_allclose(Xtr[:, 0], a["TRAIN"][0])
    np.testing.assert_array_equal(yte, a["TEST"][1])
    # NaNs: forward fill, 0 at the first step (reference data.py:125-147)
    Xtr, _, _, _ = corbeille.data.load(str(tmp_path / "Beta"))
    src = b["TRAIN"][0]
    assert not np.isnan(Xtr).any()
    assert Xtr[0, 0, 0] == 0.0
    np.testing.assert_allclose(Xtr[1, 0, 5:8], src[1, 4])
    np.testing.assert_allclose(Xtr[2, 0, -1], src[2, -2])
    kept = corbeille.data.load(str(tmp_path / "Beta"), keep_nan=True)[0]
    assert np.isnan(kept[1, 0, 5:8]).all()
    assert corbeille.data.replace_nan(kept, 7.0)[1, 0, 6] == 7.0
    names = [d[0] for d in corbeille.data.load_all(str(tmp_path))]
    assert names == ["Alpha", "Beta"]
    assert [d[0] for d in corbeille.data.load_all(str(tmp_path), datasets=["Beta"])] == ["Beta"]
    with pytest.raises(FileNotFoundError):
        corbeille.data.load(str(tmp_path / "Alpha"), univariate=False)      # no .arff there


def test_corbeille_data_module_equals_the_reference(golden_dir, tmp_path):
    """The multivariate .arff reader (+ its .npy cache), multisine and the
    resampling helpers against what the reference's module returns for the same
    files / seeds (frozen by ``oracle/gen_golden.py corbeille2``)."""
    import shutil
    import corbeille
    g = np.load(os.path.join(golden_dir, "corbeille2.npz"))
    root = str(tmp_path / "Zeta")
    shutil.copytree(os.path.join(golden_dir, "ucr_mv", "Zeta"), root)
    for keep, tag in ((False, ""), (True, "_keep_nan")):
        got = corbeille.data.load(root, univariate=False, cache=False, keep_nan=keep)
        for key, a in zip(("X_train", "y_train", "X_test", "y_test"), got):
            want = g[f"arff_{key}{tag}"]
            assert a.dtype == want.dtype and np.array_equal(a, want, equal_nan=True), key
    assert not [f for f in os.listdir(root) if f.endswith(".npy")]
    first = corbeille.data.load(root, univariate=False)                # writes the cache
    assert sorted(f for f in os.listdir(root) if f.endswith(".npy")) == [
        "Zeta_XTEST.npy", "Zeta_XTRAIN.npy", "Zeta_yTEST.npy", "Zeta_yTRAIN.npy"]
    os.remove(os.path.join(root, "Zeta_TRAIN.arff"))                  # ... and reads it back
    for a, b in zip(first, corbeille.data.load(root, univariate=False)):
        assert np.array_equal(a, b)
    assert [d[0] for d in corbeille.data.load_all(str(tmp_path), univariate=False)] == ["Zeta"]
    X = g["util_X"]
    assert np.array_equal(corbeille.data.lengthen(X, 0.2), g["lengthen"])
    assert np.array_equal(corbeille.data.downsample(X, 0.34), g["downsample"])
    assert np.array_equal(corbeille.data.upsample(X), g["upsample"])
    assert np.array_equal(corbeille.data.upsample(X[:, :1]), g["upsample_1d"])
    mid = corbeille.data.upsample(X[:, :1])
    assert np.array_equal(mid[:, :, 0::2], X[:, :1]) and mid.shape[2] == 59
    for tag, sl in (("a", 0.1), ("b", 0.5)):
        np.random.seed(5)
        assert np.array_equal(corbeille.data.implant_stuttering(X, sl), g["stutter_" + tag])
        assert np.random.random() == float(g[f"stutter_{tag}_rng"])
    np.random.seed(9)
    ms = corbeille.data.multisine(train_size=11, test_size=7, length=20, n_classes=3)
    assert np.random.random() == float(g["multisine_rng"])
    for key, a in zip(("X_train", "y_train", "X_test", "y_test"), ms):
        assert np.array_equal(a, g["multisine_" + key]), key
    sized = corbeille.data.multisine(train_size=4, test_size=5, length=9, noise=lambda: 0.25)
    assert sized[0].shape == (4, 1, 9) and sorted(sized[3]) == [0, 0, 1, 1, 1]


def test_corbeille_split_index_equals_the_reference(golden_dir):
    """tools.split_index over every prepared / iterated-sum / feature index of
    a fruit with chained ISS and sieves of several features."""
    import corbeille
    g = np.load(os.path.join(golden_dir, "corbeille2.npz"))
    fruit = specs.build_fruit(fruits, specs.SPECS["R_mixed"])
    for level in ("prepared", "iterated sums", "features"):
        want = g["split_" + level.replace(" ", "_")]
        got = [corbeille.tools.split_index(fruit, i, level) for i in range(len(want))]
        assert np.array_equal(np.array(got), want), level
        for bad in (-1, len(want)):
            with pytest.raises(ValueError):
                corbeille.tools.split_index(fruit, bad, level)
    with pytest.raises(ValueError):
        corbeille.tools.split_index(fruit, 0, "words")


def test_corbeille_decide_which_fruit_logic():
    """decide_which_fruit with stand-in fruits (no GPU): the candidate whose
    features classify best on the validation splits wins, a pair returns its
    second member, a class with one sample short-cuts to the first candidate."""
    import corbeille

    class Stub:
        def __init__(self, name, informative):
            self.name, self.informative = name, informative

        def fit(self, X):
            pass

        def transform(self, X):
            signal = X[:, 0, :1] if self.informative else np.zeros((len(X), 1))
            return np.concatenate((signal, np.ones((len(X), 1))), axis=1)

        def deepcopy(self):
            return Stub(self.name + "'", self.informative)

    rng = np.random.default_rng(0)
    y = np.repeat([0, 1], 30)
    X = rng.standard_normal((60, 1, 8)) * 0.1
    X[:, 0, 0] += 3.0 * y
    blind, sharp, big = Stub("blind", False), Stub("sharp", True), Stub("big", True)
    np.random.seed(0)
    assert corbeille.decide_which_fruit([blind, sharp], n_splits=3)(X, y).name == "sharp'"
    np.random.seed(0)
    assert corbeille.decide_which_fruit([blind, (sharp, big)])(X, y).name == "big'"
    lonely = y.copy()
    lonely[0] = 5
    assert corbeille.decide_which_fruit([(blind, big), sharp])(X, lonely).name == "big'"


def test_copy_threads_copy_rows():
    """Host staging helper: the row blocks of the copy threads tile the array."""
    from fruits_b200 import _hoststage as hs
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 1000):
        src = rng.random((n, 3, 5))
        dst = np.zeros_like(src)
        hs.finish(hs.copy_rows(dst, src))
        np.testing.assert_array_equal(dst, src)
    with pytest.raises(ValueError):
        hs.finish(hs.copy_rows(np.zeros((4, 2)), np.zeros((4, 3))))


def test_emission_ranges_respect_the_dimension_limit():
    """Words over more dimensions than one launch stages are materialised in
    consecutive emission ranges, each within the limit (ISS._dim_pieces)."""
    from fruits_b200 import _backend as be
    iss = fruits.ISS(fruits.words.of_weight(2, 12), mode=fruits.ISSMode.EXTENDED)
    trie = iss.trie()
    pieces = iss._dim_pieces(None)
    assert len(pieces) > 1 and pieces[0][0] == 0 and pieces[-1][1] == len(trie.emits)
    assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
    for lo, hi in pieces:
        assert len(trie.subset(lo, hi).used_dims()) <= be.FB_MAX_USED_DIMS
    sub = iss._dim_pieces((3, 40))
    assert sub[0][0] == 3 and sub[-1][1] == 40
    small = fruits.ISS(fruits.words.of_weight(3, 3), mode=fruits.ISSMode.EXTENDED)
    assert small._dim_pieces(None) == [None] and small._dim_pieces((2, 9)) == [(2, 9)]
    wide = fruits.ISS([fruits.words.SimpleWord("[12345678]")])
    with pytest.raises(NotImplementedError):
        wide._dim_pieces(None)


@pytest.mark.parametrize("script,n_slices,n_features", [("fruit_reduced", 4, 4431),
                                                        ("fruit_general", 4, 20167),
                                                        ("fruit_twi", 2, 1725)])
def test_reference_experiment_scripts_build_unchanged(script, n_slices, n_features):
    """The reference's own experiment definitions (experiments/fruit_*.py) run
    unmodified against the ``fruits`` alias of this repository and describe the
    same feature space.  Needs the reference checkout (build container only)."""
    import importlib.util
    import os
    path = os.path.join(os.environ.get("FRUITS_REF", "/root/reference"), "experiments",
                        script + ".py")
    if not os.path.exists(path):
        pytest.skip("reference checkout not available")
    import fruits as alias
    assert alias.Fruit is fruits.Fruit
    spec = importlib.util.spec_from_file_location("ref_" + script, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert isinstance(mod.fruit, fruits.Fruit)
    assert len(mod.fruit) == n_slices and mod.fruit.nfeatures() == n_features


def test_corbeille_loader_equals_the_reference(golden_dir):
    """tests/golden/ucr/* read by ``corbeille.data.load`` equals, bit for bit, what
    the reference's loader returns for the same files (frozen in corbeille.npz by
    oracle/gen_golden.py: labels, NaN forward fill, comma and blank separated)."""
    import corbeille
    g = np.load(os.path.join(golden_dir, "corbeille.npz"))
    root = os.path.join(golden_dir, "ucr")
    names = []
    for name, Xtr, ytr, Xte, yte in corbeille.data.load_all(root):
        names.append(name)
        for got, key in ((Xtr, "X_train"), (ytr, "y_train"), (Xte, "X_test"), (yte, "y_test")):
            want = g[f"{name}_{key}"]
            assert got.shape == want.shape and got.dtype == want.dtype, (name, key)
            assert np.array_equal(got, want), (name, key)
        kept = corbeille.data.load(os.path.join(root, name), keep_nan=True)[0]
        assert np.array_equal(kept, g[f"{name}_X_train_keep_nan"], equal_nan=True)
    assert names == ["Delta", "Eps"]
    assert np.isnan(g["Eps_X_train_keep_nan"]).sum() == 5


def test_coswiss_separable_plan_reproduces_the_reference_on_the_host(golden_dir):
    """``CosWISS._separable_plan`` -- the cosine weighted ISS as a recurrence with
    exponent+1 states per level instead of (exponent+1)^(p-1) expansion terms --
    evaluated with numpy (no GPU): equal to the reference's outputs (cos.npz, frozen
    from fruits/iss/cos.py) to 1e-12 of the row maximum on all five cases."""
    from cases import COS_CASES, make_iss_input
    from helpers import rowmax_rel_err
    g = np.load(os.path.join(golden_dir, "cos.npz"))
    for name, (desc, shape, kind) in COS_CASES.items():
        X = make_iss_input(shape, kind)
        iss = specs.build_iss(fruits, desc)
        nodes, emits, rows = iss._separable_plan(shape[1])
        n, d, T = shape
        W = []
        for f, coeff, a, b in rows:
            arg = np.pi * np.arange(T) / (float(np.float32(iss._freqs[f])) * (T - 1))
            w = np.ones(T)
            for _ in range(a):
                w = w * np.sin(arg)
            for _ in range(b):
                w = w * np.cos(arg)
            W.append(coeff * w)
        S = [None] * len(nodes)
        for i, (level, letter, preds) in enumerate(nodes):      # predecessors come first
            lt = np.ones((n, T))
            for dim, div in letter:
                lt = lt / X[:, dim, :] if div else lt * X[:, dim, :]
            if not preds:
                val = lt
            else:
                z = np.zeros((n, T))
                for u, r in preds:
                    if u < 0:
                        z = z + W[r]
                    else:
                        assert u < i
                        prev = np.concatenate([np.zeros((n, 1)), S[u][:, :-1]], axis=1)
                        z = z + W[r] * prev
                val = lt * z
            S[i] = np.cumsum(val, axis=1)
        out = []
        for terms in emits:
            y = np.zeros((n, T))
            for r, u in terms:
                y = y + (S[u] if r < 0 else W[r] * S[u])
            out.append(y)
        out = np.stack(out)
        assert out.shape == g[name].shape
        assert rowmax_rel_err(out, g[name]).max() < 1e-12, name
    # the C3 slice: 1,830 running sums instead of the 15,050 of the expansion
    big = specs.build_iss(fruits, specs.SPECS["C3_cos"]["slices"][1]["iss"][0])
    nodes, emits, rows = big._separable_plan(12)
    assert (len(nodes), len(emits), len(rows)) == (1830, 575, 60)


def test_all_reference_preparateurs_exist_with_reference_strings():
    """Every preparateur name of the reference constructs, copies and prints
    like the reference (fruits/preparation/transform.py, filter.py: ``__str__``
    is what ``Fruit.summary`` and ``label`` show)."""
    P = fruits.preparation
    f = abs
    want = {
        P.INC(): "INC(1, 1, True)", P.STD(): "STD(True, True)", P.NRM(True): "NRM(True)",
        P.MAV(0.1): "MAV(0.1)", P.LAG(): "LAG()", P.FFN(2, 3, False, True): "FFN(2, 3, False, True)",
        P.RIN(2, True, 1, True): "RIN(2, True, 1, True, None)", P.RDW("uniform"): "RDW('uniform')",
        P.JLD(4, True, True): "JLD(4, True, True)",
        P.SPE(0.5, "additive", None, "L1", 9): "SPE(0.5, additive, None, L1, 9)",
        P.RPE(0.3, 7): "RPE(0.3, 7)", P.CTS(0.2, True): "CTS(0.2, True)",
        P.QTC(0.4, True, 1.0): "QTC(0.4, True, 1.0)", P.FUN(f): f"FUN({f})",
        P.DIL(0.3): "DIL(clusters=0.3)", P.WIN(0.1, 0.9): "WIN(start=0.1, end=0.9)",
        P.DOT(3, 1): "DOT(n=3, first=1)", P.PDD(0.2, 0.4): "PDD(density=0.2, proportion=0.4)",
    }
    for prep, text in want.items():
        assert str(prep) == text and str(prep.copy()) == text
        assert type(prep.copy()) is type(prep) and prep.copy() is not prep
    assert P.MAV(3) == P.MAV(3) and P.MAV(3) != P.MAV(4)
    assert P.CTS(2) == P.CTS(2) and P.LAG() == P.LAG() and P.QTC(0.5) != P.QTC(0.6)
    assert P.FUN(f) != P.FUN(f) and P.FFN() != P.FFN()      # (no __eq__ in the reference)
    with pytest.raises(TypeError):
        P.WIN(0.1, 0.9) == 3
    assert not P.LAG().requires_fitting and not P.WIN(0, 1).requires_fitting
    assert P.MAV().requires_fitting and P.CTS(1).requires_fitting      # (Seed default)
    # what a row-sharded fit and the chunked host path need to know
    assert not P.QTC(0.5)._row_independent_fit() and P.RIN()._row_independent_fit()
    assert not P.FUN(f)._row_independent_transform() and not P.NEW(P.FUN(f))._row_independent_transform()
    assert P.WIN(0, 1)._needs_raw_cache() and P.SPE(0.5, step_transform="L2")._needs_raw_cache()
    assert not P.SPE(0.5)._needs_raw_cache()


def test_preparateur_fits_draw_like_the_oracle():
    """``fit`` of the random preparateurs only looks at the shape of its input
    and draws from the global numpy RNG: same weights and same generator state
    as the oracle restatement (which gen_golden.py pinned to the reference)."""
    import torch
    from cases import PREP2_CASES, make_prep2_inputs
    from oracle import preps as more
    for name in ("ffn", "ffn_out3_relu", "rin", "rin_w4_adaptive", "rin_outdim2_sum1",
                 "rin_callable", "rdw_uniform", "jld", "jld_distribute_bias", "jld_float",
                 "dil", "dil_clusters", "dot_float", "pdd", "pdd_dense", "mav_float"):
        desc = PREP2_CASES[name]
        X, _ = make_prep2_inputs(name)
        np.random.seed(7)
        st = more.fit_prep(desc, X)
        after = np.random.random()
        prep = specs._prep(fruits, desc)
        np.random.seed(7)
        prep._fit_device(torch.from_numpy(X))       # (shape only: no GPU needed)
        assert np.random.random() == after, name
        got = {"w1": "_weights1", "b": "_biases", "w2": "_weights2", "kernel": "_kernel",
               "ndim": "_ndim_per_kernel", "dims": "_dims_per_kernel", "weights": "_weights",
               "bias": "_bias_weights", "indices": "_indices", "lengths": "_lengths",
               "n": "_n", "first": "_first", "width": "_width", "w": "_w"}
        for key, val in st.items():
            np.testing.assert_array_equal(np.asarray(getattr(prep, got[key])), np.asarray(val),
                                          err_msg=f"{name}.{key}")


def test_letters_and_generic_words():
    """Host side of words over Python letters (reference:
    fruits/iss/words/letters.py, word.py:9-125, creation.py:53-83)."""
    W = fruits.words
    assert W.letters.get_available()[:2] == ["DIM", "ABS"]
    el = W.ExtendedLetter("ABS(1)DIM(3)")
    assert str(el) == "[ABS(1)DIM(3)]" and len(el) == 2 and str(el.copy()) == str(el)
    X = np.array([[1.0, -2.0], [3.0, 4.0], [-5.0, 6.0]])
    np.testing.assert_array_equal(el[0](X), [1.0, 2.0])
    np.testing.assert_array_equal([f(X) for f in el][1], [-5.0, 6.0])
    el.append("DIM", 1)
    assert str(el) == "[ABS(1)DIM(3)DIM(2)]"
    with pytest.raises(RuntimeError, match="does not exist"):
        W.ExtendedLetter("NOPE(1)")
    name = "T_HALF"
    if name not in W.letters.get_available():
        @W.letter(name=name)
        def half(X, i):
            return X[i, :] / 2
    with pytest.raises(RuntimeError, match="already exists"):
        W.letter(name=name)(lambda X, i: X[i, :])
    with pytest.raises(ValueError):
        W.letter()
    with pytest.raises(RuntimeError):
        W.letter(abs, abs)
    word = W.Word("[ABS(1)][T_HALF(2)DIM(1)]")
    assert len(word) == 2 and str(word) == "[ABS(1)][T_HALF(2)DIM(1)]"
    word.multiply(W.ExtendedLetter("DIM(2)"))
    word.multiply(W.Word("[ABS(2)]"))
    assert str(word) == "[ABS(1)][T_HALF(2)DIM(1)][DIM(2)][ABS(2)]" and str(word.copy()) == str(word)
    with pytest.raises(TypeError):
        word.multiply(3)
    np.testing.assert_array_equal(word.alpha, np.ones(4, dtype=np.float32))
    swapped = W.replace_letters(W.SimpleWord("[112][2]"), iter(["ABS", "T_HALF"]))
    assert str(swapped) == "[ABS(1)T_HALF(1)DIM(2)][DIM(2)]"
    iss = fruits.ISS([word, swapped, W.SimpleWord("[13]")], mode=fruits.ISSMode.EXTENDED)
    assert iss.n_iterated_sums() == 4 + 2 + 1 and iss.max_dim() == 3
    assert iss.label(1) == "[ABS(1)][T_HALF(2)DIM(1)]" and iss.label(6) == "[13]"
    assert str(iss.copy().words[0]) == str(word)


def test_shifted_rows_move_emissions_back_in_time():
    """Generic words in the Bayesian semiring run through the unshifted kernels on
    letter rows moved left; ``_ShiftedRows`` moves emission ``e`` back ``shifts[e]``
    steps (zeros in front) in every access path: whole, ranges, chunks, batches."""
    import torch
    from fruits_b200.iss.iss import _ShiftedRows

    rows = torch.arange(1, 4 * 2 * 6 + 1, dtype=torch.float64).reshape(4, 2, 6)

    class Inner:
        def materialize(self, X, emit_range=None, lookup=None, trusted=False):
            lo, hi = (0, 4) if emit_range is None else emit_range
            return rows[lo:hi].clone()

        def iter_chunks(self, X, max_bytes=0, emit_range=None, rows_for_size=None):
            for lo in range(0, 4, 3):
                yield lo, rows[lo:lo + 3].clone()

        def batch_transform(self, X, batch_size=1):
            yield rows[:1].clone()
            yield rows[1:].clone()

    twin = _ShiftedRows(Inner(), [0, 2, 6, 9])
    twin._cache = "shared"
    assert twin._inner._cache == "shared"
    want = torch.zeros_like(rows)
    want[0] = rows[0]
    want[1, :, 2:] = rows[1, :, :4]                     # (6 and 9 steps: nothing is left)
    assert torch.equal(twin.materialize(None), want)
    assert torch.equal(twin.materialize(None, (1, 3)), want[1:3])
    assert torch.equal(torch.cat([c for _, c in twin.iter_chunks(None)]), want)
    assert [lo for lo, _ in twin.iter_chunks(None)] == [0, 3]
    assert torch.equal(torch.cat(list(twin.batch_transform(None, 2))), want)
