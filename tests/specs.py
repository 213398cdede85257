"""Pipeline descriptions shared by the parity tests, the oracle and the
golden-vector generator.

A ``spec`` is a plain dict (format documented in ``oracle/pipeline.py``).
``build_fruit(mod, spec)`` instantiates it against any module that exposes the
reference API -- ``fruits_b200`` (the product) or the real reference package
(only inside ``oracle/gen_golden.py``, in the build container).
"""
import numpy as np


def _seven_sieves():
    # experiments/fruit_reduced.py:33-39
    return ([["NPI", {"q": [0.5, 1.0], "inc": i}] for i in range(3)]
            + [["MPI", {"q": [0.5, 1.0], "inc": i}] for i in range(3)]
            + [["END", {}]])


def _alt(words):
    return {"alternate_sign": words}


SPECS = {
    # README.md:67-99 of the reference
    "C1_readme": {"slices": [
        {"preps": [["INC", {}]],
         "iss": [{"words": {"of_weight": [2, 3]}, "mode": "extended"}],
         "sieves": [["NPI", {"q": [0.5, 1.0]}], ["END", {}]]},
        {"preps": [],
         "iss": [{"words": {"of_weight": [2, 3]}, "mode": "extended"}],
         "sieves": [["NPI", {}], ["END", {}]]},
    ]},
    # experiments/fruit_reduced.py:27-49 (slices 0-1; CosWISS slices are "next")
    "C2_reduced": {"slices": [
        {"preps": [["NEW", ["INC", {}]], ["STD", {}]],
         "iss": [{"words": {"of_weight": [4, 2]}, "mode": "extended",
                  "semiring": "reals", "weighting": ["Indices", {}]}],
         "sieves": _seven_sieves(), "fit_sample_size": 1.0},
        {"preps": [["NEW", ["INC", {}]]],
         "iss": [{"words": _alt([24 * "[1]", 24 * "[2]", 12 * "[1][2]",
                                 12 * "[2][1]"]),
                  "mode": "extended", "semiring": "arctic"}],
         "sieves": _seven_sieves(), "fit_sample_size": 1.0},
    ]},
    # experiments/fruit_general.py:28-51 (slices 0-1)
    "C3_general": {"slices": [
        {"preps": [["NEW", ["INC", {}]], ["STD", {}]],
         "iss": [{"words": {"of_weight": [6, 2]}, "mode": "extended",
                  "semiring": "reals", "weighting": ["Indices", {}]}],
         "sieves": _seven_sieves(), "fit_sample_size": 1.0},
        {"preps": [["NEW", ["INC", {}]]],
         "iss": [{"words": _alt([48 * "[1]", 48 * "[2]", 24 * "[1][2]",
                                 24 * "[2][1]"]),
                  "mode": "extended", "semiring": "arctic"}],
         "sieves": _seven_sieves(), "fit_sample_size": 1.0},
    ]},
    # experiments/fruit_twi.py:3-31
    "C4_twi": {"slices": [
        {"preps": [["INC", {}]],
         "iss": [{"words": {"of_weight": [9, 1]}, "mode": "extended",
                  "semiring": "reals", "weighting": ["L1", {}]}],
         "sieves": [["NPI", {}], ["MPI", {}], ["END", {}]],
         "fit_sample_size": 1.0},
        {"preps": [],
         "iss": [{"words": _alt([48 * "[1]"]), "mode": "extended",
                  "semiring": "arctic"}],
         "sieves": [["NPI", {}], ["END", {}]], "fit_sample_size": 1.0},
    ]},
    # experiments/fruit_reduced.py:52-68 (slices 2-3: cosine weighted ISS, exponents 1 and 2)
    "C2_cos": {"slices": [
        {"preps": [["NEW", ["INC", {}]], ["STD", {}]],
         "iss": [{"words": {"concat": [{"of_weight": [1, 2]}, {"of_weight": [2, 2]},
                                       {"of_weight": [3, 2]}]},
                  "coswiss": {"freqs": [i / 20 for i in range(1, 11, 2)], "exponent": e,
                              "total": True}}],
         "sieves": _seven_sieves(), "fit_sample_size": 1.0}
        for e in (1, 2)]},
    # throughput sweep (BASELINE.json configs[4])
    "C5_sweep": {"slices": [
        {"preps": [],
         "iss": [{"words": {"of_weight": [4, 3]}, "mode": "extended"}],
         "sieves": [["NPI", {"q": [0.5, 1.0]}], ["PPV", {}], ["MAX", {}],
                    ["MIN", {}], ["END", {}]]},
    ]},
}


# experiments/fruit_general.py:53-69 (slices 2-3: 115 words up to four letters)
SPECS["C3_cos"] = {"slices": [
    {"preps": [["NEW", ["INC", {}]], ["STD", {}]],
     "iss": [{"words": {"concat": [{"of_weight": [w, 2]} for w in (1, 2, 3, 4)]},
              "coswiss": {"freqs": [i / 20 for i in range(1, 11, 2)], "exponent": e,
                          "total": True}}],
     "sieves": _seven_sieves(), "fit_sample_size": 1.0}
    for e in (1, 2)]}
SPECS["C3_full"] = {"slices": SPECS["C3_general"]["slices"] + SPECS["C3_cos"]["slices"]}

# SURVEY.md section 8(f) ranks 2-3 in one fruit: a Bayesian slice with rank-2
# sieves and wrappers, a slice with two chained ISS (fruits/fruit.py:440-454)
SPECS["R_mixed"] = {"slices": [
    {"preps": [],
     "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended", "semiring": "bayesian"}],
     "sieves": [["NPI", {"q": [0.4, 1.0]}], ["CPV", {"quantile": [0.3, 0.8]}],
                ["CUR", {"cut": [9, -1], "q": [-1.0, 0.5, 1.0]}],
                ["INC", {"sieve": ["MAX", {"q": [-1.0, 0.6]}]}],
                ["INT", {"sieve": ["END", {}]}], ["LPI", {}], ["END", {}]],
     "fit_sample_size": 1.0},
    {"preps": [["INC", {}]],
     "iss": [{"words": ["[1]", "[1][2]"], "mode": "extended"},
             {"words": ["[1]", "[1][1]"], "mode": "single", "semiring": "arctic"}],
     "sieves": [["NPI", {"q": [0.5, 1.0]}], ["XPI", {}], ["MIN", {}], ["END", {"cut": [5, -1]}]],
     "fit_sample_size": 0.5}]}

# order in which fit consumes the global numpy RNG (fruits/fruit.py:430-438, then per
# iterated sum and non-constant PPV quantile fruits/sieving/implicit.py:103-108)
SPECS["R_rng"] = {"slices": [
    {"preps": [["INC", {}]],
     "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended"}],
     "sieves": [["PPV", {"quantile": [0.3, 0.0, 0.8], "constant": [False, True, False],
                         "sample_size": 0.5}], ["NPI", {"q": [0.4, 1.0]}], ["END", {}]],
     "fit_sample_size": 0.7},
    {"iss": [{"words": ["[1][2]", "[2]"], "mode": "single", "semiring": "arctic"}],
     "sieves": [["PPV", {"sample_size": 0.25}], ["MAX", {"q": [-1.0, 0.5]}]],
     "fit_sample_size": 1}]}

# the preparateurs beside INC / STD in front of every kind of slice: a prepared copy, then
# the fused kernels (slices 0-2, FruitSlice._transform_prepared_fused) or the composed
# route (slice 3: two cuts); fit draws from the RNG in the preparateurs, too
SPECS["R_preps"] = {"slices": [
    {"preps": [["LAG", {}], ["INC", {}]],
     "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended"}],
     "sieves": [["NPI", {"q": [0.5, 1.0]}], ["END", {}]], "fit_sample_size": 1.0},
    {"preps": [["DOT", {"n": 3}], ["NEW", ["INC", {}]], ["STD", {}]],
     "iss": [{"words": {"of_weight": [2, 2]}, "mode": "extended",
              "weighting": ["Indices", {}]}],
     "sieves": _seven_sieves(), "fit_sample_size": 1.0},
    {"preps": [["INC", {}], ["RIN", {"width": 3}], ["MAV", {"width": 4}]],
     "iss": [{"words": ["[1][2]", "[2][-1][1]"], "mode": "extended", "semiring": "arctic"}],
     "sieves": [["MAX", {}], ["MIN", {}], ["PPV", {}]], "fit_sample_size": 0.5},
    {"preps": [["WIN", {"start": 0.1, "end": 0.9}], ["JLD", {"dim": 2, "bias": True}],
               ["QTC", {"q": 0.9}]],
     "iss": [{"words": ["[1]", "[1][2]"], "mode": "extended"}],
     "sieves": [["NPI", {"q": [0.5, 1.0]}], ["END", {"cut": [7, -1]}]],
     "fit_sample_size": 1}]}

# words over Python letters (fruits/iss/words/letters.py; Semiring._iterated_sum): a slice
# of generic words behind a preparateur, a mixed ISS, an arctic one, chained behind an ISS
SPECS["R_letters"] = {"slices": [
    {"preps": [["INC", {}]],
     "iss": [{"words": ["[ABS(1)DIM(2)][DIM(1)]", "[ABS(1)DIM(2)][RELU(2)][DIM(1)DIM(1)]",
                        "[LAGDIFF(1)]", "[RELU(2)][DIM(1)DIM(1)]"], "mode": "extended"}],
     "sieves": [["NPI", {"q": [0.5, 1.0]}], ["MAX", {}], ["END", {}]], "fit_sample_size": 1.0},
    {"preps": [],
     "iss": [{"words": ["[1][2]", "[DIM(1)][DIM(2)]", "[12]", "[ABS(1)ABS(2)]"],
              "mode": "single"}],
     "sieves": [["PPV", {}], ["MPI", {"q": [0.3, 1.0]}]], "fit_sample_size": 1.0},
    {"preps": [],
     "iss": [{"words": ["[ABS(1)][DIM(2)RELU(1)][DIM(1)]", "[ABS(1)][LAGDIFF(2)]"],
              "mode": "extended", "semiring": "arctic"}],
     "sieves": [["MIN", {}], ["NPI", {"inc": 0, "q": [0.5, 1.0]}], ["END", {"cut": [5, -1]}]],
     "fit_sample_size": 0.5},
    {"preps": [],
     "iss": [{"words": ["[1]", "[2]"], "mode": "single"},
             {"words": ["[ABS(1)][RELU(1)]"], "mode": "extended"}],
     "sieves": [["NPI", {}], ["END", {}]], "fit_sample_size": 1}]}

# the randomised CosWISS variants inside a fruit: ISS.fit draws between the preparateurs
# and the sieves (fruits/fruit.py:478-481)
SPECS["R_cosrand"] = {"slices": [
    {"preps": [["NEW", ["INC", {}]], ["STD", {}]],
     "iss": [{"words": ["[1]", "[1][2]", "[3][1]"],
              "coswiss": {"freqs": [0.1, 0.35], "exponent": 2, "total": True, "ffn_size": 3}}],
     "sieves": [["NPI", {"q": [0.5, 1.0]}], ["MPI", {"q": [0.5, 1.0]}], ["END", {}]],
     "fit_sample_size": 1.0},
    {"preps": [["INC", {}]],
     "iss": [{"words": ["[1][2]", "[2][1][1]"],
              "coswiss": {"freqs": [0.2], "exponent": 1, "dropout": 0.25}}],
     "sieves": [["PPV", {}], ["MAX", {}], ["END", {}]], "fit_sample_size": 0.5}]}

# Arctic(argmax=True) inside a fruit: maxima and position rows alike are sieved
SPECS["R_argmax"] = {"slices": [
    {"preps": [["INC", {}]],
     "iss": [{"words": ["[1][2]", "[2][1][1]"], "mode": "extended", "semiring": "arctic_argmax"}],
     "sieves": [["NPI", {"q": [0.5, 1.0]}], ["MAX", {}], ["END", {}]], "fit_sample_size": 1.0}]}

# two small candidates for corbeille.decide_which_fruit
SPECS["R_decide_a"] = {"slices": [
    {"preps": [["INC", {}]], "iss": [{"words": {"of_weight": [2, 1]}, "mode": "extended"}],
     "sieves": [["NPI", {"q": [0.5, 1.0]}], ["END", {}]], "fit_sample_size": 1.0}]}
SPECS["R_decide_b"] = {"slices": [
    {"preps": [], "iss": [{"words": ["[1]"], "mode": "single", "semiring": "arctic"}],
     "sieves": [["MAX", {}]], "fit_sample_size": 1.0}]}

# the complete experiments/fruit_reduced.py pipeline: all four slices (4,431 features)
SPECS["C2_full"] = {"slices": SPECS["C2_reduced"]["slices"] + SPECS["C2_cos"]["slices"]}


def make_input(name: str, n: int = None) -> np.ndarray:
    """Seeded synthetic input of SURVEY.md section 8(d) for a config; ``n``
    overrides the number of series (same generator, first ``n`` rows)."""
    shapes = {"C1_readme": (200, 3, 100), "C2_reduced": (1000, 1, 512), "C2_cos": (1000, 1, 512), "C2_full": (1000, 1, 512),
              "C3_cos": (10000, 6, 1024), "C3_full": (10000, 6, 1024), "R_mixed": (40, 2, 60), "R_rng": (30, 2, 50), "R_preps": (36, 2, 48), "R_letters": (33, 2, 45), "R_cosrand": (30, 2, 44), "R_argmax": (28, 2, 52),
              "C3_general": (10000, 6, 1024), "C4_twi": (100000, 3, 2048),
              "C5_sweep": (4096, 3, 1024)}
    N, D, T = shapes[name]
    n = N if n is None else n
    if name == "C1_readme":
        return np.random.default_rng(0).random((N, D, T))[:n]
    if name == "C5_sweep":
        return np.random.default_rng(1234).standard_normal((n, D, T))
    if name == "R_mixed":
        return np.random.default_rng(42).random((n, D, T)) + 0.1
    if name == "R_argmax":
        return np.random.default_rng(47).standard_normal((n, D, T)).cumsum(axis=2) / 3
    if name == "R_cosrand":
        return np.random.default_rng(46).standard_normal((n, D, T)).cumsum(axis=2) / 3
    if name == "R_letters":
        return np.random.default_rng(45).standard_normal((n, D, T)).cumsum(axis=2) / 4
    if name == "R_preps":
        return np.random.default_rng(44).standard_normal((n, D, T)).cumsum(axis=2)
    if name == "R_rng":
        return np.random.default_rng(43).standard_normal((n, D, T)).cumsum(axis=2)
    return np.random.default_rng(0).standard_normal((n, D, T)).cumsum(axis=2)


# ---------------------------------------------------------------------------

def _words(mod, desc):
    if isinstance(desc, dict):
        if "of_weight" in desc:
            return list(mod.words.of_weight(*desc["of_weight"]))
        if "alternate_sign" in desc:
            return list(mod.words.alternate_sign(_words(mod, desc["alternate_sign"])))
        if "concat" in desc:
            return [w for d in desc["concat"] for w in _words(mod, d)]
        raise ValueError(desc)
    _register_letters(mod)
    return [mod.words.Word(w) if any(c.isalpha() for c in w) else mod.words.SimpleWord(w)
            for w in desc]


def _register_letters(mod):
    """The letters the generic words of the specs use beside DIM / ABS, registered
    once per package through its own decorator (fruits/iss/words/letters.py:137-206);
    the same functions as ``oracle.pipeline.LETTERS``."""
    if "RELU" in mod.words.letters.get_available():
        return

    @mod.words.letter(name="RELU")
    def relu(X, i):
        return X[i, :] * (X[i, :] > 0)

    @mod.words.letter
    def LAGDIFF(X, i):
        return X[i, :] - np.roll(X[i, :], 2)


def _prep(mod, desc):
    name, args = desc
    if name == "NEW":
        return mod.preparation.NEW(None if args is None else _prep(mod, args))
    if name == "DIM":
        return mod.preparation.DIM(_prep(mod, args["preparateur"]), args["dim"])
    if any(isinstance(v, str) and v.startswith("@") for v in args.values()) or "kernel" in args:
        from oracle.preps import resolve          # named callables, kernel arrays (tests only)
        args = resolve(args)
    return getattr(mod.preparation, name)(**args)


def _weighting(mod, desc):
    if desc is None:
        return None
    name, args = desc
    return getattr(mod.iss.weighting, name)(**args)


def _semiring(mod, name):
    if name == "arctic_argmax":
        return mod.iss.semiring.Arctic(argmax=True)
    return {"reals": mod.iss.semiring.Reals,
            "arctic": mod.iss.semiring.Arctic,
            "bayesian": mod.iss.semiring.Bayesian}[name]()


def _sieve(mod, desc):
    name, args = desc
    args = dict(args)
    if name in ("INC", "INT"):           # sieve wrappers (fruits/sieving/wrapper.py)
        args["sieve"] = _sieve(mod, args["sieve"])
        return getattr(mod.sieving, name)(**args)
    for key in ("q", "cut"):
        if isinstance(args.get(key), list):
            args[key] = tuple(args[key])
    return getattr(mod.sieving, name)(**args)


def build_iss(mod, desc):
    words = _words(mod, desc["words"])
    if desc.get("coswiss") is not None:
        c = desc["coswiss"]
        return mod.CosWISS(words=words, freqs=list(c["freqs"]), exponent=c.get("exponent", 2),
                           total_weighting=c.get("total", False), ffn_size=c.get("ffn_size"),
                           dropout=c.get("dropout"))
    if desc.get("alphas") is not None:
        for w, a in zip(words, desc["alphas"]):
            if a is not None:
                w.alpha = a
    mode = (mod.ISSMode.EXTENDED if desc.get("mode", "single") == "extended"
            else mod.ISSMode.SINGLE)
    return mod.ISS(words, mode=mode,
                   semiring=_semiring(mod, desc.get("semiring", "reals")),
                   weighting=_weighting(mod, desc.get("weighting")))


def build_fruit(mod, spec, name="fruit"):
    fruit = mod.Fruit(name)
    for slc in spec["slices"]:
        fruit.cut()
        for p in slc.get("preps", []):
            fruit.add(_prep(mod, p))
        for i in slc["iss"]:
            fruit.add(build_iss(mod, i))
        for s in slc["sieves"]:
            fruit.add(_sieve(mod, s))
        fruit.get_slice().fit_sample_size = slc.get("fit_sample_size", 1)
    return fruit
