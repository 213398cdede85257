"""Seeded test cases shared by the golden-vector generator
(``oracle/gen_golden.py``), the oracle tests and the GPU parity tests."""
import numpy as np

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
    # Bayesian (max, times) semiring, SURVEY.md section 8(f) rank 2
    "bayes_w3d2_ext": ({"words": {"of_weight": [3, 2]}, "mode": "extended",
                        "semiring": "bayesian"}, (5, 2, 300), "unit"),
    "bayes_single": ({"words": ["[1][2][1]", "[2]", "[12][1]"], "mode": "single",
                      "semiring": "bayesian"}, (4, 2, 70), "unit"),
    "bayes_neg": ({"words": ["[-1][2]", "[1][-2-2][1]", "[11][2]"], "mode": "extended",
                   "semiring": "bayesian"}, (3, 2, 257), "uniform1"),
    "bayes_indices": ({"words": ["[1][1]", "[1][2][1]"], "mode": "extended",
                       "semiring": "bayesian", "weighting": ["Indices", {"scale": 3}]},
                      (3, 2, 60), "unit"),
    "bayes_L1_total": ({"words": ["[1][1]", "[1][2][1]"], "mode": "extended",
                        "semiring": "bayesian",
                        "weighting": ["L1", {"total": True, "scale": 2}],
                        "alphas": [[0.5, 1.0], [1.0, 0.25, 2.0]]},
                       (3, 2, 600), "unit"),
}


COS_CASES = {
    # cosine weighted ISS (reference: fruits/iss/cos.py)
    "cos_e2": ({"words": ["[1]", "[1][2]", "[12][1]", "[1][2][11]"],
                "coswiss": {"freqs": [0.25, 0.05], "exponent": 2}}, (4, 2, 70), "std"),
    "cos_e1_total": ({"words": {"concat": [{"of_weight": [1, 2]}, {"of_weight": [2, 2]}]},
                      "coswiss": {"freqs": [0.05, 0.15, 0.45], "exponent": 1, "total": True}},
                     (3, 2, 64), "std"),
    "cos_e3_total": ({"words": ["[2]", "[1][-1]", "[1][2][1]"],
                      "coswiss": {"freqs": [0.3], "exponent": 3, "total": True}},
                     (3, 2, 50), "uniform1"),
    "cos_e4": ({"words": ["[1][1]", "[11][2]"],
                "coswiss": {"freqs": [0.1, 0.7], "exponent": 4}}, (2, 2, 41), "normal"),
    # four letters with the total weighting: 81 expansion terms (the C3 slices)
    "cos_e2_total_4letters": ({"words": ["[1][2][1][2]", "[2][1][1][-2]"],
                               "coswiss": {"freqs": [0.05, 0.45], "exponent": 2, "total": True}},
                              (3, 2, 40), "uniform1"),
}


def make_iss_input(shape, kind, seed=7):
    rng = np.random.default_rng(seed)
    if kind == "normal":
        return rng.standard_normal(shape)
    if kind == "walk":
        return rng.standard_normal(shape).cumsum(axis=2)
    if kind == "uniform1":
        return rng.random(shape) + 0.5
    if kind == "unit":
        return rng.random(shape)
    if kind == "std":
        x = rng.standard_normal(shape).cumsum(axis=2)
        return (x - x.mean(axis=2, keepdims=True)) / x.std(axis=2, keepdims=True)
    raise ValueError(kind)


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
    # SURVEY.md section 8(f) rank 2
    "cur_default": ["CUR", {}],
    "cur_q_cuts": ["CUR", {"cut": [12, 0.6, -1], "q": [-1.0, 0.4, 1.0]}],
    "avg_default": ["AVG", {}],                      # runs CUR's backend in the reference
    "std_q": ["STD", {"q": [0.3, 1.0]}],             # so does STD
    "cpv_default": ["CPV", {}],
    "cpv_multi": ["CPV", {"quantile": [0.2, 0.0, 0.9],
                          "constant": [False, True, False]}],
    "cpv_segments": ["CPV", {"quantile": [0.2, 0.5, 0.9], "segments": True}],
    # sieve wrappers (fruits/sieving/wrapper.py)
    "inc_of_max": ["INC", {"sieve": ["MAX", {"q": [-1.0, 0.5, 1.0]}]}],
    "inc_shift3_of_npi": ["INC", {"sieve": ["NPI", {"q": [0.3, 1.0], "inc": 0}], "depth": 2,
                                  "shift": 3}],
    "inc_of_npi_cocut": ["INC", {"sieve": ["NPI", {"cut": [0.5, -1]}]}],
    "int_of_end": ["INT", {"sieve": ["END", {"cut": [7, -1]}]}],
    "int_of_ppv": ["INT", {"sieve": ["PPV", {"quantile": [0.4]}]}],
}

# sieves whose value is a floating-point sum (order unspecified under numba fastmath)
SUMMING_SIEVES = ("MPI", "XPI", "CUR", "AVG", "STD")
IMPLICIT_SIEVES = ("PPV", "CPV")



def make_sieve_input():
    """-> (raw [12,2,48] for the coquantile cache, Y [12,48] to sieve)"""
    rng = np.random.default_rng(11)
    raw = rng.standard_normal((12, 2, 48)).cumsum(axis=2)
    Y = rng.standard_normal((12, 48)).cumsum(axis=1)
    Y[3] = Y[3, 0] + 0.0 * Y[3]  # constant row
    Y[5, 7:11] = Y[5, 6]         # ties
    Y -= np.median(Y, axis=1, keepdims=True) - 0.25  # every row straddles 0.25
    return raw, Y


PREP_CASES = {
    "inc": ["INC", {}],
    "inc_shift3_depth2": ["INC", {"shift": 3, "depth": 2}],
    "inc_nopad": ["INC", {"zero_padding": False}],
    "std": ["STD", {}],
    "std_novar": ["STD", {"var": False}],
    "nrm": ["NRM", {}],
    "new_inc": ["NEW", ["INC", {}]],
    "new_none": ["NEW", None],
    "dim_inc": ["DIM", {"preparateur": ["INC", {}], "dim": 1}],
    "dim_std_reordered": ["DIM", {"preparateur": ["STD", {}], "dim": [2, 0]}],
}



# the preparateurs beside INC / STD / NRM (fruits/preparation/transform.py:212-1048,
# filter.py); "@name" = a callable of oracle/preps.py::CALLABLES.  Frozen from the
# reference by ``oracle/gen_golden.py preps2`` (fit under np.random.seed(7) on the first
# input, transform of both inputs, state of the RNG behind fit).
PREP2_CASES = {
    "nrm_scale_dim": ["NRM", {"scale_dim": True}],
    "mav": ["MAV", {}],
    "mav_float": ["MAV", {"width": 0.25}],
    "mav_wide": ["MAV", {"width": 64}],
    "lag": ["LAG", {}],
    "ffn": ["FFN", {}],
    "ffn_out3_relu": ["FFN", {"d_out": 3, "d_hidden": 5, "center": False, "relu_out": True}],
    "rin": ["RIN", {}],
    "rin_w4_adaptive": ["RIN", {"width": 4, "adaptive_width": True}],
    "rin_outdim2_sum1": ["RIN", {"width": 3, "out_dim": 2, "force_sum_one": True}],
    "rin_kernel": ["RIN", {"kernel": [[1.0, -0.5], [0.25, 0.5], [2.0, 0.0]]}],
    "rin_callable": ["RIN", {"width": "@half"}],
    "rdw": ["RDW", {}],
    "rdw_uniform": ["RDW", {"dist": "uniform"}],
    "jld": ["JLD", {"dim": 2}],
    "jld_distribute_bias": ["JLD", {"dim": 2, "distribute": True, "bias": True}],
    "jld_float": ["JLD", {"dim": 0.99, "bias": True}],
    "spe": ["SPE", {"freq": 0.5}],
    "spe_additive_maxlen": ["SPE", {"freq": 0.3, "operation": "additive", "max_length": 100}],
    "spe_L1": ["SPE", {"freq": 0.5, "step_transform": "L1"}],
    "spe_L2_cos_maxlen": ["SPE", {"freq": 0.7, "step_transform": "L2", "function": "@cos",
                                  "max_length": 30}],
    "rpe": ["RPE", {"freq": 0.5}],
    "rpe_maxlen": ["RPE", {"freq": 0.25, "max_length": 100}],
    "cts": ["CTS", {"s": 3}],
    "cts_float": ["CTS", {"s": 0.25}],
    "cts_pseudo": ["CTS", {"s": 5, "pseudo_shift": True}],
    "cts_long": ["CTS", {"s": 100}],
    "qtc": ["QTC", {"q": 0.7}],
    "qtc_lower_bound": ["QTC", {"q": 0.2, "lower": True, "bound": -1.5}],
    "fun": ["FUN", {"f": "@sq"}],
    "dil": ["DIL", {}],
    "dil_clusters": ["DIL", {"clusters": 0.2}],
    "win": ["WIN", {"start": 0.1, "end": 0.8}],
    "win_negative_start": ["WIN", {"start": -0.1, "end": 0.5}],
    "dot": ["DOT", {}],
    "dot_float": ["DOT", {"n": 0.1, "first": 0.3}],
    "dot_first": ["DOT", {"n": 3, "first": 1}],
    "pdd": ["PDD", {}],
    "pdd_dense": ["PDD", {"density": 0.5, "proportion": 0.3}],
}
# plain numpy in the reference (copies, masks, np.where): compared bit for bit; the others
# are numba fastmath loops, BLAS or libm calls: 1e-12 of the row maximum
PREP2_EXACT = ("NRM", "LAG", "CTS", "QTC", "FUN", "DIL", "WIN", "DOT", "PDD")


def make_prep2_inputs(name: str):
    """(fit / first transform input, second transform input) of a PREP2 case."""
    X = make_prep_input()
    X2 = np.random.default_rng(6).standard_normal((4, 3, 40)).cumsum(axis=2)
    if PREP2_CASES[name][0] == "RPE":          # two-dimensional series only
        X, X2 = np.ascontiguousarray(X[:, :2]), np.ascontiguousarray(X2[:, :2])
    if PREP2_CASES[name][0] == "RDW":          # x ** w: positive values (plus one negative row)
        X, X2 = np.abs(X) + 0.5, np.abs(X2) + 0.5
        X2[1, 2] *= -1.0
    return X, X2


def make_prep_input():
    X = np.random.default_rng(5).standard_normal((6, 3, 40)).cumsum(axis=2)
    X[2, 1] = 4.0
    return X


PIPE_CASES = {
    # name: (spec name, number of series)
    "C1_readme": ("C1_readme", 200),
    "C2_reduced": ("C2_reduced", 24),
    "C3_general": ("C3_general", 4),
    "C4_twi": ("C4_twi", 8),
    "C5_sweep": ("C5_sweep", 32),
}

# Arctic(argmax=True): maxima and the positions that produced them
# (fruits/iss/semiring.py:234-279), frozen by ``oracle/gen_golden.py argmax``
ARGMAX_CASES = {
    "argmax": ({"words": ["[1]", "[1][2]", "[2][1][12]"], "mode": "extended",
                "semiring": "arctic_argmax"}, (5, 2, 40), "walk"),
    "argmax_signs_long": ({"words": ["[-1][2][-2][1][11][-2-2]", "[2]"], "mode": "extended",
                           "semiring": "arctic_argmax"}, (3, 2, 300), "normal"),
    "argmax_indices": ({"words": ["[1][2]", "[2][1][1]"], "mode": "extended",
                        "semiring": "arctic_argmax", "weighting": ["Indices", {"scale": 2}],
                        "alphas": [[0.5, 0.2], [0.3, 0.1, 0.7]]}, (4, 2, 64), "walk"),
    "argmax_L1_total": ({"words": ["[1][1][2]"], "mode": "extended", "semiring": "arctic_argmax",
                         "weighting": ["L1", {"total": True, "scale": 3}]}, (4, 2, 50), "walk"),
}

# words over Python letters through ISS.transform (Semiring._iterated_sum,
# fruits/iss/semiring.py:54-75, :428-446), frozen by ``oracle/gen_golden.py letters``
LETTER_CASES = {
    "letters_reals": ({"words": ["[ABS(1)DIM(2)][DIM(1)]", "[ABS(1)DIM(2)][RELU(2)][DIM(1)DIM(1)]",
                                 "[LAGDIFF(1)]", "[12][1]"], "mode": "extended"},
                      (5, 2, 40), "std"),
    "letters_arctic": ({"words": ["[ABS(1)][DIM(2)RELU(1)][DIM(1)]", "[ABS(1)][LAGDIFF(2)]",
                                  "[1][-2]"], "mode": "extended", "semiring": "arctic"},
                       (4, 2, 33), "walk"),
    "letters_bayesian": ({"words": ["[ABS(1)][DIM(2)DIM(2)][ABS(1)DIM(2)]", "[ABS(1)][ABS(2)]",
                                    "[1][2]", "[RELU(1)]"], "mode": "extended",
                          "semiring": "bayesian"}, (4, 2, 30), "unit"),
    "letters_mixed_weighted": ({"words": ["[1][2]", "[ABS(1)][DIM(2)]", "[2][1][1]", "[RELU(2)]"],
                                "mode": "extended", "weighting": ["Indices", {"scale": 3}],
                                "alphas": [[0.6, 0.2], None, [0.3, 0.1, 0.5], None]},
                               (4, 2, 36), "std"),
    "letters_mixed_weighted_arctic": ({"words": ["[ABS(1)][LAGDIFF(2)]", "[1][-2]", "[2]"],
                                       "mode": "single", "semiring": "arctic",
                                       "weighting": ["L1", {"total": True, "scale": 2}]},
                                      (4, 2, 33), "walk"),
    "letters_bayesian_single": ({"words": ["[ABS(2)][ABS(1)][DIM(2)]", "[2][1]"], "mode": "single",
                                 "semiring": "bayesian"}, (3, 2, 260), "unit"),
}

COS_PIPE_CASES = {"C2_cos": ("C2_cos", 16)}

# the randomised CosWISS variants (fruits/iss/cos.py:243-260, :306-324): fit under
# np.random.seed(3), then transform; frozen by ``oracle/gen_golden.py cos2``
COS_RANDOM_CASES = {
    "cos_ffn": ({"words": ["[1]", "[1][2]", "[12][1]"],
                 "coswiss": {"freqs": [0.25, 0.05], "exponent": 2, "ffn_size": 4}},
                (4, 2, 60), "std"),
    "cos_ffn_total_e1": ({"words": ["[2][1]", "[1][2][11]"],
                          "coswiss": {"freqs": [0.15], "exponent": 1, "total": True,
                                      "ffn_size": 3}}, (3, 2, 50), "std"),
    "cos_dropout": ({"words": ["[1]", "[1][2]", "[12][1]", "[1][-2][11]"],
                     "coswiss": {"freqs": [0.25, 0.05], "exponent": 2, "dropout": 0.2}},
                    (4, 2, 60), "uniform1"),
    "cos_dropout_total_e3": ({"words": ["[2][1]", "[1][2][1]"],
                              "coswiss": {"freqs": [0.3], "exponent": 3, "total": True,
                                          "dropout": 0.5}}, (3, 2, 50), "std"),
    "cos_ffn_and_dropout": ({"words": ["[1][2]"],
                             "coswiss": {"freqs": [0.1, 0.4], "exponent": 2, "ffn_size": 2,
                                         "dropout": 0.3}}, (3, 2, 40), "std"),
}

# pipelines of the rank 2-3 components, frozen from the reference
EXTRA_PIPE_CASES = {"R_mixed": ("R_mixed", 40), "R_rng": ("R_rng", 30),
                    "R_preps": ("R_preps", 36), "R_letters": ("R_letters", 33),
                    "R_cosrand": ("R_cosrand", 30), "R_argmax": ("R_argmax", 28)}




def sieve_kind(desc) -> str:
    """Name of the innermost sieve of a (possibly wrapped) sieve description."""
    while desc[0] in ("INC", "INT"):
        desc = desc[1]["sieve"]
    return desc[0]


def unwrap(sieve):
    """Innermost sieve object (product, reference: ``_sieve``; oracle: ``inner``)."""
    while hasattr(sieve, "_sieve") or hasattr(sieve, "inner"):
        sieve = getattr(sieve, "_sieve", None) or sieve.inner
    return sieve


# Known answers of the reference's own sieve tests, as data: (sieve description,
# row block of KAT_X the sieve is applied to, expected features).  Sources:
# tests/sieving/test_explicit.py:11-165 (MAX :11-40, MIN :43-72, END :75-97,
# NPI :100-137, MPI :140-148, XPI :151-159, LPI :162-170) and
# tests/sieving/test_implicit.py:12-74 (PPV, CPV) of the reference checkout.
KAT_X = np.array([
    [[-4., .8, 0., 5., -3.], [2., 1., 0., 0., -7.]],
    [[5., 8., 2., 6., 0.], [-5., -1., -4., -.5, -8.]],
])

SIEVE_KATS = [
    (["MAX", {}], 0, [[5], [2]]),
    (["MAX", {"cut": 3}], 0, [[0.8], [2]]),
    (["MAX", {"cut": 0.5}], 0, [[5], [2]]),
    (["MAX", {"cut": [-1, 3, 1]}], 0, [[-4, 0.8, 5], [2, 1, 0]]),
    (["MAX", {"cut": [-1, 0.2, 0.7, 0.5]}], 0, [[-4, 5, 0, -3], [2, 0, 0, -7]]),
    (["MIN", {}], 1, [[0], [-8]]),
    (["MIN", {"cut": 3}], 0, [[-4], [0]]),
    (["MIN", {"cut": 0.5}], 1, [[2], [-5]]),
    (["MIN", {"cut": [-1, 3, 1]}], 1, [[5, 2, 0], [-5, -4, -8]]),
    (["MIN", {"cut": [-1, 0.2, 0.7, 0.5]}], 1, [[5, 2, 6, 0], [-5, -4, 0, -8]]),
    (["END", {}], 0, [[-3], [-7]]),
    (["END", {"cut": 0.2}], 0, [[-4], [0]]),
    (["END", {"cut": [1, 0.2, 0.8, 4, -1]}], 0, [[-4, -4, 5, 5, -3], [2, 0, 0, 0, -7]]),
    (["NPI", {}], 0, [[2], [0]]),
    (["NPI", {"cut": 3}], 0, [[1], [0]]),
    (["NPI", {"cut": 0.5}], 1, [[1], [2]]),
    (["NPI", {"cut": [-1, 3, 1]}], 1, [[0, 1, 1], [0, 1, 1]]),
    (["NPI", {"cut": [-1, 0.2, 0.7, 0.5]}], 1, [[1, 0, 1, 0], [1, 1, 0, 0]]),
    (["MPI", {}], 0, [[4.9], [0]]),
    (["MPI", {}], 1, [[3.5], [3.75]]),
    (["XPI", {}], 0, [[2], [0]]),
    (["XPI", {}], 1, [[2], [2]]),
    (["LPI", {}], 0, [[1], [0]]),
    (["LPI", {}], 1, [[1], [1]]),
    (["PPV", {"quantile": 0, "constant": True}], 0, [[3 / 5], [4 / 5]]),
    (["PPV", {"quantile": 0.5, "constant": False, "sample_size": 1}], 1, [[1], [0]]),
    (["CPV", {"quantile": 0, "constant": True}], 0, [[1 / 3], [0.0]]),
    (["PPV", {"quantile": [0.5, 0.1, 0.7], "constant": False, "sample_size": 1}], 1,
     [[1., 1., 3 / 5], [0., 4 / 5, 0.]]),
    (["PPV", {"quantile": [0.5, 0.1, 0.7], "constant": False, "sample_size": 1,
              "segments": True}], 1, [[0., 2 / 5], [4 / 5, 0.]]),
    (["PPV", {"quantile": [-5, 0, 2], "constant": True, "sample_size": 1}], 1,
     [[1., 1., 4 / 5], [4 / 5, 0., 0.]]),
    (["PPV", {"quantile": [0, -5, 2], "constant": True, "sample_size": 1, "segments": True}], 1,
     [[0., 1 / 5], [4 / 5, 0.]]),
]
