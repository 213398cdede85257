"""numpy restatement of the reference's fit/transform orchestration.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

A pipeline is described by a plain ``spec`` dict (see ``tests/specs.py``)::

    {"slices": [
        {"preps": [["INC", {}], ["NEW", ["INC", {}]], ["STD", {}]],
         "iss": [{"words": ["[1][2]", ...], "mode": "extended",
                  "semiring": "reals" | "arctic",
                  "weighting": None | ["Indices", {...}] | ["L1", {...}],
                  "alphas": None | [[...], ...]}],
         "sieves": [["NPI", {"q": [0.5, 1.0], "inc": 1}], ["END", {}]],
         "fit_sample_size": 1}]}

``OracleFruit(spec).fit(X)`` / ``.transform(X)`` reproduce
``fruits.Fruit.fit`` / ``.transform`` (fruits/fruit.py:121-173), including
the order in which the global ``np.random`` state is consumed.
"""
import ctypes
import itertools
import re

import numpy as np

from .build import load_oracle


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------
# words (fruits/iss/words/word.py:189-245, creation.py:9-50, :86-103)

_WORD_RE = re.compile(r"(\[(-?\d|\(-?\d+\))+\])+")


def parse_word(string: str) -> np.ndarray:
    """Exponent matrix ``[p, max_dim]`` of a SimpleWord string
    (fruits/iss/words/word.py:189-245)."""
    if not _WORD_RE.fullmatch(string):
        raise ValueError(f"bad word {string!r}")
    letters = []
    for el in string.split("]")[:-1]:
        el = el[1:]
        toks = re.findall(r"\((-?\d*)\)|(-\d?)|(\d)", el)
        ints = []
        for par, neg, pos in toks:
            if pos:
                ints.append(int(pos))
            elif neg:
                ints.append(-1 if neg == "-" else int(neg))
            else:
                ints.append(int(par) if par != "" else 1)
        letters.append(ints)
    max_dim = max(abs(x) for el in letters for x in el)
    mat = np.zeros((len(letters), max_dim), dtype=np.int32)
    for k, el in enumerate(letters):
        for x in el:
            mat[k, abs(x) - 1] += 1 if x > 0 else -1
    return mat


def _partitions(n, start=1):
    # fruits/iss/words/creation.py:9-13
    yield (n,)
    for i in range(start, n // 2 + 1):
        for p in _partitions(n - i, i):
            yield (i,) + p


def of_weight(w: int, dim: int = 1) -> list:
    """Word strings of weight ``w`` in the reference's enumeration order
    (fruits/iss/words/creation.py:26-50).  The order of the permutations of
    one partition is CPython's iteration order of a ``set`` of int tuples,
    exactly as in the reference (creation.py:45)."""
    els = []
    for i in range(1, w + 1):
        els.append([
            "[" + "".join(f"({x})" if x > 9 else str(x) for x in comb) + "]"
            for comb in itertools.combinations_with_replacement(
                range(1, dim + 1), i)
        ])
    words = []
    for part in _partitions(w):
        for perm in set(itertools.permutations(part)):
            for raw in itertools.product(*[els[k - 1] for k in perm]):
                words.append("".join(raw))
    return words


def alternate_sign(words: list) -> list:
    # fruits/iss/words/creation.py:86-103
    out = []
    for w in words:
        mat = parse_word(w)
        w1, w2 = "", ""
        for i, el in enumerate(mat):
            neg = "".join(int(c) * f"-{d + 1}" for d, c in enumerate(el))
            pos = neg.replace("-", "")
            if i % 2 == 0:
                w1 += f"[{neg}]"
                w2 += f"[{pos}]"
            else:
                w1 += f"[{pos}]"
                w2 += f"[{neg}]"
        out += [w1, w2]
    return out


def expand_words(desc) -> list:
    """``desc`` is a list of strings or a small generator description."""
    if isinstance(desc, dict):
        if "of_weight" in desc:
            return of_weight(*desc["of_weight"])
        if "alternate_sign" in desc:
            return alternate_sign(expand_words(desc["alternate_sign"]))
        if "concat" in desc:
            return [w for d in desc["concat"] for w in expand_words(d)]
        raise ValueError(desc)
    return list(desc)


def cache_plan(words: list) -> list:
    """Number of new prefixes per word (fruits/iss/cache.py:17-37)."""
    plan = []
    for i, wstr in enumerate(words):
        els = wstr.split("[")[1:]
        depth = len(els)
        for j in range(len(els)):
            prefix = "[" + "[".join(els[:j + 1])
            if any(words[k].startswith(prefix) for k in range(i)):
                depth -= 1
            else:
                break
        plan.append(depth)
    return plan


# ---------------------------------------------------------------------------
# raw-input cache (fruits/cache.py:51-135)

class RawCache:
    def __init__(self, X):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self._store = {}

    def lsum(self, key):
        if key not in self._store:
            n, d, t = self.X.shape
            out = np.zeros((n, t))
            load_oracle().fo_lsum(_p(self.X), _p(out), n, d, t,
                                  1 if key == "L2" else 0)
            self._store[key] = out
        return self._store[key]

    def coquantile(self, c: float, norm: str):
        key = f"{c}:{norm}"
        if key not in self._store:
            s = self.lsum(norm)
            out = np.zeros(s.shape[0], dtype=np.int64)
            load_oracle().fo_coquantile(_p(s), _p(out), s.shape[0],
                                        s.shape[1], float(c))
            self._store[key] = out
        return self._store[key]


# ---------------------------------------------------------------------------
# preparateurs

def increments(X, k=1):
    X = np.ascontiguousarray(X, dtype=np.float64)
    out = np.zeros_like(X)
    n, d, t = X.shape
    load_oracle().fo_increments(_p(X), _p(out), n, d, t, int(k))
    return out


def nrm(X, scale_dim=False):
    # fruits/preparation/transform.py:184-198
    min_ = np.min(X, axis=2)
    max_ = np.max(X, axis=2)
    if scale_dim:
        min_ = np.min(min_, axis=1)[:, np.newaxis]
        max_ = np.max(max_, axis=1)[:, np.newaxis]
    mask = (min_ != max_)
    if scale_dim:
        mask = mask[:, 0]
    min_ = min_[mask][:, np.newaxis]
    max_ = max_[mask][:, np.newaxis]
    out = np.zeros_like(X)
    out[mask] = (X[mask] - min_) / (max_ - min_)
    out[~mask] = 0
    return out


_STATELESS_PREPS = ("INC", "STD", "NRM")


def fit_prep_state(prep, X):
    """What ``fit`` leaves on a preparateur (None for INC / STD / NRM; the state
    of the wrapped preparateur for NEW / DIM; ``oracle/preps.py`` otherwise)."""
    name, args = prep
    if name in _STATELESS_PREPS:
        return None
    if name == "NEW":
        return None if args is None else fit_prep_state(args, X)
    if name == "DIM":
        dims = args["dim"] if isinstance(args["dim"], (list, tuple)) else [args["dim"]]
        return fit_prep_state(args["preparateur"], np.ascontiguousarray(X[:, list(dims), :]))
    from . import preps as more
    return more.fit_prep(prep, X)


def apply_prep(prep, X, state=None, cache=None):
    """Transform with one preparateur description ``[name, args]``; ``state``
    from ``fit_prep_state`` (fitted on X itself if omitted), ``cache`` the
    ``RawCache`` of the raw batch."""
    name, args = prep
    if name == "INC":
        # fruits/preparation/transform.py:62-76
        shift = args.get("shift", 1)
        depth = args.get("depth", 1)
        zero_padding = args.get("zero_padding", True)
        if isinstance(shift, float):
            shift = int(np.ceil(shift * X.shape[2]))
        out = X
        for _ in range(depth):
            out = increments(out, shift)
            if not zero_padding:
                out[:, :, :shift] = X[:, :, :shift]
        return out
    if name == "STD":
        # fruits/preparation/transform.py:132-144 (separately=True only)
        eps = args.get("std_eps", 1e-5)
        mean_ = np.mean(X, axis=2)[:, :, np.newaxis]
        std_ = np.ones((X.shape[0], X.shape[1], 1))
        if args.get("var", True):
            std_ = np.std(X, axis=2)[:, :, np.newaxis]
        return (X - mean_) / (std_ + eps)
    if name == "NRM":
        return nrm(X, args.get("scale_dim", False))
    if name == "NEW":
        # fruits/preparation/wrapper.py:78-96
        if args is None:
            return np.concatenate((X, X), axis=1)
        return np.concatenate((X, apply_prep(args, X, state, cache)), axis=1)
    if name == "DIM":
        # fruits/preparation/wrapper.py:37-43: the chosen dimensions (in the given
        # order) transformed and appended behind the untouched ones
        dims = args["dim"] if isinstance(args["dim"], (list, tuple)) else [args["dim"]]
        inner = apply_prep(args["preparateur"], np.ascontiguousarray(X[:, list(dims), :]),
                           state, cache)
        return np.concatenate((np.delete(X, list(dims), axis=1), inner), axis=1)
    from . import preps as more          # MAV, LAG, FFN, RIN, ... DIL, WIN, DOT, PDD
    if state is None:
        state = more.fit_prep(prep, X)
    return more.transform_prep(prep, state, X, cache)


# ---------------------------------------------------------------------------
# weightings (fruits/iss/weighting.py)

def lookup(weighting, X, cache: RawCache):
    name, args = weighting
    n, _, l = X.shape
    scale = args.get("scale", 50)
    if name == "Indices":
        # weighting.py:100-110
        r = np.arange(1, l + 1)
        if args.get("relative", True):
            r = r / l
        r = nrm(r[np.newaxis, np.newaxis, :])[0, 0, :] * scale
        return np.ones((n, l)) * r
    if name in ("L1", "L2"):
        # weighting.py:148-160 / 198-210: raw-input cache, dim 0
        if args.get("on_prepared", False):
            r = RawCache(X).lsum(name)
        else:
            r = cache.lsum(name)
        if args.get("relative", False):
            r = r / (r[:, -1:] + 1e-5)
        return nrm(r[:, np.newaxis, :])[:, 0, :] * scale
    if name == "Plateaus":
        # weighting.py:245-256
        npl = args["n"]
        r = np.ones(l)
        step = int(l / npl)
        for i in range(npl):
            r[i * step:(i + 1) * step] = i / (npl - 1)
        if args.get("reverse", False):
            r = r[::-1]
        r = nrm(np.ascontiguousarray(r)[np.newaxis, np.newaxis, :])[0, 0, :] * scale
        return np.ones((n, l)) * r
    raise NotImplementedError(name)


# ---------------------------------------------------------------------------
# ISS (fruits/iss/iss.py:21-67, fruits/iss/semiring.py:14-41)

def iss_word(X, word: str, extended: int, semiring: str, weighting,
             alpha, cache: RawCache):
    """Iterated sums of one word: ``[extended, n, t]``."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, d, t = X.shape
    mat = np.ascontiguousarray(parse_word(word))
    p, md = mat.shape
    if md > d:
        raise IndexError("word uses more dimensions than the input has")
    if weighting is not None:
        lk = np.ascontiguousarray(lookup(weighting, X, cache))
        a = (np.ones(p, dtype=np.float32) if alpha is None
             else np.asarray(alpha, dtype=np.float32))
        total = bool(weighting[1].get("total", False))
    else:
        lk = np.zeros((n, t))
        a = np.zeros(p, dtype=np.float32)
        total = True
    res = np.zeros((n, extended, t))
    load_oracle().fo_iterated_sums(
        _p(X), _p(mat), _p(a), _p(lk), _p(res), n, d, t, p, md, extended,
        {"reals": 0, "arctic": 1, "bayesian": 2}[semiring], 1 if total else 0)
    return np.ascontiguousarray(np.swapaxes(res, 0, 1))


def arctic_argmax_word(X, word: str, weighting, alpha, cache):
    """``[p + p(p+1)/2, n, t]``: Arctic(argmax=True) for one word --
    _arctic_argmax_single (fruits/iss/semiring.py:234-279) series by series, as
    Arctic._iterated_sum_fast calls it (:385-392; the total / non-total flag and
    ``extended`` play no role there)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, d, t = X.shape
    mat = parse_word(word)
    p = mat.shape[0]
    if weighting is not None:
        lk = np.ascontiguousarray(lookup(weighting, X, cache))
        a = (np.ones(p, dtype=np.float32) if alpha is None else np.asarray(alpha, dtype=np.float32))
    else:
        lk, a = np.zeros((n, t)), np.zeros(p, dtype=np.float32)
    rows = p + p * (p + 1) // 2
    out = np.zeros((rows, n, t))
    for j in range(n):
        Z, w = X[j], lk[j]
        result = np.zeros((2 * p, t))
        tmp = np.zeros(t)
        for k in range(p):
            if not np.any(mat[k]):
                continue
            C = np.zeros(t)
            for dim, el in enumerate(mat[k]):
                C = C + el * Z[dim, :]
            tmp = tmp + C
            if k > 0:
                tmp = tmp - w * a[k - 1]
            result[2 * k, 0] = tmp[0]
            for i in range(1, t):
                if result[2 * k, i - 1] >= tmp[i]:
                    result[2 * k, i] = result[2 * k, i - 1]
                    result[2 * k + 1, i] = result[2 * k + 1, i - 1]
                else:
                    result[2 * k, i] = tmp[i]
                    result[2 * k + 1, i] = i
            if k < p - 1:
                tmp = np.maximum.accumulate(tmp + w * a[k])
        for k in range(p - 1, -1, -1):
            index = k + k * (k + 1) // 2
            out[index, j] = result[2 * k]
            out[index + k + 1, j] = result[2 * k + 1]
            for s_ in range(k, 0, -1):
                c = int(out[index + s_ + 1, j, -1]) + 1
                out[index + s_, j, :c] = result[2 * (s_ - 1) + 1, :c]
                out[index + s_, j, c:] = result[2 * (s_ - 1) + 1, c - 1]
    return out


# letters of generic words (fruits/iss/words/letters.py:95-110 DIM / ABS; the others are
# what the tests register through ``fruits.words.letter`` -- tests/specs.py)
LETTERS = {
    "DIM": lambda X, i: X[i, :],
    "ABS": lambda X, i: np.abs(X[i, :]),
    "RELU": lambda X, i: X[i, :] * (X[i, :] > 0),
    "LAGDIFF": lambda X, i: X[i, :] - np.roll(X[i, :], 2),
}


def is_generic_word(word: str) -> bool:
    return any(c.isalpha() for c in word)


def iss_generic_word(X, word: str, extended: int, semiring: str):
    """Iterated sums of a word over Python letters, ``[extended, n, t]``:
    Semiring._iterated_sum (fruits/iss/semiring.py:54-75: Reals -- identity ones,
    product, cumsum over ``tmp[k:]`` behind a shift) and Arctic._iterated_sum
    (:428-446: identity zeros, sum, running maximum, no shift).  Weightings do
    not reach this function in the reference (:40)."""
    if semiring not in ("reals", "arctic", "bayesian"):
        raise NotImplementedError(semiring)
    els = [[(part.split("(")[0], int(part.split("(")[1]) - 1)
            for part in el[1:].split(")")[:-1]] for el in word.split("]")[:-1]]
    n, _, t = X.shape
    out = np.zeros((n, extended, t))
    for i in range(n):
        Z = X[i]
        tmp = np.zeros(t) if semiring == "arctic" else np.ones(t)
        for k, el in enumerate(els):
            C = np.zeros(t) if semiring == "arctic" else np.ones(t)
            for name, dim in el:
                C = C + LETTERS[name](Z, dim) if semiring == "arctic" else C * LETTERS[name](Z, dim)
            if semiring in ("reals", "bayesian"):
                # (Bayesian inherits the general recursion WITH the shift, :54-75 -- unlike
                # its fast path for SimpleWords, :530-566 -- and a running maximum, :598-601)
                if k > 0:
                    tmp = np.roll(tmp, 1)
                    tmp[0] = 0
                tmp[k:] = tmp[k:] * C[k:]
                tmp[k:] = np.cumsum(tmp[k:]) if semiring == "reals" else np.maximum.accumulate(tmp[k:])
            else:
                tmp = np.maximum.accumulate(tmp + C)
            if len(els) - k <= extended:
                out[i, extended - (len(els) - k), :] = tmp.copy()
    return np.ascontiguousarray(np.swapaxes(out, 0, 1))


def coswiss_weightings(n_letters: int, exponent: int, total: bool) -> np.ndarray:
    """Expansion of ``cos(a-b)**exponent`` into products of powers of sines
    and cosines (fruits/iss/cos.py:265-287): row = (binomial coefficient
    product, sin/cos exponent of every level [, of the total weighting])."""
    p = n_letters + 1 if total else n_letters
    binom = [1]
    for k in range(exponent):
        binom.append(binom[-1] * (exponent - k) // (k + 1))
    # term k of one factor: coefficient C(e, k), cos^(e-k) sin^k spread over two levels
    rows = []
    for comb in itertools.product(range(exponent + 1), repeat=p - 1):
        w = np.zeros(2 * p + 1, dtype=np.int32)
        w[0] = 1
        for i, k in enumerate(comb):
            w[0] *= binom[k]
            w[2 * i + 1] += exponent - k
            w[2 * i + 3] += exponent - k
            w[2 * i + 2] += k
            w[2 * i + 4] += k
        rows.append(w)
    return np.array(rows, dtype=np.int32).reshape(len(rows), 2 * p + 1)


def coswiss_fit(iss, X):
    """State of a randomised CosWISS after ``fit`` (fruits/iss/cos.py:243-260):
    the uniform weights of the per-(word, frequency) two-layer network
    (``ffn_size``) and / or the dropped time steps per (word, frequency, level)
    (``dropout``); the draws in the reference's order."""
    c = iss["coswiss"]
    words = expand_words(iss["words"])
    nw, nf, d, t = len(words), len(c["freqs"]), X.shape[1], X.shape[2]
    st = {}
    if c.get("ffn_size") is not None:
        h = c["ffn_size"]
        st["A"] = np.random.random((nw, nf, h, d))
        st["b"] = np.random.random((nw, nf, h))
        st["C"] = np.random.random((nw, nf, d, h))
    if c.get("dropout") is not None:
        rate = int(c["dropout"] * t)
        depth = max(len(parse_word(w)) for w in words)
        st["drop"] = np.array([[[np.random.choice(t, size=(rate,), replace=False)
                                 for _ in range(depth)] for _ in range(nf)]
                               for _ in range(nw)], dtype=np.int32)
    return st


def coswiss_word(X, word: str, freqs, exponent: int, total: bool, ffn=None, drop=None):
    """``[n_freqs, n, t]`` cosine weighted iterated sums of one word
    (fruits/iss/cos.py:16-49, :171-181), same operation order.  ``ffn`` =
    ``(A, b, C)`` of this word: every frequency sees the input through its own
    two-layer network first (``_ffn`` :96-113, ``_ffn_coswiss`` :116-138);
    ``drop[f][k]``: time steps set to zero ahead of the cumulative sum of level
    ``k`` (``_leaky_coswiss_single`` :52-93).  The network wins if both are set
    (:306-324)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, d, t = X.shape
    X_raw = X
    mat = parse_word(word)
    p = mat.shape[0]
    wts = coswiss_weightings(p, exponent, total)
    out = np.zeros((len(freqs), n, t))
    for f, freq in enumerate(freqs):
        den = float(np.float32(freq)) * (t - 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            arg = np.pi * np.arange(t) / den
        sin_w, cos_w = np.sin(arg), np.cos(arg)
        if ffn is not None:
            A, b, C = ffn[0][f], ffn[1][f], ffn[2][f]
            Y = (A[np.newaxis, :, :, np.newaxis] * X_raw[:, np.newaxis, :, :]).sum(axis=2) \
                + b[np.newaxis, :, np.newaxis]
            Y = Y * (Y > 0)
            X = (C[np.newaxis, :, :, np.newaxis] * Y[:, np.newaxis, :, :]).sum(axis=2)
        for row in wts:
            tmp = np.ones((n, t))
            for k in range(p):
                if k > 0:
                    tmp = np.roll(tmp, 1, axis=1)
                    tmp[:, 0] = 0
                for letter, occ in enumerate(mat[k]):
                    for _ in range(abs(int(occ))):
                        tmp = tmp * X[:, letter, :] if occ > 0 else tmp / X[:, letter, :]
                for _ in range(row[2 * k + 1]):
                    tmp = tmp * sin_w
                for _ in range(row[2 * k + 2]):
                    tmp = tmp * cos_w
                if drop is not None and ffn is None:
                    tmp[:, drop[f][k]] = 0
                tmp = np.cumsum(tmp, axis=1)
            if total:
                for _ in range(row[2 * p + 1]):
                    tmp = tmp * sin_w
                for _ in range(row[2 * p + 2]):
                    tmp = tmp * cos_w
            out[f] += row[0] * tmp
    return out


def iss_iter(X, iss, cache: RawCache):
    """Yield ``[n, t]`` arrays in emission order for one ISS description."""
    words = expand_words(iss["words"])
    if iss.get("coswiss") is not None:
        c = iss["coswiss"]
        st = iss.get("_state") or {}
        if (c.get("ffn_size") is not None or c.get("dropout") is not None) and not st:
            raise RuntimeError("randomised CosWISS: coswiss_fit first")
        for i, w in enumerate(words):      # word-major, frequency-minor (fruits/iss/cos.py:289-333)
            out = coswiss_word(X, w, c["freqs"], c.get("exponent", 2), c.get("total", False),
                               ffn=(st["A"][i], st["b"][i], st["C"][i]) if "A" in st else None,
                               drop=st["drop"][i] if "drop" in st else None)
            for f in range(out.shape[0]):
                yield out[f]
        return
    extended = iss.get("mode", "single") == "extended"
    alphas = iss.get("alphas")
    if iss.get("semiring") == "arctic_argmax":
        if not extended:            # fruits/iss/iss.py:37-40
            raise NotImplementedError("Arctic argmax is not implemented when using ISSMode.SINGLE")
        for i, w in enumerate(words):
            out = arctic_argmax_word(X, w, iss.get("weighting"),
                                     None if alphas is None else alphas[i], cache)
            for e in range(out.shape[0]):
                yield out[e]
        return
    plan = cache_plan(words) if extended else [1] * len(words)
    for i, w in enumerate(words):
        if plan[i] == 0:
            continue
        if is_generic_word(w):
            out = iss_generic_word(X, w, plan[i], iss.get("semiring", "reals"))
        else:
            out = iss_word(X, w, plan[i], iss.get("semiring", "reals"),
                           iss.get("weighting"),
                           None if alphas is None else alphas[i], cache)
        for e in range(out.shape[0]):
            yield out[e]


def n_iterated_sums(iss):
    words = expand_words(iss["words"])
    if iss.get("coswiss") is not None:
        return len(words) * len(iss["coswiss"]["freqs"])
    if iss.get("semiring") == "arctic_argmax":      # fruits/iss/iss.py:139-144
        return sum(len(parse_word(w)) + len(parse_word(w)) * (len(parse_word(w)) + 1) // 2
                   for w in words)
    if iss.get("mode", "single") == "extended":
        return sum(cache_plan(words))
    return len(words)


def iss_label(iss, index):
    # fruits/iss/iss.py:195-204, fruits/iss/cache.py:55-66
    words = expand_words(iss["words"])
    if iss.get("coswiss") is not None:
        c = iss["coswiss"]
        d, r = divmod(index, len(c["freqs"]))
        return (f"{words[d]}!{c['freqs'][r]} : ^{c.get('exponent', 2)}"
                + (" : total" if c.get("total", False) else ""))
    if iss.get("mode", "single") == "extended":
        plan = cache_plan(words)
        for i, w in enumerate(words):
            index -= plan[i]
            if index < 0:
                return "]".join(w.split("]")[:int(index)]) + "]"
    return words[index]


def iterate_iss(X, iss_list, cache, idx=0):
    # fruits/fruit.py:440-454
    if idx == len(iss_list):
        yield X[:, 0, :]
    else:
        for itsum in iss_iter(X, iss_list[idx], cache):
            yield from iterate_iss(itsum[:, np.newaxis, :], iss_list, cache,
                                   idx + 1)


# ---------------------------------------------------------------------------
# sieves

# AVG and STD run CUR's backend in the reference (fruits/sieving/segment.py:303-307, :346-350)
_SEGMENT_KINDS = {"NPI": 0, "MPI": 1, "MAX": 2, "MIN": 3, "XPI": 4, "LPI": 5,
                  "CUR": 6, "AVG": 6, "STD": 6}
_IMPLICIT = ("PPV", "CPV")
_INCREMENT_SIEVES = ("NPI", "MPI", "XPI", "LPI")


class OracleWrapper:
    """fruits/sieving/wrapper.py: INC (:9-64) / INT (:67-104).  The wrapped
    sieve runs through its public fit / transform, i.e. with a cache of its
    own built from the wrapped input (fruits/seed.py:26-51)."""

    def __init__(self, desc):
        self.name, args = desc
        self.args = dict(args)
        self.inner = make_sieve(self.args["sieve"])
        self.depth = self.args.get("depth", 1)
        self.shift = self.args.get("shift", 1)

    def nfeatures(self):
        return self.inner.nfeatures()

    def requires_fitting(self):
        return self.inner.requires_fitting()

    def _wrapped(self, X):
        if self.name == "INT":
            return np.cumsum(X, axis=1)
        inc = X[:, np.newaxis, :]
        for _ in range(self.depth):          # wrapper.py:44-45: always from X
            inc = increments(X[:, np.newaxis, :], self.shift)
        return np.ascontiguousarray(inc[:, 0, :])

    def fit(self, X):
        self.inner.fit(self._wrapped(X))

    def transform(self, X, cache=None):
        W = self._wrapped(X)
        return self.inner.transform(W, RawCache(W[:, np.newaxis, :]))

    @property
    def quantiles(self):
        return self.inner.quantiles

    def label(self, index):
        return f"{self.name} of {self.inner.label(index)}"


def make_sieve(desc):
    return OracleWrapper(desc) if desc[0] in ("INC", "INT") else OracleSieve(desc)


class OracleSieve:
    def __init__(self, desc):
        self.name, args = desc
        args = dict(args)
        self.args = args
        if self.name in _IMPLICIT:
            q = args.get("quantile", 0.5)
            c = args.get("constant", False)
            q = q if isinstance(q, list) else [q]
            c = c if isinstance(c, list) else [c] * len(q)
            self.segments = args.get("segments", False)
            qc = list(zip(q, c))
            if self.segments:
                qc = sorted(zip(list(set(q)), c), key=lambda x: x[0])
            self.q_c = qc
            self.sample_size = args.get("sample_size", 1.0)
        else:
            cut = args.get("cut", -1)
            self.cut = tuple(cut) if isinstance(cut, (list, tuple)) else (cut,)
            default_q = (0.0, 1.0) if self.name in _INCREMENT_SIEVES else (-1.0, 1.0)
            q = args.get("q")
            self.q = tuple(q) if q is not None else default_q
            self.inc = args.get("inc", 1) if self.name in _INCREMENT_SIEVES else 0
            self.norm = args.get("coquantile_norm", "L2")

    # -- bookkeeping -------------------------------------------------------
    def nfeatures(self):
        if self.name in _IMPLICIT:
            return len(self.q_c) - 1 if self.segments else len(self.q_c)
        return len(self.cut) * (len(self.q) - 1)

    def requires_fitting(self):
        if self.name in _IMPLICIT:
            return True
        return any(q not in (-1, 0, 1) for q in self.q)

    # -- numerics ----------------------------------------------------------
    def pre(self, X):
        # fruits/sieving/increment.py:63-71
        arr = X
        if self.name in _INCREMENT_SIEVES:
            if self.inc > 0:
                for _ in range(self.inc):
                    arr = increments(arr[:, np.newaxis, :], 1)[:, 0, :]
            elif self.inc < 0:
                for _ in range(-self.inc):
                    arr = np.cumsum(arr, axis=1)
        return np.ascontiguousarray(arr)

    def fit(self, X):
        if self.name in _IMPLICIT:
            # fruits/sieving/implicit.py:99-112
            self.fitted_q = [x[0] for x in self.q_c]
            for i, q in enumerate(list(self.fitted_q)):
                if not self.q_c[i][1]:
                    sample_size = max(int(self.sample_size * len(X)), 1)
                    sel = np.random.choice(np.arange(len(X)), size=sample_size,
                                           replace=False)
                    self.fitted_q[i] = np.quantile(
                        np.array([X[j] for j in sel]).flatten(), q)
            return
        # fruits/sieving/segment.py:66-75
        arr = self.pre(X)
        qs = np.zeros(len(self.q))
        for i, q in enumerate(self.q):
            if q == 1.0:
                qs[i] = np.inf
            elif q == -1.0:
                qs[i] = -np.inf
            elif q != 0:
                qs[i] = np.quantile(arr, q)
        self.quantiles = np.sort(qs)

    def unfitted_quantiles(self):
        # fruits/sieving/segment.py:77-85 (not sorted there)
        qs = np.zeros(len(self.q))
        for i, q in enumerate(self.q):
            if q == 1.0:
                qs[i] = np.inf
            elif q == -1.0:
                qs[i] = -np.inf
            elif q != 0:
                raise RuntimeError("Sieve has not been fitted properly")
        self.quantiles = qs

    def cuts(self, X, cache: RawCache):
        # fruits/sieving/segment.py:51-64
        new = np.zeros((X.shape[0], len(self.cut) + 1))
        for i, cut in enumerate(self.cut):
            if isinstance(cut, float):
                new[:, i + 1] = cache.coquantile(cut, self.norm)
            else:
                new[:, i + 1] = cut if cut >= 0 else X.shape[1] + cut + 1
        return np.ascontiguousarray(np.sort(new).astype(np.int64))

    def transform(self, X, cache: RawCache):
        lib = load_oracle()
        X = np.ascontiguousarray(X, dtype=np.float64)
        n, t = X.shape
        if self.name in _IMPLICIT:
            q = np.ascontiguousarray(np.array(self.fitted_q, dtype=np.float64))
            res = np.zeros((n, self.nfeatures()))
            lib.fo_ppv(_p(X), _p(q), _p(res), n, t, len(q),
                       (1 if self.segments else 0) | (2 if self.name == "CPV" else 0))
            return res
        if not self.requires_fitting():
            self.unfitted_quantiles()
        arr = self.pre(X)
        cuts = self.cuts(arr, cache)
        if self.name == "END":
            # fruits/sieving/segment.py:210-219
            res = np.zeros((n, cuts.shape[1] - 1))
            for j in range(cuts.shape[1] - 1):
                res[:, j] = np.take_along_axis(arr, cuts[:, j + 1:j + 2] - 1,
                                               axis=1)[:, 0]
            return res
        q = np.ascontiguousarray(self.quantiles, dtype=np.float64)
        res = np.zeros((n, self.nfeatures()))
        if _SEGMENT_KINDS[self.name] == 6:
            # fruits/sieving/segment.py:246-249: cuts and thresholds of X applied to
            # the second-order increments
            arr = np.ascontiguousarray(
                increments(increments(arr[:, np.newaxis, :], 1), 1)[:, 0, :])
        lib.fo_segment_sieve(_p(arr), _p(cuts), _p(q), _p(res), n, t,
                             cuts.shape[1], len(q), _SEGMENT_KINDS[self.name])
        return res

    def label(self, index):
        if self.name in _IMPLICIT:
            return "PPV"
        r, m = divmod(index, len(self.q) - 1)
        lab = f"{self.name}!{self.cut[r]}![{self.q[m]}, {self.q[m + 1]}]"
        if self.name in _INCREMENT_SIEVES:
            lab = lab[:3] + f"[inc={self.inc}]" + lab[3:]
        return lab


# ---------------------------------------------------------------------------
# Fruit (fruits/fruit.py)

class OracleSlice:
    def __init__(self, spec):
        self.spec = spec
        self.sieves = [make_sieve(s) for s in spec["sieves"]]
        self.sieves_extended = []
        self.fit_sample_size = spec.get("fit_sample_size", 1)

    def niteratedsums(self):
        return int(np.prod([n_iterated_sums(i) for i in self.spec["iss"]]))

    def nfeatures(self):
        return sum(s.nfeatures() for s in self.sieves) * self.niteratedsums()

    def _sample(self, X):
        # fruits/fruit.py:430-438
        fs = self.fit_sample_size
        if isinstance(fs, int) and fs == 1:
            ind = np.random.randint(0, X.shape[0])
            return X[ind:ind + 1, :, :]
        s = max(int(fs * X.shape[0]), 1)
        idx = np.random.choice(X.shape[0], size=s, replace=False)
        return X[idx, :, :]

    def fit(self, X, cache):
        # fruits/fruit.py:456-496
        prepared = self._sample(X)
        self.prep_states = []
        for prep in self.spec.get("preps", []):
            self.prep_states.append(fit_prep_state(prep, prepared))
            prepared = apply_prep(prep, prepared, self.prep_states[-1], cache)
        for iss in self.spec["iss"]:          # fruits/fruit.py:478-481
            c = iss.get("coswiss")
            if c is not None and (c.get("ffn_size") is not None or c.get("dropout") is not None):
                iss["_state"] = coswiss_fit(iss, prepared)
        if not any(s.requires_fitting() for s in self.sieves):
            self.sieves_extended = []
            return
        self.sieves_extended = []
        for itsum in iterate_iss(prepared, self.spec["iss"], cache):
            copies = [make_sieve([s.name, s.args]) for s in self.sieves]
            for s in copies:
                s.fit(itsum)
            self.sieves_extended.append(copies)

    def transform(self, X, cache):
        # fruits/fruit.py:498-553
        prepared = X
        for prep, state in zip(self.spec.get("preps", []), self.prep_states):
            prepared = apply_prep(prep, prepared, state, cache)
        out = np.zeros((prepared.shape[0], self.nfeatures()))
        k = 0
        for i, itsum in enumerate(iterate_iss(prepared, self.spec["iss"], cache)):
            sieves = self.sieves_extended[i] if self.sieves_extended else self.sieves
            for s in sieves:
                nf = s.nfeatures()
                out[:, k:k + nf] = s.transform(itsum, cache)
                k += nf
        return out


class OracleFruit:
    def __init__(self, spec):
        self.spec = spec
        self.slices = [OracleSlice(s) for s in spec["slices"]]

    def nfeatures(self):
        return sum(s.nfeatures() for s in self.slices)

    def fit(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        cache = RawCache(X)
        for s in self.slices:
            s.fit(X, cache)

    def transform(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        cache = RawCache(X)
        res = np.zeros((X.shape[0], self.nfeatures()))
        i = 0
        for s in self.slices:
            k = s.nfeatures()
            res[:, i:i + k] = s.transform(X, cache)
            i += k
        return np.nan_to_num(res, copy=False, nan=0.0)

    def fit_transform(self, X):
        self.fit(X)
        return self.transform(X)
