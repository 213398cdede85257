"""The ISS seed (reference: ``fruits/iss/iss.py:14-204``).

``ISS.transform`` materialises the iterated sums ``[n_itsums, n_series,
length]`` with the CUDA kernel of ``csrc/lns.cuh``.  Inside a
:class:`~fruits_b200.fruit.FruitSlice` the same plan drives the fused kernel
that applies the sieves in registers and never writes this tensor.
"""
from enum import Enum, auto
from typing import Generator, Optional, Sequence

import numpy as np
import torch

from .. import _backend as be
from .._plan import DevicePlan, Trie
from ..cache import SharedSeedCache
from ..seed import Seed
from .cache import CachePlan
from .semiring import Arctic, Bayesian, Reals, Semiring
from .weighting import Weighting
from .words.word import SimpleWord, Word


class ISSMode(Enum):
    """Reference: iss.py:14-18."""
    SINGLE = auto()
    EXTENDED = auto()


class ISS(Seed):
    """Iterated sums signature of a list of words.

    Args:
        words: ``SimpleWord`` objects, or generic ``Word`` objects over Python
            letters (evaluated on the host, see ``_lettered``).
        mode: ``ISSMode.SINGLE`` (one iterated sum per word) or
            ``ISSMode.EXTENDED`` (additionally every prefix of every word
            that was not already emitted by an earlier word).
        semiring: ``Reals()`` (default), ``Arctic()`` or ``Bayesian()``.
        weighting: optional exponential weighting.
    """

    def __init__(self, words: Sequence[Word], /, *, mode: ISSMode = ISSMode.SINGLE,
                 semiring: Optional[Semiring] = None,
                 weighting: Optional[Weighting] = None) -> None:
        for w in words:
            if not isinstance(w, Word):
                raise TypeError(f"ISS takes words, got {type(w)}")
        self.words = words
        self.mode = mode
        self.semiring = semiring if semiring is not None else Reals()
        if not isinstance(self.semiring, (Reals, Arctic, Bayesian)):
            raise NotImplementedError(
                f"semiring {type(self.semiring).__name__} is not supported")
        # words over Python letters (reference: Semiring._iterated_sum,
        # semiring.py:54-75, Arctic :428-446)
        self._generic = any(not isinstance(w, SimpleWord) for w in words)
        if self._argmax and self._generic:
            raise NotImplementedError("Arctic(argmax=True) takes SimpleWords only")
        if self._generic:
            if any(len(w) == 0 for w in words):
                raise NotImplementedError("a word needs at least one extended letter")
        self._cache_plan = CachePlan(self.words if mode == ISSMode.EXTENDED else [])
        self.weighting = weighting
        self._trie_memo = None
        self._plan_memo: dict = {}

    @property
    def requires_fitting(self) -> bool:
        return False

    @property
    def _argmax(self) -> bool:
        return isinstance(self.semiring, Arctic) and self.semiring._argmax

    @property
    def _fusable_iss(self) -> bool:
        # the Bayesian semiring has no generic trie kernel: unweighted it is compiled
        # into the generated kernel (_jit.py), weighted it is sieved on materialised sums;
        # generic words are fused through their SimpleWord twin (FruitSlice)
        if self._generic or self._argmax:
            return False
        return not isinstance(self.semiring, Bayesian) or self.weighting is None

    # -- words over Python letters -----------------------------------------------
    def _lettered(self, X: torch.Tensor):
        """``(Xs, twin)`` for an ISS with generic words: every distinct extended
        letter is evaluated once for the batch -- on the host, series by series,
        because a letter is the caller's Python function of one series
        ``[n_dims, length]`` (reference: semiring.py:61-66) -- and appended to X as
        one more input dimension; ``twin`` is the ISS of SimpleWords over those
        dimensions.  The semiring's operation is applied once per extended letter
        (``tmp * C``, ``tmp + C``), exactly what a single-occurrence letter of a
        SimpleWord does in the kernels, so every route (trie kernel, block scan,
        generated kernels) serves generic words bit for bit like the reference.
        Weightings are ignored for generic words, like in the reference
        (semiring.py:40): in an ISS that also holds SimpleWords the twin keeps the
        weighting and the generic words get ``alpha = 0`` -- a factor ``exp(0) = 1.0``
        (Arctic: ``+ 0.0``) that changes no bit."""
        n, d, t = X.shape
        arctic = isinstance(self.semiring, Arctic)
        # Bayesian: the reference's general recursion keeps the shift between levels
        # (semiring.py:66-71), unlike its fast path for SimpleWords (:530-566).  With the
        # row of level k moved k steps to the left the shifted recursion IS the unshifted
        # one, D_k[u] = max_{v<=u} D_{k-1}[v] * C_k[v + k]; the emitted rows are moved back
        # k steps to the right (zeros in front) afterwards (_ShiftedRows).
        shifted = isinstance(self.semiring, Bayesian)
        Z = X.cpu().numpy()
        index, rows, base = {}, [], {}
        for w in self.words:
            if isinstance(w, SimpleWord):
                continue
            for k, el in enumerate(w._extended_letters):
                key = (str(el), k if shifted else 0)
                if key in index:
                    continue
                index[key] = d + len(rows)
                C = base.get(key[0])
                if C is None:
                    C = np.empty((n, t), dtype=np.float64)
                    fns = [el[q] for q in range(len(el))]
                    for i in range(n):
                        c = np.zeros(t) if arctic else np.ones(t)
                        for fn in fns:
                            c = c + fn(Z[i]) if arctic else c * fn(Z[i])
                        C[i] = c
                    base[key[0]] = C
                if key[1]:
                    moved = np.ones((n, t), dtype=np.float64)
                    moved[:, :max(t - k, 0)] = C[:, k:]
                    C = moved
                rows.append(C)
        extra = be.to_device(np.ascontiguousarray(np.stack(rows, axis=1)))
        Xs = torch.cat((X, extra), dim=1).contiguous()
        memo = self.__dict__.get("_twin_memo")
        sig = (d, tuple(str(w) for w in self.words), self.mode)
        if memo is None or memo[0] != sig:
            twins = [w if isinstance(w, SimpleWord) else SimpleWord(
                "".join(f"[({index[(str(el), k if shifted else 0)] + 1})]"
                        for k, el in enumerate(w._extended_letters)))
                for w in self.words]
            weighted = self.weighting is not None and any(
                isinstance(w, SimpleWord) for w in self.words)
            if weighted:
                for w, tw in zip(self.words, twins):
                    if tw is not w:
                        tw.alpha = np.zeros(len(tw), dtype=np.float32)
            twin = ISS(twins, mode=self.mode, semiring=self.semiring,
                       weighting=self.weighting if weighted else None)
            if shifted:
                shifts = []
                for i, w in enumerate(self.words):
                    ext = (self._cache_plan.unique_el_depth(i) if self.mode == ISSMode.EXTENDED
                           else 1)
                    generic = not isinstance(w, SimpleWord)
                    shifts += [k if generic else 0 for k in range(len(w) - ext, len(w))]
                twin = _ShiftedRows(twin, shifts)
            memo = (sig, twin)
            self._twin_memo = memo
        twin = memo[1]
        if hasattr(self, "_cache"):
            twin._cache = self._cache
        return Xs, twin

    # -- plan ------------------------------------------------------------------
    def _weight_mode(self) -> int:
        if self.weighting is None:
            return be.WEIGHT_NONE
        return be.WEIGHT_TOTAL if self.weighting.total else be.WEIGHT_NONTOTAL

    def _signature(self):
        # (evaluated on every transform call: no array is built for the default alphas)
        return (tuple(str(w) for w in self.words),
                tuple(None if w._alpha is None else w._alpha.tobytes() for w in self.words)
                if self.weighting is not None else None,
                self.mode, self.semiring._code, self._weight_mode())

    def trie(self, trusted: bool = False) -> Trie:
        """``trusted``: the caller validated the memo in this call already
        (chunk loops), skip rebuilding the signature of every word."""
        if trusted and self._trie_memo is not None:
            return self._trie_memo[1]
        sig = self._signature()
        if self._trie_memo is None or self._trie_memo[0] != sig:
            plan = self._cache_plan._plan if self.mode == ISSMode.EXTENDED else None
            self._trie_memo = (sig, Trie(self.words, plan, self.weighting is not None))
            self._plan_memo = {}
            self._piece_memo = {}
        return self._trie_memo[1]

    def _jit_trie(self, n_dims: int):
        """-> (trie, number of shared extra rows) for the kernel generator."""
        return self.trie(), 0

    def device_plan(self, rows_max: int, emit_range=None, dim_desc=None,
                    trusted: bool = False) -> DevicePlan:
        if isinstance(self.semiring, Bayesian):
            raise NotImplementedError("the Bayesian semiring has no trie kernel")
        trie = self.trie(trusted)
        key = (rows_max, emit_range, None if dim_desc is None else tuple(dim_desc))
        plan = self._plan_memo.get(key)
        if plan is None:
            sub = trie if emit_range is None else trie.subset(*emit_range)
            plan = DevicePlan(sub, self.semiring._code, self._weight_mode(), rows_max, dim_desc)
            self._plan_memo[key] = plan
        return plan

    def _dim_pieces(self, emit_range, trusted: bool = False) -> list:
        """Split ``emit_range`` (None = all) into consecutive emission ranges
        whose sub-tries reference at most ``FB_MAX_USED_DIMS`` distinct input
        dimensions and, in a weighted ISS, at most ``FB_MAX_ALPHAS`` distinct
        alpha values each (one range if the whole already does)."""
        trie = self.trie(trusted)
        lo, hi = (0, len(trie.emits)) if emit_range is None else emit_range
        memo = self.__dict__.setdefault("_piece_memo", {})
        key = (id(trie), lo, hi)
        if key in memo:
            return memo[key]
        weighted = self.weighting is not None

        def needs_of(v):
            dims, alphas = set(), set()
            while v >= 0:
                n = trie.nodes[v]
                dims.update(d for d, e in enumerate(n.expo) if e != 0)
                if weighted:
                    alphas.add(n.alpha)
                v = n.parent
            return dims, alphas

        pieces, start, used, used_a = [], lo, set(), set()
        for e in range(lo, hi):
            need, need_a = needs_of(trie.emits[e])
            if len(need) > be.FB_MAX_USED_DIMS:
                raise NotImplementedError(
                    f"one word references {len(need)} distinct dimensions; the kernel "
                    f"supports {be.FB_MAX_USED_DIMS}")
            if len(need_a) > be.FB_MAX_ALPHAS:
                raise NotImplementedError(
                    f"one word carries {len(need_a)} distinct alpha values; the kernel "
                    f"supports {be.FB_MAX_ALPHAS}")
            if (len(used | need) > be.FB_MAX_USED_DIMS
                    or len(used_a | need_a) > be.FB_MAX_ALPHAS):
                pieces.append((start, e))
                start, used, used_a = e, set(), set()
            used |= need
            used_a |= need_a
        pieces.append((start, hi))
        if len(pieces) == 1:
            pieces = [emit_range]
        memo[key] = pieces
        return pieces

    def max_dim(self) -> int:
        if self._generic:
            return max([len(el) for w in self.words if isinstance(w, SimpleWord) for el in w]
                       + [dim + 1 for w in self.words if not isinstance(w, SimpleWord)
                          for el in w._extended_letters for dim in el._dimensions] + [0])
        memo = self.__dict__.get("_max_dim_memo")
        if memo is None or memo[0] is not self.words or memo[1] != len(self.words):
            memo = (self.words, len(self.words),
                    max(len(el) for w in self.words for el in w))
            self._max_dim_memo = memo
        return memo[2]

    def n_iterated_sums(self) -> int:
        """Number of iterated sums ``transform`` returns (reference :135-150)."""
        if self._argmax:
            if self.mode != ISSMode.EXTENDED:
                raise NotImplementedError(
                    "Arctic argmax is not implemented when using ISSMode.SINGLE")
            return sum(self._argmax_rows(w) for w in self.words)
        if self.mode == ISSMode.EXTENDED:
            return self._cache_plan.n_iterated_sums()
        return len(self.words)

    @staticmethod
    def _argmax_rows(word) -> int:
        return len(word) + len(word) * (len(word) + 1) // 2

    def _materialize_argmax(self, X: torch.Tensor, emit_range, lookup) -> torch.Tensor:
        """``Arctic(argmax=True)``: per word the running maxima of every level and
        the positions that produced them (``fb_arctic_argmax_word``; reference:
        iss.py:41-58 with ``_arctic_argmax_single``, semiring.py:234-279).  Every
        word is computed from its first letter and contributes all its rows --
        the cache plan does not apply (iss.py:53-54)."""
        n, d, t = X.shape
        g, g_ld = self._lookup(X) if lookup is None else lookup
        lo, hi = (0, self.n_iterated_sums()) if emit_range is None else emit_range
        out = be.empty((hi - lo, n, t))
        first = 0
        for word in self.words:
            rows = self._argmax_rows(word)
            last = first + rows
            if last > lo and first < hi:
                mat = np.ascontiguousarray(np.array(list(word), dtype=np.int32))
                alpha = (np.asarray(word.alpha, dtype=np.float32) if self.weighting is not None
                         else np.zeros(len(mat), dtype=np.float32))
                mat_d = torch.from_numpy(mat).to(X.device)
                alpha_d = torch.from_numpy(np.ascontiguousarray(alpha)).to(X.device)
                work = be.empty((max(be.lib().fb_arctic_argmax_workspace(n, t, len(mat)), 8),),
                                dtype=torch.uint8)
                whole = first >= lo and last <= hi
                dst = out[first - lo:last - lo] if whole else be.empty((rows, n, t))
                be.check(be.lib().fb_arctic_argmax_word(
                    X.data_ptr(), n, d, t, mat_d.data_ptr(), mat.shape[0], mat.shape[1],
                    alpha_d.data_ptr(), be.ptr(g), g_ld, dst.data_ptr(), work.data_ptr(),
                    be.stream_ptr()))
                if not whole:
                    a, b = max(first, lo), min(last, hi)
                    out[a - lo:b - lo] = dst[a - first:b - first]
            first = last
        return out

    # -- execution ---------------------------------------------------------------
    def _lookup(self, X: torch.Tensor):
        """-> (g tensor or None, row stride)"""
        if self.weighting is None or self._generic:
            return None, 0
        if hasattr(self, "_cache"):
            self.weighting._cache = self._cache
        else:
            self.weighting._cache = SharedSeedCache(X)
        g, shared = self.weighting.get_lookup_device(X)
        g = g.contiguous()
        if g.shape[-1] != X.shape[2]:
            raise ValueError("weighting lookup has the wrong length")
        if not shared and g.shape[0] < X.shape[0]:
            raise ValueError("weighting lookup has fewer rows than the input")
        return g, (0 if shared else g.shape[-1])

    def _check_input(self, X: torch.Tensor) -> None:
        if X.dim() != 3:
            raise ValueError("input must have shape (n_series, n_dimensions, length)")
        if self.max_dim() > X.shape[1]:
            raise IndexError(
                f"words use dimension {self.max_dim()} but the input has {X.shape[1]}")

    def batch(self, X: torch.Tensor, g, g_ld, stats=None) -> be.FbBatch:
        b = be.FbBatch()
        b.X = X.data_ptr()
        b.n, b.d, b.t = X.shape
        b.g = be.ptr(g)
        b.g_ld = g_ld
        b.stats = be.ptr(stats)
        return b

    def materialize(self, X: torch.Tensor, emit_range=None, lookup=None,
                    trusted: bool = False) -> torch.Tensor:
        """Iterated sums ``[emit_hi-emit_lo, n, t]`` of the emissions in
        ``emit_range`` (all if None) on the device.  ``trusted``: input and
        plan memo were validated by the caller (second and later chunks)."""
        if not trusted:
            self._check_input(X)
        X = X.contiguous()
        if self._generic:
            Xs, twin = self._lettered(X)
            return twin.materialize(Xs, emit_range)
        if self._argmax:
            return self._materialize_argmax(X, emit_range, lookup)
        if isinstance(self.semiring, Bayesian):
            return self._materialize_scan(X, emit_range, lookup)
        if isinstance(self.semiring, Arctic) and self._arctic_scan(X.shape[0], emit_range):
            return self._materialize_scan(X, emit_range, lookup, arctic=True)
        rows = be.lib().fb_slice_rows(be.POLICY_MAT)
        pieces = self._dim_pieces(emit_range, trusted)
        if len(pieces) > 1:
            # the emissions reference more distinct dimensions than one launch
            # stages: consecutive emission ranges, each within the limit
            g_lookup = self._lookup(X) if lookup is None else lookup
            return torch.cat([self.materialize(X, piece, g_lookup, trusted=True)
                              for piece in pieces], dim=0)
        plan = self.device_plan(rows, emit_range, trusted=trusted)
        g, g_ld = self._lookup(X) if lookup is None else lookup
        out = be.empty((plan.n_emit, X.shape[0], X.shape[2]))
        import ctypes
        batch = self.batch(X, g, g_ld)
        be.check(be.lib().fb_iss_materialize(plan.byref(), ctypes.byref(batch), out.data_ptr(),
                                             be.stream_ptr()))
        return out

    def _arctic_scan(self, n_series: int, emit_range) -> bool:
        """Arctic sums through the block scan over T (``fb_arctic_word``, one CTA
        per series and word) instead of the lane-per-node interpreter?
        ``FRUITS_B200_ARCTIC_SCAN``: "1" always, "0" never.  Default, from
        ``scripts/arctic_scan_vs_serial.py`` on the 48-letter chains of
        ``fruit_general.py`` (same bits either way): the scan wins whenever a whole
        ISS of chain-like words is materialised in one call (1.0 vs 1.1 ms for one
        series, 8.9 vs 20.1 ms for 2,048 x 1,024); it recomputes every word from its
        first letter, so bushy tries and the emission-range chunks of ``fit`` keep
        the trie kernel."""
        import os
        mode = os.environ.get("FRUITS_B200_ARCTIC_SCAN", "auto")
        if mode in ("0", "1"):
            return mode == "1"
        if n_series <= 4:
            return True
        from .._jit_chain import chain_like
        return emit_range is None and n_series <= 4096 and chain_like(self.trie())

    def _materialize_scan(self, X: torch.Tensor, emit_range, lookup,
                          arctic: bool = False) -> torch.Tensor:
        """Word by word through the block scan over T -- ``fb_bayes_word``
        (reference: iss.py:42-63 with ``Bayesian.iterated_sum_fast``,
        semiring.py:530-566) or ``fb_arctic_word`` (``Arctic._iterated_sum_fast``,
        :354-404): word ``i`` emits its last ``extended_i`` prefixes (all but
        those an earlier word emitted)."""
        kernel = be.lib().fb_arctic_word if arctic else be.lib().fb_bayes_word
        n, d, t = X.shape
        g, g_ld = self._lookup(X) if lookup is None else lookup
        lo, hi = (0, self.n_iterated_sums()) if emit_range is None else emit_range
        out = be.empty((hi - lo, n, t))
        tabs = self.__dict__.setdefault("_bayes_tables", {})
        first = 0
        for i, word in enumerate(self.words):
            ext = (self._cache_plan.unique_el_depth(i) if self.mode == ISSMode.EXTENDED else 1)
            last = first + ext
            if ext > 0 and last > lo and first < hi:
                key = (i, str(X.device), tuple(float(a) for a in word.alpha)
                       if self.weighting is not None else None)
                if key not in tabs:
                    mat = np.ascontiguousarray(np.array(list(word), dtype=np.int32))
                    alpha = (np.asarray(word.alpha, dtype=np.float32) if self.weighting is not None
                             else np.zeros(len(mat), dtype=np.float32))
                    tabs[key] = (torch.from_numpy(mat).to(X.device),
                                 torch.from_numpy(np.ascontiguousarray(alpha)).to(X.device),
                                 mat.shape)
                mat, alpha, (p, md) = tabs[key]
                whole = first >= lo and last <= hi
                dst = out[first - lo:last - lo] if whole else be.empty((ext, n, t))
                be.check(kernel(
                    X.data_ptr(), n, d, t, mat.data_ptr(), p, md, alpha.data_ptr(), be.ptr(g),
                    g_ld, self._weight_mode(), ext, dst.data_ptr(), be.stream_ptr()))
                if not whole:
                    a, b = max(first, lo), min(last, hi)
                    out[a - lo:b - lo] = dst[a - first:b - first]
            first = last
        return out

    def iter_chunks(self, X: torch.Tensor, max_bytes: int = 1 << 30, emit_range=None,
                    rows_for_size: int = None):
        """Yield ``(emit_lo, tensor[g, n, t])`` over all emissions (or those in
        ``emit_range``), at most ``max_bytes`` per chunk (sized as if ``X`` had
        ``rows_for_size`` rows: ranks with different shards then cut the same
        chunks) and at most 65,535 emissions (one select launch)."""
        if self._generic:
            self._check_input(X)
            Xs, twin = self._lettered(X.contiguous())       # the letters once, not per chunk
            yield from twin.iter_chunks(Xs, max_bytes, emit_range, rows_for_size)
            return
        n_emit = self.n_iterated_sums()
        first, last = (0, n_emit) if emit_range is None else emit_range
        per = (X.shape[0] if rows_for_size is None else rows_for_size) * X.shape[2] * 8
        step = max(1, min(n_emit, max_bytes // max(per, 1), 65535))
        lookup = self._lookup(X.contiguous())
        for lo in range(first, last, step):
            hi = min(last, lo + step)
            rng = None if (lo == 0 and hi == n_emit) else (lo, hi)
            yield lo, self.materialize(X, rng, lookup, trusted=lo > first)

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        return self.materialize(X)

    def batch_transform(self, X, batch_size: int = 1) -> Generator:
        """Yields the iterated sums of ``batch_size`` words at a time
        (reference :152-185)."""
        if batch_size > len(self.words):
            raise ValueError("batch_size too large, has to be < len(words)")
        Xd = be.to_device(X)
        if self._generic:
            self._check_input(Xd)
            Xs, twin = self._lettered(Xd.contiguous())
            for res in twin.batch_transform(Xs, batch_size):
                yield res if isinstance(X, torch.Tensor) else res.cpu().numpy()
            return
        had_cache = hasattr(self, "_cache")
        if not had_cache:
            self._cache = SharedSeedCache(Xd)
        try:
            lookup = self._lookup(Xd.contiguous())
            i, e = 0, 0
            while i < len(self.words):
                nb = min(batch_size, len(self.words) - i)
                if self._argmax:
                    self.n_iterated_sums()                 # (raises in SINGLE mode)
                    ne = sum(self._argmax_rows(w) for w in self.words[i:i + nb])
                elif self.mode == ISSMode.EXTENDED:
                    ne = self._cache_plan.n_iterated_sums(range(i, i + nb))
                else:
                    ne = nb
                if ne > 0:
                    res = self.materialize(Xd, (e, e + ne), lookup)
                else:
                    res = be.empty((0, Xd.shape[0], Xd.shape[2]))
                yield res if isinstance(X, torch.Tensor) else res.cpu().numpy()
                i += nb
                e += ne
        finally:
            if not had_cache:
                del self._cache

    def _copy(self) -> "ISS":
        return ISS(self.words, mode=self.mode, semiring=self.semiring,
                   weighting=self.weighting)

    def _label(self, index: int) -> str:
        if self.mode == ISSMode.EXTENDED:
            string = self._cache_plan.get_word_string(index)
        else:
            string = str(self.words[index])
        if not isinstance(self.semiring, Reals):
            string += " : " + self.semiring.__class__.__name__
        if self.weighting is not None:
            string += " : " + self.weighting.__class__.__name__
        return string


class _ShiftedRows:
    """The SimpleWord twin of an ISS with generic words in the Bayesian semiring:
    emission ``e`` comes out of the kernels ``shifts[e]`` time steps early (see
    ``ISS._lettered``) and is moved back here, zeros in front."""

    def __init__(self, inner: ISS, shifts: list) -> None:
        self._inner, self._shifts = inner, shifts

    def __setattr__(self, name, value) -> None:
        if name == "_cache":
            self._inner._cache = value
        else:
            object.__setattr__(self, name, value)

    def _back(self, rows: torch.Tensor, first: int) -> torch.Tensor:
        t = rows.shape[2]
        for j in range(rows.shape[0]):
            k = self._shifts[first + j]
            if k:
                kept = rows[j, :, :max(t - k, 0)].clone()
                rows[j, :, :k] = 0.0
                rows[j, :, k:] = kept
        return rows

    def materialize(self, X, emit_range=None, lookup=None, trusted: bool = False):
        return self._back(self._inner.materialize(X, emit_range),
                          0 if emit_range is None else emit_range[0])

    def iter_chunks(self, X, max_bytes: int = 1 << 30, emit_range=None, rows_for_size=None):
        for lo, chunk in self._inner.iter_chunks(X, max_bytes, emit_range, rows_for_size):
            yield lo, self._back(chunk, lo)

    def batch_transform(self, X, batch_size: int = 1):
        first = 0
        for rows in self._inner.batch_transform(X, batch_size):
            yield self._back(rows, first)
            first += rows.shape[0]
