"""Cosine weighted ISS (reference: ``fruits/iss/cos.py:184-351``).

``CosWISS`` is an ISS over the real semiring whose summands are weighted with
``cos(pi |i-j| / (f (T-1)))**s`` for every pair of consecutive summation
indices.  The power of the cosine of a difference expands into products of
powers of ``sin`` and ``cos`` of the single indices (``_get_weightings``), so
every expansion term is an ordinary iterated sum with extra per-level
factors; ``csrc/cos.cu`` evaluates all terms of one word for all frequencies
in one launch.  Emission order: word-major, frequency-minor.

The randomised variants of the reference (``ffn_size``, ``dropout``) are out
of scope and raise ``NotImplementedError``.
"""
import itertools
from typing import Generator, Optional, Sequence

import numpy as np
import torch

from .. import _backend as be
from .iss import ISS
from .words.word import SimpleWord, Word


class CosWISS(ISS):
    """Args:
        words: ``SimpleWord`` objects.
        freqs: frequencies ``f``; one iterated sum per word and frequency.
        exponent: exponent ``s`` of the cosine (default 2).
        total_weighting: also weight the outermost sum.
    """

    # In a FruitSlice the expansion is compiled into the plan-specialised kernel
    # (``_jit_trie``); there is no generic fused kernel for it, so small batches
    # and plans that do not fit run on materialised iterated sums.
    _fusable_iss = True
    _jit_only = True

    def __init__(self, words: Sequence[Word], freqs: Sequence[float], exponent: int = 2,
                 total_weighting: bool = False, ffn_size: Optional[int] = None,
                 dropout: Optional[float] = None) -> None:
        for word in words:
            if not isinstance(word, SimpleWord):
                raise ValueError("CosWISS only implemented for simple words")
        if ffn_size is not None or dropout is not None:
            raise NotImplementedError(
                "the randomised CosWISS variants (ffn_size, dropout) are not built")
        super().__init__(words)
        self._total_weighting = total_weighting
        self._freqs = freqs
        self._exponent = exponent
        self._ffn_size = ffn_size
        self._dropout = dropout
        self._tables: dict = {}

    @property
    def requires_fitting(self) -> bool:
        return False

    def n_iterated_sums(self) -> int:
        return len(self._freqs) * len(self.words)

    def trie(self):
        raise NotImplementedError("CosWISS has no prefix trie of its own (see _jit_trie)")

    def _emit_costs(self) -> list:
        """Relative cost of every emitted sum (expansion terms x letters of its
        word): lets a multi-GPU fit give every rank the same amount of work."""
        costs = []
        for word in self.words:
            p = len(word) + 1 if self._total_weighting else len(word)
            costs += [float((self._exponent + 1) ** (p - 1) * len(word))] * len(self._freqs)
        return costs

    def _jit_trie(self, n_dims: int):
        """-> (trie, number of shared rows) for the kernel generator.

        Every expansion term of every (word, frequency) pair is a word over the
        input dimensions *augmented with the sin / cos rows of the frequency*:
        level k multiplies by x^e, then sin^a, then cos^b -- the reference's
        order, since the trig rows sort after the real dimensions.  The terms
        form a prefix trie like any word list; the emitted value is the
        combination ``sum_i coeff_i * S_i [* sin^a cos^b of the total
        weighting]`` (a virtual node with a ``combo`` list)."""
        from .._plan import Trie, _Node
        key = ("jit", n_dims)
        if key in self._tables:
            return self._tables[key]
        nf = len(self._freqs)

        class _Term(list):
            alpha = None

        terms, owner = [], []
        for w, word in enumerate(self.words):
            mat = [list(int(x) for x in el) + [0] * (n_dims - len(el)) for el in word]
            wts = self._get_weightings(word)
            p = len(mat)
            for f in range(nf):
                for r, row in enumerate(wts):
                    letters = []
                    for k in range(p):
                        trig = [0] * (2 * nf)
                        trig[2 * f], trig[2 * f + 1] = int(row[2 * k + 1]), int(row[2 * k + 2])
                        letters.append(mat[k] + trig)
                    terms.append(_Term(letters))
                    owner.append((w, f, r))
        trie = Trie(terms, None, False)
        term_node = list(trie.emits)
        for node in trie.nodes:
            node.emit = -1
        trie.emits = []
        i = 0
        for w, word in enumerate(self.words):
            wts = self._get_weightings(word)
            p = len(word)
            total = wts.shape[1] == 2 * p + 3
            for f in range(nf):
                combo = []
                for row in wts:
                    sp, cp = (int(row[2 * p + 1]), int(row[2 * p + 2])) if total else (0, 0)
                    combo.append((int(row[0]), term_node[i], sp, n_dims + 2 * f,
                                  cp, n_dims + 2 * f + 1))
                    i += 1
                # the marker exponents make the trig rows "used" dimensions
                marker = [0] * (n_dims + 2 * nf)
                marker[n_dims + 2 * f] = marker[n_dims + 2 * f + 1] = 1
                node = _Node(-1, tuple(marker), 0.0, 1, len(trie.emits))
                node.combo = combo
                trie.nodes.append(node)
                trie.emits.append(len(trie.nodes) - 1)
        self._tables[key] = (trie, 2 * nf)
        return self._tables[key]

    # -- expansion table ---------------------------------------------------------
    def _get_weightings(self, word: Word) -> np.ndarray:
        """``[n_terms, 2p+1]`` int32: binomial coefficient product, then the
        exponent of sin and of cos for every level (reference :265-287).  Like
        the reference, only one decimal digit of every binomial coefficient
        is used (exponents up to 4 are exact)."""
        p = len(word) + 1 if self._total_weighting else len(word)
        e = self._exponent
        binom = [1]
        for k in range(e):
            binom.append(binom[-1] * (e - k) // (k + 1))
        rows = np.zeros(((e + 1) ** (p - 1), 2 * p + 1), dtype=np.int32)
        rows[:, 0] = 1
        for c, comb in enumerate(itertools.product(range(e + 1), repeat=p - 1)):
            for i, k in enumerate(comb):
                rows[c, 0] *= int(str(binom[k])[0])
                rows[c, 2 * i + 1] += int(str(e - k)[0])
                rows[c, 2 * i + 3] += int(str(e - k)[0])
                rows[c, 2 * i + 2] += int(str(k)[0])
                rows[c, 2 * i + 4] += int(str(k)[0])
        return rows

    def _word_tables(self, index: int, dev):
        """Exponent matrix and expansion table of word ``index`` on the device.
        All words are uploaded together on first use (one copy, not two per
        word: every small pageable copy is a synchronisation)."""
        key = ("tables", str(dev))
        if key not in self._tables:
            mats = [np.ascontiguousarray(np.array(list(w), dtype=np.int32)) for w in self.words]
            wtss = [np.ascontiguousarray(self._get_weightings(w)) for w in self.words]
            flat = np.concatenate([a.ravel() for pair in zip(mats, wtss) for a in pair])
            blob = torch.from_numpy(flat).to(dev)
            views, off = [], 0
            for mat, wts in zip(mats, wtss):
                m = blob[off:off + mat.size]
                off += mat.size
                w = blob[off:off + wts.size]
                off += wts.size
                views.append((m, w, mat.shape + (int(np.abs(mat).sum(axis=1).max()),
                                           int(wts[:, 1:].max(initial=0))), wts.shape))
            self._tables[key] = views
        return self._tables[key][index]

    def _trig(self, X: torch.Tensor) -> torch.Tensor:
        freqs = torch.tensor(np.asarray(self._freqs, dtype=np.float32), device=X.device)
        trig = be.empty((len(self._freqs), 2, X.shape[2]))
        be.check(be.lib().fb_cos_trig(freqs.data_ptr(), len(self._freqs), X.shape[2],
                                      trig.data_ptr(), be.stream_ptr()))
        return trig

    # -- execution ---------------------------------------------------------------
    def _lookup(self, X: torch.Tensor):
        return None, 0

    def materialize(self, X: torch.Tensor, emit_range=None, lookup=None,
                    trusted: bool = False) -> torch.Tensor:
        """Iterated sums ``[emit_hi-emit_lo, n, t]`` (word-major, frequency-minor)."""
        if not trusted:
            self._check_input(X)
        X = X.contiguous()
        n, d, t = X.shape
        nf = len(self._freqs)
        lo, hi = (0, self.n_iterated_sums()) if emit_range is None else emit_range
        out = be.empty((hi - lo, n, t))
        trig = self._trig(X)
        L = be.lib()
        for w in range(lo // nf, -(-hi // nf)):
            mat, wts, mshape, wshape = self._word_tables(w, X.device)
            first, last = w * nf, (w + 1) * nf
            whole = first >= lo and last <= hi
            dst = out[first - lo:last - lo] if whole else be.empty((nf, n, t))
            be.check(L.fb_coswiss_word(X.data_ptr(), n, d, t, mat.data_ptr(), mshape[0],
                                       mshape[1], mshape[2], mshape[3], trig.data_ptr(), nf,
                                       wts.data_ptr(),
                                       wshape[0], wshape[1], dst.data_ptr(), be.stream_ptr()))
            if not whole:
                a, b = max(first, lo), min(last, hi)
                out[a - lo:b - lo] = dst[a - first:b - first]
        return out

    def batch_transform(self, X, batch_size: int = 1) -> Generator:
        """Yields the iterated sums of ``batch_size`` words at a time, each
        word contributing one array per frequency (reference :289-333)."""
        Xd = be.to_device(X)
        nf = len(self._freqs)
        i = 0
        while i < len(self.words):
            nb = min(batch_size, len(self.words) - i)
            res = self.materialize(Xd, (i * nf, (i + nb) * nf))
            yield res if isinstance(X, torch.Tensor) else res.cpu().numpy()
            i += nb

    def _copy(self) -> "CosWISS":
        return CosWISS(freqs=self._freqs, words=self.words, exponent=self._exponent,
                       total_weighting=self._total_weighting)

    def _label(self, index: int) -> str:
        d, r = divmod(index, len(self._freqs))
        string = str(self.words[d])
        string += f"!{self._freqs[r]} : ^{self._exponent}"
        if self._total_weighting:
            string += " : total"
        return string
