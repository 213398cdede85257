"""Cosine weighted ISS (reference: ``fruits/iss/cos.py:184-351``).

``CosWISS`` is an ISS over the real semiring whose summands are weighted with
``cos(pi |i-j| / (f (T-1)))**s`` for every pair of consecutive summation
indices.  The power of the cosine of a difference expands into products of
powers of ``sin`` and ``cos`` of the single indices (``_get_weightings``), so
every expansion term is an ordinary iterated sum with extra per-level
factors; ``csrc/cos.cu`` evaluates all terms of one word for all frequencies
in one launch.  Emission order: word-major, frequency-minor.

The randomised variants of the reference (``ffn_size``: every word and
frequency sees the input through its own random two-layer network; ``dropout``:
random time steps are zeroed ahead of every cumulative sum) are served per
(word, frequency) by the same kernels on a transformed input: the network output
(``fb_ffn``), or the input plus one 0/1 mask dimension per level that joins the
letters of the word (a factor 1.0 leaves a product unchanged bit for bit, a
factor 0.0 zeroes the summand like the reference's assignment does).
"""
import itertools
import os
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
    # and plans that do not fit run on materialised iterated sums (and so do the
    # randomised variants).
    _jit_only = True

    @property
    def _fusable_iss(self) -> bool:
        return self._ffn_size is None and self._dropout is None

    def __init__(self, words: Sequence[Word], freqs: Sequence[float], exponent: int = 2,
                 total_weighting: bool = False, ffn_size: Optional[int] = None,
                 dropout: Optional[float] = None) -> None:
        for word in words:
            if not isinstance(word, SimpleWord):
                raise ValueError("CosWISS only implemented for simple words")
        super().__init__(words)
        self._total_weighting = total_weighting
        self._freqs = freqs
        self._exponent = exponent
        self._ffn_size = ffn_size
        self._dropout = dropout
        self._tables: dict = {}

    @property
    def requires_fitting(self) -> bool:
        return self._ffn_size is not None or self._dropout is not None

    def _fit_device(self, X: torch.Tensor) -> None:
        """The reference's draws in its order (cos.py:243-260): uniform network
        weights per (word, frequency), then dropped time steps per (word, frequency,
        level)."""
        d, t = X.shape[1], X.shape[2]
        nw, nf = len(self.words), len(self._freqs)
        if (h := self._ffn_size) is not None:
            self._A = np.random.random((nw, nf, h, d))
            self._b = np.random.random((nw, nf, h))
            self._C = np.random.random((nw, nf, d, h))
        if (p := self._dropout) is not None:
            rate = int(p * t)
            self._dropout_indices = np.array([
                [[np.random.choice(t, size=(rate,), replace=False)
                  for _ in range(max(map(len, self.words)))]
                 for _ in range(nf)]
                for _ in range(nw)], dtype=np.int32)
        self._tables.pop("randomised", None)

    def n_iterated_sums(self) -> int:
        return len(self._freqs) * len(self.words)

    def _materialize_randomised(self, X: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
        """``ffn_size`` / ``dropout``: one plain CosWISS of one word and one
        frequency per emission, on the input that (word, frequency) sees
        (reference: ``_ffn_coswiss`` cos.py:116-138, ``_leaky_coswiss`` :141-160;
        the network wins if both are set, :306-324)."""
        ffn = self._ffn_size is not None
        if not hasattr(self, "_A" if ffn else "_dropout_indices"):
            raise RuntimeError("Missing call of self.fit")
        n, d, t = X.shape
        nf = len(self._freqs)
        out = be.empty((hi - lo, n, t))
        subs = self._tables.setdefault("randomised", {})
        for e in range(lo, hi):
            w, f = divmod(e, nf)
            word = self.words[w]
            if ffn:
                if self._A.shape[3] != d:
                    raise ValueError(f"CosWISS was fitted on {self._A.shape[3]} dimensions, got {d}")
                A, b, C = (torch.from_numpy(np.ascontiguousarray(a[w, f])).to(X.device)
                           for a in (self._A, self._b, self._C))
                Xe = be.empty((n, d, t))
                be.check(be.lib().fb_ffn(X.data_ptr(), 0, A.data_ptr(), b.data_ptr(),
                                         C.data_ptr(), Xe.data_ptr(), n, d, t,
                                         self._ffn_size, d, 0, be.stream_ptr()))
                masked = word
            else:
                if self._dropout_indices.shape[3] > t:
                    raise IndexError("CosWISS was fitted on longer series")
                p = len(word)
                mask = np.ones((p, t))
                for k in range(p):
                    mask[k, self._dropout_indices[w, f, k]] = 0.0
                extra = torch.from_numpy(mask).to(X.device)[None].expand(n, p, t)
                Xe = torch.cat((X, extra), dim=1).contiguous()
                # level k of the word additionally multiplies by mask dimension d + k
                masked = subs.get(("word", w))
                if masked is None or masked[0] != d:
                    text = "".join(
                        "[" + "".join((f"({i + 1})" if c > 0 else f"(-{i + 1})") * abs(int(c))
                                      for i, c in enumerate(el)) + f"({d + k + 1})]"
                        for k, el in enumerate(word))
                    masked = subs[("word", w)] = (d, SimpleWord(text))
                masked = masked[1]
            sub = subs.get((w, f, ffn, d))
            if sub is None:
                sub = subs[(w, f, ffn, d)] = CosWISS(
                    [masked], [self._freqs[f]], exponent=self._exponent,
                    total_weighting=self._total_weighting)
            out[e - lo] = sub.materialize(Xe)[0]
        return out

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

    # -- separable form ------------------------------------------------------------
    def _junction(self) -> list:
        """``[(coefficient, sin exponent, cos exponent)]`` of one junction between
        two consecutive summation indices, one entry per binomial term:
        ``cos(a-b)^s = sum_k C(s,k) (cos a cos b)^k (sin a sin b)^(s-k)`` -- the
        factor of index ``a`` goes to the level of ``a``, the factor of ``b`` to the
        level of ``b`` (reference :265-287, including its habit of keeping only
        the first decimal digit of every number)."""
        e = self._exponent
        binom = [1]
        for k in range(e):
            binom.append(binom[-1] * (e - k) // (k + 1))
        return [(int(str(binom[k])[0]), int(str(e - k)[0]), int(str(k)[0])) for k in range(e + 1)]

    def _separable_plan(self, n_dims: int):
        """The cosine weighted ISS as a recurrence with ``s+1`` states per level.

        The reference expands every word into ``(s+1)^(p-1)`` terms and computes
        each as an iterated sum of its own (:16-49).  The weight of a junction is a
        sum of ``s+1`` products ``u_k(i) v_k(j)``, so the sum over all terms
        factorises level by level: with ``A[i][k]`` = the sum of all partial terms
        whose junction ``i`` took binomial term ``k``,

            A[0][k][t] = cumsum( letter_0 * F(s_k, c_k) )
            A[i][k][t] = cumsum( letter_i * sum_k' d_k' F(s_k'+s_k, c_k'+c_k) * A[i-1][k'][t-1] )
            y[t]       = sum_k d_k F(s_k, c_k)[t] * A[p-1][k][t]            (total weighting)

        (``F(a, b) = sin^a cos^b`` of the level's own index; the last level of a
        word without total weighting has no junction to its right and a single
        state).  ``(s+1) p`` running sums per word and frequency instead of
        ``sum_i (s+1)^i``; prefixes shared between words are computed once.  Same
        value as the reference's expansion up to the order of the additions
        (1e-14 of the row maximum on the goldens).

        -> (nodes, emits, rows): ``nodes`` = ``[(level, letter, [(pred, row)])]``
        (``pred`` = index into nodes or -1, ``row`` = index into rows or -1),
        ``emits[w * nf + f]`` = ``[(row, node)]``, ``rows`` = ``[(f, coeff, a, b)]``."""
        key = ("separable",)
        if key in self._tables:
            return self._tables[key]
        J = self._junction()
        ns = len(J)
        rows, row_index = [], {}

        def row(f, coeff, a, b):
            k = (f, coeff, a, b)
            if k not in row_index:
                row_index[k] = len(rows)
                rows.append(k)
            return row_index[k]

        nodes, node_index, emits = [], {}, []

        def node(prefix, f, state, level, letter, preds):
            k = (prefix, f, state)
            if k not in node_index:
                node_index[k] = len(nodes)
                nodes.append((level, letter, preds))
            return node_index[k]

        for word in self.words:
            mat = [tuple(int(x) for x in el) for el in word]
            p = len(mat)
            nj = p if self._total_weighting else p - 1            # junctions of this word
            for f in range(len(self._freqs)):
                prev = None                                         # states of the previous level
                for i in range(p):
                    prefix = tuple(mat[:i + 1])
                    letter = tuple((d, e < 0) for d, e in enumerate(mat[i]) for _ in range(abs(e)))
                    if i <= nj - 1:                                 # a junction to the right
                        cur = []
                        for k in range(ns):
                            if prev is None:
                                preds = [(-1, row(f, 1, J[k][1], J[k][2]))]
                            else:
                                preds = [(prev[k2], row(f, J[k2][0], J[k2][1] + J[k][1],
                                                        J[k2][2] + J[k][2])) for k2 in range(ns)]
                            cur.append(node(prefix, f, k, i, letter, preds))
                    else:                                           # last level, not total
                        if prev is None:
                            preds = []
                        else:
                            preds = [(prev[k2], row(f, J[k2][0], J[k2][1], J[k2][2]))
                                     for k2 in range(ns)]
                        cur = [node(prefix, f, "last", i, letter, preds)]
                    prev = cur
                if nj == p:
                    emits.append([(row(f, J[k][0], J[k][1], J[k][2]), prev[k]) for k in range(ns)])
                else:
                    emits.append([(-1, prev[0])])
        # emission index = word-major, frequency-minor
        nf = len(self._freqs)
        order = [w * nf + f for w in range(len(self.words)) for f in range(nf)]
        assert order == list(range(len(emits)))
        self._tables[key] = (nodes, emits, rows)
        return self._tables[key]

    def _jit_trie(self, n_dims: int):
        """-> (trie, number of shared rows) for the kernel generator: the
        separable recurrence (``_separable_plan``) as nodes with several
        predecessors (``dp``), the weight rows as shared extra "dimensions"
        ``n_dims + r`` and every emission as a combination of the last level's
        states."""
        from .._plan import Trie, _Node
        key = ("jit", n_dims)
        if key in self._tables:
            return self._tables[key]
        nodes, emits, rows = self._separable_plan(n_dims)
        width = n_dims + len(rows)
        trie = object.__new__(Trie)
        trie.nodes, trie.emits = [], []

        def marker(dims):
            m = [0] * width
            for d in dims:
                m[d] = 1
            return tuple(m)

        for level, letter, preds in nodes:
            nd = _Node(-1, marker([d for d, _ in letter] + [n_dims + r for _, r in preds]),
                       0.0, level + 1)
            nd.dp = (letter, [(u, n_dims + r) for u, r in preds])
            trie.nodes.append(nd)
        # emissions in an order that keeps words with a common prefix together
        # (frequency-major, words sorted by their letters): parts of the generated
        # kernel then share the states of the prefix
        nf = len(self._freqs)
        mats = [tuple(tuple(int(x) for x in el) for el in w) for w in self.words]
        trie.emits = [None] * len(emits)
        for f in range(nf):
            for w in sorted(range(len(self.words)), key=lambda w: mats[w]):
                e = w * nf + f
                nd = _Node(-1, marker([n_dims + r for r, _ in emits[e] if r >= 0]), 0.0, 1, e)
                if emits[e][0][0] < 0:
                    nd.combo = [(1, emits[e][0][1], 0, 0, 0, 0)]
                else:
                    nd.combo = [(None, u, 0, n_dims + r, 0, 0) for r, u in emits[e]]
                trie.emits[e] = len(trie.nodes)
                trie.nodes.append(nd)
        trie.max_depth = max(n.depth for n in trie.nodes)
        self._tables[key] = (trie, len(rows))
        return self._tables[key]

    def _sep_table(self, dev) -> torch.Tensor:
        """Row indices of the separable recurrence for ``fb_coswiss_sep_word``:
        ``[n_freq][ns + ns*ns + ns]`` = first level, inner levels ``[k'][k]``,
        last level / output (the same for every word)."""
        key = ("septab", str(dev))
        if key not in self._tables:
            _, _, rows = self._separable_plan(0)
            index = {r: i for i, r in enumerate(rows)}
            J = self._junction()
            ns = len(J)
            tab = []
            for f in range(len(self._freqs)):
                # (rows a word set never uses are absent from the plan: any index will do)
                get = lambda *k: index.get((f,) + k, 0)          # noqa: E731
                tab += [get(1, J[k][1], J[k][2]) for k in range(ns)]
                tab += [get(J[k2][0], J[k2][1] + J[k][1], J[k2][2] + J[k][2])
                        for k2 in range(ns) for k in range(ns)]
                tab += [get(J[k][0], J[k][1], J[k][2]) for k in range(ns)]
            self._tables[key] = torch.tensor(np.asarray(tab, dtype=np.int32), device=dev)
        return self._tables[key]

    def _rows(self, X: torch.Tensor) -> torch.Tensor:
        """The weight rows of the separable form, ``[n_rows, t]`` on the device:
        ``coeff * sin^a cos^b`` of ``pi t / (f (t_len - 1))``."""
        _, _, rows = self._separable_plan(0)
        key = ("rows", str(X.device), X.shape[2])
        if key not in self._tables:
            spec = torch.tensor(np.asarray(rows, dtype=np.int32).reshape(-1, 4), device=X.device)
            freqs = torch.tensor(np.asarray(self._freqs, dtype=np.float32), device=X.device)
            # (one-letter words without total weighting have no junction: no row at all --
            # the kernels still get a valid pointer)
            out = be.zeros((max(len(rows), 1), X.shape[2]))
            if rows:
                be.check(be.lib().fb_cos_rows(freqs.data_ptr(), len(self._freqs), X.shape[2],
                                              spec.data_ptr(), len(rows), out.data_ptr(),
                                              be.stream_ptr()))
            self._tables = {k: v for k, v in self._tables.items() if k[0] != "rows"}
            self._tables[key] = out
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
        if self.requires_fitting:
            return self._materialize_randomised(X, lo, hi)
        out = be.empty((hi - lo, n, t))
        L = be.lib()
        ns = self._exponent + 1
        # separable recurrence (``_separable_plan``): exponent+1 states per level
        # instead of (exponent+1)^(p-1) expansion terms
        sep_ok = 2 <= ns <= 5 and os.environ.get("FRUITS_B200_COS_EXPANSION", "0") != "1"
        trig = None
        for w in range(lo // nf, -(-hi // nf)):
            mat, wts, mshape, wshape = self._word_tables(w, X.device)
            first, last = w * nf, (w + 1) * nf
            whole = first >= lo and last <= hi
            dst = out[first - lo:last - lo] if whole else be.empty((nf, n, t))
            if sep_ok and mshape[0] <= 6 and mshape[1] <= 8:
                be.check(L.fb_coswiss_sep_word(
                    X.data_ptr(), n, d, t, mat.data_ptr(), mshape[0], mshape[1],
                    self._rows(X).data_ptr(), self._sep_table(X.device).data_ptr(), nf, ns,
                    int(bool(self._total_weighting)), dst.data_ptr(), be.stream_ptr()))
            else:
                if trig is None:
                    trig = self._trig(X)
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
                       total_weighting=self._total_weighting, ffn_size=self._ffn_size,
                       dropout=self._dropout)

    def _label(self, index: int) -> str:
        d, r = divmod(index, len(self._freqs))
        string = str(self.words[d])
        string += f"!{self._freqs[r]} : ^{self._exponent}"
        if self._total_weighting:
            string += " : total"
        return string
