"""``Fruit`` and ``FruitSlice``: the pipeline container and the dispatch of the
hot path (reference: ``fruits/fruit.py``).

The reference walks a slice with one numba launch per word and per sieve
(fruit.py:538-550).  Here a slice is compiled into one device plan and --
whenever its preparateurs and sieves fit the fused kernel -- evaluated by a
single CUDA launch per slice that reads X once and writes only the features
(``csrc/lns.cuh``).  Everything else (several ISS in one slice, sieves with
several cuts or quantile intervals, callbacks, fit) runs on the composed
route: materialise chunks of iterated sums on the GPU and sieve them with the
array kernels of ``csrc/sieve.cu``.  There is no CPU route.
"""
import ctypes
import inspect
import os
from typing import Callable, Generator, Literal, Optional, Union

import numpy as np
import torch

from . import _backend as be
from . import _hoststage as hs
from . import _jit
from . import _jit_chain
from .cache import SharedSeedCache
from .iss.iss import ISS
from .iss.semiring import Reals
from .preparation.abstract import Preparateur
from .preparation.wrapper import NEW
from .seed import Seed
from .sieving.abstract import FeatureSieve, quantile_multi, quantile_rows
from .sieving.implicit import PPV
from .sieving.segment import SegmentSieve

_FEAT_CODE = {"CNT": be.FEAT_CNT, "AVG": be.FEAT_AVG, "PPV": be.FEAT_PPV,
              "MAX": be.FEAT_MAX, "MIN": be.FEAT_MIN, "END": be.FEAT_END,
              "XPI": be.FEAT_XPI, "LPI": be.FEAT_LPI, "CUR": be.FEAT_CUR, "CPV": be.FEAT_CPV}


class Fruit:
    """Feature extractor using iterated sums; a list of
    :class:`FruitSlice` objects whose features are concatenated
    (reference: fruit.py:14-277, same methods)."""

    def __init__(self, name: str = "") -> None:
        self.name: str = name
        self._slices: list = []
        self._slc_index: int = 0
        self._fitted: bool = False
        self._iterator_index: int = -1

    def cut(self, slice: Optional["FruitSlice"] = None) -> None:
        """Adds a new (or the given) slice and switches to it."""
        if slice is None:
            slice = FruitSlice()
        self._slices.append(slice)
        self._slc_index = len(self._slices) - 1
        self._fitted = False

    def copycut(self) -> None:
        """Adds a deep copy of the current slice and switches to it."""
        self.cut(self.get_slice().deepcopy())

    def get_slice(self, index: Optional[int] = None) -> "FruitSlice":
        if index is None:
            return self._slices[self._slc_index]
        return self._slices[index]

    def switch_slice(self, index: int) -> None:
        if not (0 <= index < len(self._slices)):
            raise IndexError("Index has to be in [0, len(self)-1]")
        self._slc_index = index

    def add(self, *objects: Union[Seed, Callable[[], Seed]]) -> None:
        """Adds preparateurs, ISS or sieves to the current slice."""
        if len(self._slices) == 0:
            self.cut()
        self._slices[self._slc_index].add(*objects)
        self._fitted = False

    def nfeatures(self) -> int:
        return sum(slc.nfeatures() for slc in self._slices)

    def fit(self, X, cache: Optional[SharedSeedCache] = None) -> None:
        """Fits all slices (reference: fruit.py:121-136)."""
        Xd = X if self._fit_from_host(X) else be.to_device(X)
        cache_ = SharedSeedCache(Xd) if cache is None else cache
        for slc in self._slices:
            slc._fit_device(Xd, cache_)
        self._fitted = True

    def _fit_from_host(self, X) -> bool:
        """A host array whose fit samples are a small part of it stays on the
        host: only the sampled rows are uploaded (the default ``fit_sample_size``
        is one series)."""
        if isinstance(X, torch.Tensor) or not isinstance(X, np.ndarray) or X.ndim != 3:
            return False
        if X.dtype != np.float64:
            return False                   # be.to_device raises the reference's TypeError
        n = X.shape[0]
        for slc in self._slices:
            fs = slc.fit_sample_size
            if not (isinstance(fs, int) and fs == 1) and max(int(fs * n), 1) * 4 > n:
                return False
        return n > 0

    def transform(self, X, callbacks: Optional[list] = None,
                  cache: Optional[SharedSeedCache] = None, out=None):
        """Feature matrix ``[n_series, nfeatures]`` of all slices
        (reference: fruit.py:138-173).

        ``out`` (not in the reference) is an optional preallocated float64
        numpy array ``[n_series, nfeatures]`` that receives the features;
        with page-locked ``X`` and ``out`` the copies run asynchronously."""
        if callbacks is None:
            callbacks = []
        if not self._fitted:
            raise RuntimeError("Missing call of self.fit")
        if isinstance(X, torch.Tensor) and X.is_cuda:
            return self.transform_device(X, callbacks, cache)
        return self._transform_host(X, callbacks, cache, out)

    def _transform_host(self, X, callbacks, cache, out):
        """Host input: stream row chunks through the GPU so that the upload of
        chunk i+1 and the download of chunk i-1 overlap the kernels of chunk i
        (three streams, two device buffers per direction).  Pageable arrays --
        what a caller of the reference passes -- go through a ring of pinned
        staging buffers that copy threads fill and drain (``_hoststage``)."""
        Xh = X if isinstance(X, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(self._check_host(X)))
        if Xh.dtype != torch.float64:
            raise TypeError(f"input must be float64, got {Xh.dtype}")
        if Xh.dim() != 3:
            raise ValueError("input must have shape (n_series, n_dimensions, length)")
        n, nf = Xh.shape[0], self.nfeatures()
        if out is None:
            out = np.empty((n, nf), dtype=np.float64)
        elif out.shape != (n, nf) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array [n_series, nfeatures]")
        oh = torch.from_numpy(out)
        row_bytes = Xh[0].numel() * 8 + nf * 8 if n else 1
        chunk_mb = int(os.environ.get("FRUITS_B200_HOST_CHUNK_MB", "256"))
        rows = max(1, min(n, (chunk_mb << 20) // max(row_bytes, 1)))
        # every step of transform is independent per series, so row chunks can
        # be processed on their own (a user-supplied cache refers to all rows)
        single = bool(callbacks) or cache is not None or n <= rows or not all(
            p._row_independent_transform() for slc in self._slices
            for p in slc.get_preparateurs())       # (FUN: user code that sees the whole batch)
        if single:
            res = self.transform_device(be.to_device(Xh), callbacks, cache)
            oh.copy_(res)
            return out
        dev = be.require_cuda()
        main = torch.cuda.current_stream()
        up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        up.wait_stream(main)
        xb = [torch.empty((rows,) + tuple(Xh.shape[1:]), dtype=torch.float64, device=dev)
              for _ in range(2)]
        fb = [torch.empty((rows, nf), dtype=torch.float64, device=dev) for _ in range(2)]
        # pageable arrays go through pinned staging buffers filled / drained by copy
        # threads (_hoststage); pinned arrays are the DMA source / target themselves
        stage_in, stage_out = not Xh.is_pinned(), not oh.is_pinned()
        pin_in = [hs.pinned(f"in{b}", (rows,) + tuple(Xh.shape[1:])) for b in range(2)] \
            if stage_in else None
        pin_out = [hs.pinned(f"out{b}", (rows, nf)) for b in range(2)] if stage_out else None
        starts = list(range(0, n, rows))
        nchunks = len(starts)
        span = [(lo, min(n, lo + rows)) for lo in starts]
        x_free, f_free = [None] * 2, [None] * 2          # device buffers reusable (events)
        h2d_done, d2h_done = [None] * nchunks, [None] * nchunks
        fill = [None] * nchunks                           # futures: user array -> pin_in
        drain = [[] for _ in range(2)]                    # futures: pin_out -> user array

        def start_fill(i):
            if stage_in and i < nchunks:
                b, (lo, hi) = i % 2, span[i]
                if i >= 2:
                    h2d_done[i - 2].synchronize()         # the DMA has read pin_in[b]
                fill[i] = hs.copy_rows(pin_in[b][:hi - lo], Xh[lo:hi])

        def start_drain(i):
            if stage_out and 0 <= i < nchunks:
                b, (lo, hi) = i % 2, span[i]
                d2h_done[i].synchronize()
                drain[b] = hs.copy_rows(oh[lo:hi], pin_out[b][:hi - lo])

        start_fill(0)
        for i, (lo, hi) in enumerate(span):
            b = i % 2
            if stage_in:
                hs.finish(fill[i])
            with torch.cuda.stream(up):
                if x_free[b] is not None:
                    up.wait_event(x_free[b])
                src = pin_in[b][:hi - lo] if stage_in else Xh[lo:hi]
                xb[b][:hi - lo].copy_(src, non_blocking=True)
                h2d_done[i] = torch.cuda.Event()
                h2d_done[i].record(up)
            start_fill(i + 1)                             # host copy overlaps the queued GPU work
            main.wait_event(h2d_done[i])
            if f_free[b] is not None:
                main.wait_event(f_free[b])
            self.transform_device(xb[b][:hi - lo], None, None, out=fb[b][:hi - lo])
            x_free[b] = torch.cuda.Event()
            x_free[b].record(main)
            if stage_out:
                hs.finish(drain[b])                       # pin_out[b] (chunk i-2) has been copied out
            with torch.cuda.stream(down):
                down.wait_event(x_free[b])
                dst = pin_out[b][:hi - lo] if stage_out else oh[lo:hi]
                dst.copy_(fb[b][:hi - lo], non_blocking=True)
                d2h_done[i] = torch.cuda.Event()
                d2h_done[i].record(down)
                f_free[b] = d2h_done[i]
            start_drain(i - 1)                            # blocks until chunk i-1 has landed
        start_drain(nchunks - 1)
        for b in range(2):
            hs.finish(drain[b])
        main.wait_stream(down)
        main.synchronize()
        return out

    @staticmethod
    def _check_host(X):
        X = np.asarray(X)
        if X.dtype != np.float64:
            # the reference's numba signatures only accept float64
            raise TypeError(f"input must be float64, got {X.dtype}")
        return X

    def transform_device(self, Xd: torch.Tensor, callbacks=None, cache=None,
                         out: Optional[torch.Tensor] = None, multicast=None) -> torch.Tensor:
        """``transform`` on device tensors; ``out`` may be a preallocated
        ``[n, >= nfeatures]`` row-major tensor (or a view of rows of one).

        ``multicast`` = ``(address, row stride)`` of the same rows inside an
        NVSwitch multicast mapping (``parallel.PeerGather``): the features are
        then stored THERE -- into the matrix of every rank of the multicast group,
        this rank's ``out`` included -- by the generated kernels themselves, or
        copied there with ``fb_multimem_copy`` by the slices on other routes."""
        callbacks = callbacks or []
        cache_ = SharedSeedCache(Xd) if cache is None else cache
        n = Xd.shape[0]
        result = be.zeros((n, self.nfeatures())) if out is None else out
        index = 0
        for slc in self._slices:
            for callback in callbacks:
                callback.on_next_slice()
            k = slc.nfeatures()
            slc._transform_device(Xd, callbacks, cache_, result, index, sanitize=True,
                                  multicast=multicast)
            index += k
        return result

    def fit_transform(self, X, callbacks: Optional[list] = None):
        self.fit(X)
        return self.transform(X, callbacks=callbacks)

    def summary(self) -> str:
        """Text overview, two slices side by side (same layout as the
        reference's ``Fruit.summary``, fruit.py:186-213: the golden vectors hold
        its output for the BASELINE pipelines)."""
        bar, gap = "=" * 80, "-" * 38
        title = "Fruit" + (f" {self.name!r}" if self.name else "")
        title += f" -> Features: {self.nfeatures()}"
        lines = [bar, "<" + title.center(78) + ">", bar]
        blocks = [slc.summary().split("\n") for slc in self._slices]
        for left, right in zip(blocks[0::2], blocks[1::2]):
            height = max(len(left), len(right))
            left = left + [" " * 38] * (height - len(left))
            right = right + [" " * 38] * (height - len(right))
            lines.append(f"|{gap}||{gap}|")
            lines += [f"|{a}||{b}|" for a, b in zip(left, right)]
            lines.append(f"|{gap}||{gap}|")
        if len(blocks) % 2:
            lines.append(f"|{gap}|")
            lines += [f"|{row}|" for row in blocks[-1]]
            lines.append(f"|{gap}|")
        lines.append(bar)
        return "\n".join(lines)

    def _clone(self, suffix: str, how: str) -> "Fruit":
        twin = Fruit(self.name + suffix)
        for slc in self._slices:
            twin.cut(getattr(slc, how)())
        return twin

    def copy(self) -> "Fruit":
        """New fruit over the same seed objects, unfitted (fruit.py:215-220)."""
        return self._clone(" (Copy)", "copy")

    def deepcopy(self) -> "Fruit":
        """New fruit over copies of all seeds, unfitted (fruit.py:222-227)."""
        return self._clone(" (Deepcopy)", "deepcopy")

    def label(self, index: int,
              level: Literal["prepared", "iterated sums", "features"] = "features",
              verbose: Literal[1, 2] = 1) -> str:
        """Label of one feature / iterated sum / preparateur chain, counted over
        all slices (fruit.py:229-261)."""
        def size(slc):
            if level == "prepared":
                return len(slc.get_preparateurs())
            if level == "iterated sums":
                return slc.niteratedsums()
            return slc.nfeatures()

        if index < 0:
            raise RuntimeError("Label index out of range")
        for slc in self._slices:
            if index < size(slc):
                return slc.label(index, level, verbose)
            index -= size(slc)
        raise IndexError("Label index out of range")

    def __len__(self) -> int:
        return len(self._slices)

    def __iter__(self) -> "Fruit":
        self._iterator_index = -1
        return self

    def __next__(self) -> "FruitSlice":
        if self._iterator_index < len(self._slices)-1:
            self._iterator_index += 1
            return self._slices[self._iterator_index]
        raise StopIteration()

    def __getitem__(self, index: int) -> "FruitSlice":
        return self.get_slice(index)


# iterated sums materialised per fit chunk (the pre-transformed copies and the
# select workspace come on top: ~3x this in HBM)
_FIT_CHUNK_BYTES = 2 << 30


class FruitSlice:
    """One slice of a Fruit: preparateurs -> ISS -> sieves
    (reference: fruit.py:280-686, same methods)."""

    def __init__(self) -> None:
        self._preparateurs: list = []
        self._iss: list = []
        self._sieves: list = []
        # one list of fitted sieve copies per iterated sum
        self._sieves_extended: list = []
        self._fitted: bool = False
        self.fit_sample_size: Union[float, int] = 1
        self._thr_memo = None

    # -- configuration -----------------------------------------------------------
    def add_preparateur(self, preparateur: Preparateur) -> None:
        if not isinstance(preparateur, Preparateur):
            raise TypeError
        self._preparateurs.append(preparateur)
        self._fitted = False

    def get_preparateurs(self) -> list:
        return self._preparateurs

    def clear_preparateurs(self) -> None:
        self._preparateurs = []
        self._fitted = False

    def add_iss(self, iss: ISS) -> None:
        if not isinstance(iss, ISS):
            raise TypeError
        self._iss.append(iss)
        self._fitted = False

    def get_iss(self) -> list:
        return self._iss

    def clear_iss(self) -> None:
        self._iss = []
        self._sieves_extended = []
        self._fitted = False

    def add_sieve(self, sieve: FeatureSieve) -> None:
        if not isinstance(sieve, FeatureSieve):
            raise TypeError
        self._sieves.append(sieve)
        self._fitted = False

    def get_sieves(self) -> list:
        return self._sieves

    def clear_sieves(self) -> None:
        self._sieves = []
        self._sieves_extended = []
        self._fitted = False

    def add(self, *objects: Union[Seed, Callable[[], Seed]]) -> None:
        """Adds preparateurs, ISS or sieves (classes are instantiated)."""
        for obj in objects:
            if inspect.isclass(obj):
                obj = obj()
            if isinstance(obj, Preparateur):
                self.add_preparateur(obj)
            elif isinstance(obj, ISS):
                self.add_iss(obj)
            elif isinstance(obj, FeatureSieve):
                self.add_sieve(obj)
            else:
                raise TypeError(f"Cannot add variable of type {type(obj)}")

    def clear(self) -> None:
        self.clear_preparateurs()
        self.clear_iss()
        self.clear_sieves()
        self.fit_sample_size = 1

    def nfeatures(self) -> int:
        return sum(s.nfeatures() for s in self._sieves) * self.niteratedsums()

    def niteratedsums(self) -> int:
        return int(np.prod([iss.n_iterated_sums() for iss in self._iss]))

    def _compile(self) -> None:
        if not self._iss:
            raise RuntimeError("No ISS given")
        if not self._sieves:
            raise RuntimeError("No feature sieves given")

    def _select_fit_sample(self, X) -> torch.Tensor:
        """Rows to fit on, on the device (``X``: device tensor or host array).
        Same draws from the global numpy RNG as the reference (fruit.py:430-438)."""
        if isinstance(self.fit_sample_size, int) and self.fit_sample_size == 1:
            ind = np.random.randint(0, X.shape[0])
            return be.to_device(X[ind:ind+1, :, :])
        s = max(int(self.fit_sample_size * X.shape[0]), 1)
        indices = np.random.choice(X.shape[0], size=s, replace=False)
        if not isinstance(X, torch.Tensor):
            return be.to_device(X[indices])
        idx = torch.as_tensor(indices, device=X.device, dtype=torch.long)
        return X.index_select(0, idx)

    # -- which route ---------------------------------------------------------------
    def _fused_dims(self, n_dims: int):
        """Per prepared dimension ``(raw_dim, inc, std)`` if the preparateurs
        can be applied while the kernel loads X, else None."""
        dims = [(d, 0, 0) for d in range(n_dims)]
        for prep in self._preparateurs:
            if isinstance(prep, NEW):
                inner = prep._preparateur
                if any(std for _, _, std in dims):
                    return None
                if inner is None:
                    dims = dims + dims
                elif inner._fusable() == "inc" and all(i == 0 for _, i, _ in dims):
                    dims = dims + [(d, 1, 0) for d, _, _ in dims]
                else:
                    return None
            else:
                kind = prep._fusable()
                if kind == "inc" and all(i == 0 and s == 0 for _, i, s in dims):
                    dims = [(d, 1, 0) for d, _, _ in dims]
                elif kind == "std" and all(s == 0 for _, _, s in dims):
                    dims = [(d, i, 1) for d, i, _ in dims]
                else:
                    return None
        return dims

    def _fused_sieves(self):
        """``(features, bounded_hi, bounded_mm)`` if all sieves fit the fused
        kernel, else None.  ``features`` holds one ``(kind, arg)`` per sieve."""
        feats, unit_q = [], {}
        bounded_hi = bounded_mm = False
        cuts = {sv._cut_key() for sv in self._sieves if isinstance(sv, SegmentSieve)}
        if len(cuts) > 1:
            return None           # segment sieves with different cuts: one table per kernel
        self._fused_cut = next((sv for sv in self._sieves if isinstance(sv, SegmentSieve)
                                and sv._cut_key() is not None), None)
        for sv in self._sieves:
            f = sv._fused()
            if f is None:
                return None
            kind, arg = f
            if kind in ("CNT", "AVG", "XPI", "LPI"):
                key = ("U", arg)             # one (lo, hi] interval per increment depth
                bounded_hi |= (sv._q[1] != 1.0) or (sv._q[0] > sv._q[1])
            elif kind in ("MAX", "MIN"):
                key = (kind, 0)
                bounded_mm |= tuple(sv._q) != (-1.0, 1.0)
            elif kind == "CUR":
                key = ("CUR", 0)             # (several CUR / AVG / STD sieves share one interval)
            else:
                key = None
            if key is not None:
                if unit_q.setdefault(key, tuple(sv._q)) != tuple(sv._q):
                    return None
            feats.append((_FEAT_CODE[kind], arg))
        if len(feats) > be.FB_MAX_FEATS:
            return None
        for code in (be.FEAT_PPV, be.FEAT_CPV):          # one fitted threshold column each
            if sum(1 for k, _ in feats if k == code) > 1:
                return None
        return feats, bounded_hi, bounded_mm

    def _is_fusable(self, n_dims: int, callbacks, length: int = 0) -> bool:
        if callbacks or len(self._iss) != 1:
            return False
        if length >= 65536:
            return False          # the fused kernels count in 16 bits: long series are sieved
                                  # on materialised iterated sums (same results)
        if not getattr(self._iss[0], "_fusable_iss", True):
            return False          # e.g. CosWISS: sieved on materialised iterated sums
        w = self._iss[0].weighting
        if w is not None and getattr(w, "_on_prepared", False):
            return False
        if self._fused_dims(n_dims) is None:
            return False
        self._fs_memo = self._fused_sieves()      # (_transform_fused takes it from here)
        return self._fs_memo is not None

    # -- fit -------------------------------------------------------------------------
    def fit(self, X, cache: Optional[SharedSeedCache] = None) -> None:
        """Fits the slice (reference: fruit.py:456-496)."""
        self._fit_device(be.to_device(X), cache)

    def _prepare_device(self, X: torch.Tensor, cache, fit: bool, callbacks=()):
        prepared = X
        for prep in self._preparateurs:
            prep._cache = cache
            if fit:
                prep._fit_device(prepared)
            prepared = prep._transform_device(prepared)
            for callback in callbacks:
                callback.on_preparateur(prepared.cpu().numpy())
        return prepared.contiguous()

    def _iterate_iss_device(self, X: torch.Tensor, iss_index: int = 0,
                            max_bytes: int = 1 << 30) -> Generator:
        """Yield every iterated sum ``[n, t]`` in the reference's order
        (fruit.py:440-454), materialised chunk by chunk on the GPU."""
        if iss_index == len(self._iss):
            yield X[:, 0, :]
        else:
            for _, chunk in self._iss[iss_index].iter_chunks(X, max_bytes):
                for e in range(chunk.shape[0]):
                    yield from self._iterate_iss_device(chunk[e][:, None, :], iss_index + 1,
                                                        max_bytes)

    def _fit_device(self, X: torch.Tensor, cache: Optional[SharedSeedCache] = None) -> None:
        self._compile()
        self._thr_memo = None
        if len(X.shape) != 3:
            raise ValueError("input must have shape (n_series, n_dimensions, length)")
        if cache is None:
            cache = SharedSeedCache(X)
        sample = self._select_fit_sample(X)
        prepared = self._prepare_device(sample, cache, fit=True)
        for iss in self._iss:
            iss._cache = cache
            iss._check_input(prepared)
            if iss.requires_fitting:            # (randomised CosWISS; fruit.py:478-481)
                iss._fit_device(prepared)
        if not any(sieve.requires_fitting for sieve in self._sieves):
            self._sieves_extended = []
            self._fitted = True
            return
        self._sieves_extended = []
        if len(self._iss) == 1:
            self._fit_batched(prepared, cache)
        else:
            for itsum in self._iterate_iss_device(prepared):
                sieves_copy = [sieve.copy() for sieve in self._sieves]
                for sieve in sieves_copy:
                    sieve._cache = cache
                    sieve._fit_device(itsum.contiguous())
                self._sieves_extended.append(sieves_copy)
        self._fitted = True

    def _fit_batched(self, prepared: torch.Tensor, cache) -> None:
        """Fit all sieve copies of one chunk of iterated sums with batched
        order statistics: one radix select per (increment depth, probability)
        over ``[chunk, n_fit * t]`` instead of one np.quantile per sieve."""
        iss = self._iss[0]
        n, _, t = prepared.shape
        n_emit = iss.n_iterated_sums()
        # multi-GPU fit (parallel.fit_sharded), either
        #  * ``_row_shard``: ``prepared`` holds only this rank's rows of the sample;
        #    every rank walks all iterated sums and the selections sum their
        #    histograms over the ranks (thresholds identical everywhere), or
        #  * ``_fit_shard``: every rank holds the whole sample and fits its
        #    contiguous share of the iterated sums; the fitted sieve copies (a few
        #    numbers each) are exchanged afterwards
        rs = getattr(self, "_row_shard", None)
        shard = getattr(self, "_fit_shard", None)
        n_all = n if rs is None else rs.n_sample            # rows of the whole fit sample
        first, last = (0, n_emit) if shard is None else shard[0](
            n_emit, getattr(iss, "_emit_costs", lambda: None)())
        has_ppv = any(isinstance(sv, PPV) for sv in self._sieves)

        def skip_draws(lo, hi):
            # keep the global numpy RNG in step with the reference (and with the
            # other ranks) for the iterated sums this rank does not fit
            if has_ppv:
                for _ in range(lo, hi):
                    for sv in self._sieves:
                        if isinstance(sv, PPV):
                            sv._draw(n_all)

        def qmulti(V, pairs):
            if rs is None:
                return quantile_multi(V, t, pairs)
            # (a rank without sample rows carries one placeholder row: zero values here)
            return rs.quantile_multi(V if rs.n_local else V[:, :0], t, pairs)

        def qrows(V, q, rows_global):
            if rs is None:
                return quantile_rows(V, q)
            return rs.quantile_rows(V if rs.n_local else V[:, :0], q, rows_global * t)

        skip_draws(0, first)
        local = []
        # (row-sharded: the chunk size follows the largest shard, so that all ranks
        # run the same number of chunks -- and of collectives)
        n_chunk = n if rs is None else rs.n_local_max(prepared.device)
        for _, chunk in iss.iter_chunks(prepared, max_bytes=_FIT_CHUNK_BYTES,
                                        emit_range=None if shard is None else (first, last),
                                        rows_for_size=n_chunk):
            G = chunk.shape[0]
            copies = [[sieve.copy() for sieve in self._sieves] for _ in range(G)]
            # replay the reference's RNG consumption: node-major, sieve order
            draws = [[sv._draw(n_all) if isinstance(sv, PPV) else None for sv in row]
                     for row in copies]
            pre_cache, q_cache = {}, {}

            def pretransformed(sv):
                key = sv._inc if isinstance(sv, SegmentSieve) else 0
                if key not in pre_cache:
                    flat = chunk.reshape(G * n, t)
                    pre_cache[key] = (flat if key == 0 else
                                      sv._pre_transform_device(flat)).reshape(G, n * t)
                return key, pre_cache[key]

            # every (increment depth, probability) this chunk needs, selected
            # together in three reads of the chunk (fb_order_stats_multi)
            pairs = []
            for sv in self._sieves:
                if isinstance(sv, PPV):
                    if max(int(sv._sample_size * n_all), 1) == n_all:
                        pairs += [(0, q) for q, const in sv._q_c_input if not const]
                elif isinstance(sv, SegmentSieve) and 0 <= sv._inc <= 2:
                    pairs += [(sv._inc, q) for q in sv._q if q not in (1.0, -1.0, 0)]
            pairs = sorted(set(pairs))
            if pairs:
                q_cache.update({pair: vals for pair, vals in
                                qmulti(chunk.reshape(G, n * t), pairs).items()
                                if vals is not None})

            def quant(sv, q):
                key = sv._inc if isinstance(sv, SegmentSieve) else 0
                if (key, q) not in q_cache:
                    # other depths (cumulative sums, > 2) and buckets of equal values
                    # too large for the candidate list: select on the materialised rows
                    q_cache[(key, q)] = qrows(pretransformed(sv)[1], q, n_all)
                return q_cache[(key, q)]

            for si, sieve in enumerate(self._sieves):
                if isinstance(sieve, PPV):
                    for e in range(G):
                        sv = copies[e][si]
                        sv._q = [x[0] for x in sv._q_c_input]
                        for qi, (q, const) in enumerate(sv._q_c_input):
                            if const:
                                continue
                            sel = draws[e][si][qi]
                            if len(sel) == n_all:
                                sv._q[qi] = quant(sv, q)[e]
                            else:
                                # a subsample of the sample rows (positions in the sample)
                                mine = sel if rs is None else rs.local_rows_of(np.sort(sel))
                                idx = torch.as_tensor(mine, device=chunk.device, dtype=torch.long)
                                rows = chunk[e].index_select(0, idx).reshape(1, -1)
                                sv._q[qi] = qrows(rows, q, len(sel))[0]
                elif isinstance(sieve, SegmentSieve):
                    need = [q for q in sieve._q if q not in (1.0, -1.0, 0)]
                    vals = {q: quant(sieve, q) for q in need}
                    for e in range(G):
                        copies[e][si]._set_quantiles({q: vals[q][e] for q in need})
                else:
                    if rs is not None:
                        raise NotImplementedError(
                            f"{type(sieve).__name__} cannot be fitted on a row-sharded sample")
                    for e in range(G):
                        copies[e][si]._fit_device(chunk[e])
            local.extend(copies)
        skip_draws(last, n_emit)
        if shard is not None:
            local = shard[1](local)          # all ranks' copies in emission order
        for row in local:
            for sv in row:
                sv._cache = cache
        self._sieves_extended.extend(local)

    # -- transform -------------------------------------------------------------------
    def transform(self, X, callbacks: Optional[list] = None,
                  cache: Optional[SharedSeedCache] = None):
        """Features of this slice (reference: fruit.py:498-553)."""
        if callbacks is None:
            callbacks = []
        if not self._fitted:
            raise RuntimeError("Missing call of self.fit")
        Xd = be.to_device(X)
        out = be.zeros((Xd.shape[0], self.nfeatures()))
        self._transform_device(Xd, callbacks, cache, out, 0, sanitize=False)
        if isinstance(X, torch.Tensor):
            return out
        return out.cpu().numpy()

    def _sieves_for(self, i: int) -> list:
        return self._sieves_extended[i] if self._sieves_extended else self._sieves

    def _threshold_table(self, n_emit: int) -> torch.Tensor:
        """``[n_emit, FB_NTHR]`` table of the fused kernel (layout in
        include/fruits_b200.h)."""
        if self._thr_memo is not None and self._thr_memo[0] == n_emit:
            return self._thr_memo[1]
        tab = np.zeros((n_emit, be.FB_NTHR))
        tab[:, 1::2] = np.inf
        tab[:, 7] = 0.0
        tab[:, 8] = -np.inf
        tab[:, 10] = -np.inf
        tab[:, 12] = -np.inf
        for e in range(n_emit):
            for sv in self._sieves_for(e):
                kind, arg = sv._fused()
                if kind in ("PPV", "CPV"):
                    if not hasattr(sv, "_q"):
                        raise RuntimeError(f"Missing call of {kind}.fit()")
                    tab[e, 6 if kind == "PPV" else 7] = sv._q[0]
                    continue
                if kind == "END":
                    continue
                if not sv.requires_fitting:
                    sv._get_unfitted_quantiles()
                q = sv._quantiles
                col = {"CNT": 2 * arg, "AVG": 2 * arg, "XPI": 2 * arg, "LPI": 2 * arg,
                       "MAX": 8, "MIN": 10, "CUR": 12}[kind]
                tab[e, col], tab[e, col + 1] = q[0], q[1]
        dev = be.to_device(np.ascontiguousarray(tab))
        self._thr_memo = (n_emit, dev)
        return dev

    def _transform_device(self, X: torch.Tensor, callbacks, cache, out: torch.Tensor,
                          col0: int, sanitize: bool, multicast=None) -> None:
        """Write the features of this slice into ``out[:, col0:col0+nfeatures]``
        (``multicast``: see ``Fruit.transform_device``)."""
        self._multicast = multicast
        try:
            self._transform_device_(X, callbacks, cache, out, col0, sanitize)
        finally:
            self._multicast = None
        if multicast is not None and X.shape[0] and \
                getattr(self, "_last_launch", ("",))[0] not in ("fb_jit_slice", "fb_jit_chain"):
            # this slice wrote ordinary memory: replicate its columns
            be.check(be.lib().fb_multimem_copy(
                out.data_ptr() + 8 * col0, out.stride(0), int(multicast[0]) + 8 * col0,
                int(multicast[1]), X.shape[0], self.nfeatures(), be.stream_ptr()))

    def _transform_device_(self, X: torch.Tensor, callbacks, cache, out: torch.Tensor,
                           col0: int, sanitize: bool) -> None:
        if not self._fitted:
            raise RuntimeError("Missing call of self.fit")
        if X.dim() != 3:
            raise ValueError("input must have shape (n_series, n_dimensions, length)")
        if cache is None:
            cache = SharedSeedCache(X)
        for iss in self._iss:
            iss._cache = cache
        if X.shape[0] == 0:
            return
        if self._is_fusable(X.shape[1], callbacks, X.shape[2]):
            self._transform_fused(X.contiguous(), cache, out, col0, sanitize)
        elif not self._transform_prepared_fused(X, callbacks, cache, out, col0, sanitize):
            self._transform_composed(X, callbacks or [], cache, out, col0, sanitize)

    def _transform_prepared_fused(self, X, callbacks, cache, out, col0, sanitize) -> bool:
        """Leading preparateurs the kernels cannot apply while they load the
        input (everything but ``INC`` / ``STD`` / ``NEW``) write a prepared copy
        with their own streaming kernel; the rest of the slice then takes the
        fused route on that copy, so the iterated sums still never reach HBM.
        An ISS over Python letters is replaced by its SimpleWord twin over the
        prepared input plus one dimension per extended letter (``ISS._lettered``).
        False if the slice cannot be fused behind any prefix."""
        saved, saved_iss = self._preparateurs, self._iss
        # (Bayesian generic words: their twin's rows are moved in time afterwards, so the
        # sieves have to see materialised rows)
        generic = (len(saved_iss) == 1 and getattr(saved_iss[0], "_generic", False)
                   and saved_iss[0].semiring._code != be.SEMIRING_BAYESIAN)
        if callbacks or not (saved or generic):
            return False
        try:
            k = len(saved)
            if not generic:
                for k in range(1, len(saved) + 1):
                    self._preparateurs = saved[k:]
                    if self._fused_dims(1) is not None:
                        break
                else:
                    return False
                if not self._is_fusable(X.shape[1], callbacks, X.shape[2]):
                    return False      # ISS or sieves keep the slice off the fused route anyway
            self._preparateurs = saved[k:]
            prepared = X
            for prep in saved[:k]:
                prep._cache = cache
                prepared = prep._transform_device(prepared)
            if prepared.dim() != 3:
                raise ValueError("preparateurs must return (n_series, n_dimensions, length)")
            if generic:
                saved_iss[0]._check_input(prepared)
                prepared, twin = saved_iss[0]._lettered(prepared.contiguous())
                self._iss = [twin]
            if not self._is_fusable(prepared.shape[1], callbacks, prepared.shape[2]):
                return False
            self._transform_fused(prepared.contiguous(), cache, out, col0, sanitize)
            return True
        finally:
            self._preparateurs, self._iss = saved, saved_iss

    def _transform_fused(self, X, cache, out, col0, sanitize) -> None:
        iss = self._iss[0]
        n, d, t = X.shape
        dims = self._fused_dims(d)
        if iss.max_dim() > len(dims):
            raise IndexError(
                f"words use dimension {iss.max_dim()} but the prepared input has {len(dims)}")
        if out.stride(1) != 1:
            raise ValueError("feature matrix must be row-major")
        feats, bounded_hi, bounded_mm = self.__dict__.pop("_fs_memo", None) or self._fused_sieves()
        # rank-2 accumulators and cuts exist in the thread-per-series kernel only
        # ... and so do the Bayesian sums
        jit_only = getattr(iss, "_jit_only", False) or self._fused_cut is not None or any(
            k in (be.FEAT_XPI, be.FEAT_LPI, be.FEAT_CUR, be.FEAT_CPV) for k, _ in feats) or \
            iss.semiring._code == be.SEMIRING_BAYESIAN
        compile_ok = _jit.enabled(n) or (getattr(iss, "_jit_only", False) and _jit.enabled())
        # mid-size batches: the generated kernel only if it has been compiled already
        cached_ok = not compile_ok and n >= _jit.MIN_SERIES_CACHED and _jit.enabled()
        if not compile_ok and _jit.enabled() and self._chain_first(iss, len(dims)):
            # the chain kernel runs one lane per trie node like the generic kernel, so
            # it pays from small batches on (one translation unit, ~1 s of NVRTC)
            compile_ok = n >= _jit_chain.MIN_SERIES
            cached_ok = not compile_ok and n >= _jit_chain.MIN_SERIES_CACHED
        if compile_ok or cached_ok:
            try:
                self._transform_jit(X, cache, out, col0, sanitize, dims, feats, bounded_hi,
                                    bounded_mm, cached_only=cached_ok)
                return
            except NotImplementedError:
                pass      # plan too large for the specialised kernel (or not compiled and
                          # the batch too small to pay for it): other route below
        if jit_only:
            self._transform_composed(X, [], cache, out, col0, sanitize)
            return
        try:
            self._transform_generic(X, out, col0, sanitize, dims, feats, bounded_hi, bounded_mm)
        except NotImplementedError:
            # e.g. words over more distinct dimensions than one kernel block stages:
            # materialise (in pieces) and sieve with the stand-alone kernels
            self._transform_composed(X, [], cache, out, col0, sanitize)

    @staticmethod
    def _chain_first(iss, n_dims: int) -> bool:
        if os.environ.get("FRUITS_B200_CHAIN", "1") == "0":
            return False
        trie, n_shared = iss._jit_trie(n_dims)
        return (not n_shared and _jit_chain.suitable(trie, iss.semiring._code, iss._weight_mode())
                and _jit_chain.chain_like(trie))

    def _transform_jit(self, X, cache, out, col0, sanitize, dims, feats, bounded_hi,
                       bounded_mm, cached_only: bool = False) -> None:
        """Plan-specialised kernel (``_jit.py``): trie nodes and sieve state in
        registers, one thread per series and trie part."""
        iss = self._iss[0]
        trie, n_shared = iss._jit_trie(len(dims))
        used = trie.used_dims()
        real = [u for u in used if u < len(dims)]
        # standardised dimensions: the prepared input is materialised once
        # (2 x 8 bytes per value against ~10^3 flop per value of ISS work)
        materialise = any(dims[u][2] for u in real)
        jdims = [((u, 0) if materialise else (dims[u][0], dims[u][1])) if u < len(dims)
                 else ("row", u - len(dims)) for u in used]
        cut_sv = getattr(self, "_fused_cut", None)
        sieves = _jit.SieveSet.make(feats, bounded_hi, bounded_mm, cut=cut_sv is not None)
        g, g_ld = iss._lookup(X)
        wm = iss._weight_mode()
        # unweighted Arctic plans: the lane-per-node chain kernel (``_jit_chain.py``) first
        # for chain-like tries (the alternating-sign words), else as the second choice
        mode = os.environ.get("FRUITS_B200_CHAIN", "1")
        kinds = ["slice"]
        if mode != "0" and not n_shared and _jit_chain.suitable(trie, iss.semiring._code, wm):
            first = mode == "force" or _jit_chain.chain_like(trie)
            kinds = ["chain", "slice"] if first else ["slice", "chain"]
        small = X.shape[0] < _jit.MIN_SERIES          # (layout choice of _jit.generate)
        base_key = (tuple(jdims), tuple(feats), bounded_hi, bounded_mm, g_ld == 0,
                    X.device.index,        # (a loaded module belongs to one device)
                    cut_sv is not None, small)
        memo = getattr(iss, "_jit_memo", None)
        if memo is None or memo[0] is not trie:
            memo = (trie, {})
            iss._jit_memo = memo

        def kernel_of(kind, chain_warps=None):
            chain = kind == "chain"
            copts = _jit_chain.options()
            if chain_warps is not None:
                copts["warps"] = chain_warps
            key = base_key + ((kind, tuple(sorted(copts.items()))) if chain
                              else (kind, _jit.options_key()))
            kern = memo[1].get(key)
            gen = None
            if isinstance(kern, tuple):         # ("not compiled", planned source)
                gen, kern = kern[1], None       # a large batch pays for the compilation, a
                                                # mid-size one looks for the cubin again
            if kern is None and gen is None:
                try:
                    if chain:
                        gen = _jit_chain.generate(trie, iss.semiring._code, wm, sieves, jdims, copts)
                    else:
                        # a plan without another fused route may keep part of its sums
                        # in local memory
                        spill = 450 if getattr(iss, "_jit_only", False) else 0
                        gen = _jit.generate(trie, iss.semiring._code, wm, sieves, jdims,
                                            g_ld == 0, _jit.options(), n_shared, spill, small)
                except NotImplementedError as exc:
                    memo[1][key] = exc          # remembered: planning is host work
                    raise
            if kern is None:
                try:
                    kern = (_jit_chain.JitChain if chain else _jit.JitSlice).load(gen, cached_only)
                except _jit.NotCompiled:
                    memo[1][key] = ("not compiled", gen)      # do not plan again on every call
                    raise
                memo[1][key] = kern
            if isinstance(kern, NotImplementedError):
                raise kern
            if chain and not kern.fits(X.shape[2]):
                # the whole series of a CTA sits in shared memory: fewer series per CTA
                if kern.em.spc > 1:
                    return kernel_of(kind, max(1, kern.em.nb * (kern.em.spc // 2)))
                raise NotImplementedError("series too long for the chain kernel's staging")
            return kern

        kern = None
        for kind in kinds:
            try:
                kern = kernel_of(kind)
                break
            except NotImplementedError:
                if kind == kinds[-1]:
                    raise
        chain = kind == "chain"
        if materialise:
            X = self._prepare_device(X, cache, fit=False)
        thr = self._threshold_table(len(trie.emits))
        thr_c = None
        if kern.cols:
            tkey = (id(thr), tuple(kern.cols))
            if getattr(self, "_thr_compact", (None, None))[0] != tkey:
                idx = torch.as_tensor(kern.cols, device=thr.device, dtype=torch.long)
                self._thr_compact = (tkey, thr.index_select(1, idx).contiguous(), thr)
            thr_c = self._thr_compact[1]
        extra, extra_ld = None, 0
        if n_shared:
            extra = iss._rows(X)              # CosWISS: weight rows, shared by all series
        elif wm != be.WEIGHT_NONE:
            rows = 1 if g_ld == 0 else X.shape[0]
            if isinstance(iss.semiring, Reals):
                alphas = (ctypes.c_float * len(kern.em.p.alphas))(*kern.em.p.alphas)
                extra = be.empty((rows, 2 * len(alphas), X.shape[2]))
                be.check(be.lib().fb_exp_rows(g.data_ptr(), extra.data_ptr(), rows, X.shape[2],
                                              alphas, len(alphas), be.stream_ptr()))
                extra_ld = 0 if g_ld == 0 else 2 * len(alphas) * X.shape[2]
            else:
                extra = g
                extra_ld = g_ld
        cuts = None
        if cut_sv is not None:
            # end of the sieved segment per series: the second column of the sorted cut
            # table [0, cut] (fruits/sieving/segment.py:51-64), float cuts from the cache
            cut_sv._cache = cache
            cuts = cut_sv._cuts_device(X.shape[0], X.shape[2])[:, 1].to(torch.int32).contiguous()
        kern.launch(X.contiguous(), extra, extra_ld, thr_c, out, col0, sanitize,
                    multicast=getattr(self, "_multicast", None), cuts=cuts)
        self._last_launch = ("fb_jit_chain" if chain else "fb_jit_slice",
                             kern.n_launches(X.shape[0], X.shape[2]), kern)

    def _transform_generic(self, X, out, col0, sanitize, dims, feats, bounded_hi,
                           bounded_mm) -> None:
        """Generic trie-interpreting kernel (``csrc/lns.cuh``)."""
        if getattr(self, "_fused_cut", None) is not None:
            raise NotImplementedError("cuts are not built into the generic kernel")
        iss = self._iss[0]
        L = be.lib()
        n, d, t = X.shape
        sp = be.FbSievePlan()
        sp.n_feats = len(feats)
        for f, (kind, arg) in enumerate(feats):
            sp.kind[f], sp.arg[f] = kind, arg
        policy = L.fb_slice_policy(ctypes.byref(sp), int(bounded_hi), int(bounded_mm))
        if policy < 0:
            raise NotImplementedError("sieve combination not supported by the fused kernel")
        plan = iss.device_plan(L.fb_slice_rows(policy), None, dims)
        thr = self._threshold_table(plan.n_emit)
        sp.thresholds = thr.data_ptr()
        # statistics of the standardised dimensions the words actually use
        stats = None
        if any(dims[u][2] for u in plan.used_dims):
            stats = be.zeros((n, len(plan.used_dims), 2))
            for ui, u in enumerate(plan.used_dims):
                raw, inc, std = dims[u]
                if not std:
                    continue
                row = X[:, raw, :].contiguous()
                if inc:
                    from .preparation.transform import increments_device
                    row = increments_device(row, 1)
                std_prep = next(p for p in self._preparateurs if p._fusable() == "std")
                st = std_prep._row_stats(row, std_prep._div_std, std_prep._eps)
                stats[:, ui, :] = st
        g, g_ld = iss._lookup(X)
        batch = iss.batch(X, g, g_ld, stats)
        be.check(L.fb_slice_features_ex(plan.byref(), ctypes.byref(batch), ctypes.byref(sp),
                                        out.data_ptr(), out.stride(0), col0, policy,
                                        int(sanitize), be.stream_ptr()))
        self._last_launch = ("fb::lns_kernel", 1, None)

    def _transform_composed(self, X, callbacks, cache, out, col0, sanitize) -> None:
        prepared = self._prepare_device(X, cache, fit=False, callbacks=callbacks)
        for callback in callbacks:
            callback.on_preparation_end(prepared.cpu().numpy())
        n = prepared.shape[0]
        nf_total = self.nfeatures()
        buf = be.zeros((n, nf_total))
        k = 0
        for i, itsum in enumerate(self._iterate_iss_device(prepared)):
            itsum = itsum.contiguous()
            for callback in callbacks:
                callback.on_iterated_sum(itsum.cpu().numpy())
            pre_cache = {}
            for sieve in self._sieves_for(i):
                sieve._cache = cache
                nf = sieve.nfeatures()
                if isinstance(sieve, SegmentSieve):
                    if not sieve.requires_fitting:
                        sieve._get_unfitted_quantiles()
                    if sieve._inc not in pre_cache:
                        pre_cache[sieve._inc] = sieve._pre_transform_device(itsum)
                    sieve._apply(pre_cache[sieve._inc], buf, k)
                elif isinstance(sieve, PPV):
                    if not hasattr(sieve, "_q"):
                        raise RuntimeError("Missing call of PPV.fit()")
                    sieve._apply(itsum, buf, k)
                else:
                    buf[:, k:k+nf] = sieve._transform_device(itsum)
                for callback in callbacks:
                    callback.on_sieve(buf[k:k+nf].cpu().numpy())
                k += nf
        for callback in callbacks:
            callback.on_sieving_end(buf.cpu().numpy())
        if sanitize:
            be.check(be.lib().fb_nan_to_num(buf.data_ptr(), buf.numel(), be.stream_ptr()))
        out[:, col0:col0 + nf_total] = buf
        self._last_launch = ("composed", 0, None)

    def fit_transform(self, X):
        self.fit(X)
        return self.transform(X)

    # -- text ------------------------------------------------------------------------
    def summary(self) -> str:
        """Text block of 38 columns (layout of the reference, fruit.py:561-597)."""
        def row(text=""):
            return f"{text:<38}"

        lines = [f"{f'FruitSlice -> {self.nfeatures()}':^38}", "-" * 38,
                 row(f"Preparateurs ({len(self._preparateurs)}):")]
        lines += [row(f"    + {prep}") for prep in self._preparateurs] or [row()]
        lines.append(row(f"ISS Calculators ({len(self._iss)}):") + ("" if self._iss else row()))
        for iss in self._iss:
            weighting = "None" if iss.weighting is None else type(iss.weighting).__name__
            lines += [row(f"    + {iss} -> {iss.n_iterated_sums()}"),
                      row(f"       | words: {len(iss.words)}"),
                      row(f"       | semiring: {type(iss.semiring).__name__}"),
                      row(f"       | weighting: {weighting}")]
        if not self._iss:
            lines.append("")
        lines.append(row(f"Sieves ({len(self._sieves)}):"))
        lines += [row(f"    + {type(sv).__name__} -> {sv.nfeatures()}")
                  for sv in self._sieves] or [row()]
        return "\n".join(lines)

    def _seeds(self) -> list:
        return self._preparateurs + self._iss + self._sieves

    def copy(self) -> "FruitSlice":
        """Same seed objects in a new, unfitted slice (fruit.py:599-609: the
        reference's shallow copy does not carry ``fit_sample_size`` over)."""
        twin = FruitSlice()
        twin.add(*self._seeds())
        return twin

    def deepcopy(self) -> "FruitSlice":
        """Copies of all seeds in a new, unfitted slice (fruit.py:611-623)."""
        twin = FruitSlice()
        twin.add(*[seed.copy() for seed in self._seeds()])
        twin.fit_sample_size = self.fit_sample_size
        return twin

    def label(self, index: int,
              level: Literal["prepared", "iterated sums", "features"] = "features",
              verbose: Literal[1, 2] = 1) -> str:
        """``preparateurs | words | sieve`` of one feature (fruit.py:625-686);
        ``verbose=1`` drops the arguments of the preparateurs and the semiring /
        weighting suffix of the words."""
        def short(text, sep):
            return text.split(sep)[0] if verbose == 1 else text

        preps = [short(p.label(), "(") for p in self._preparateurs]
        if level == "prepared":
            return " -> ".join(preps[:index + 1]) or "input"
        parts = [" -> ".join(preps)] if preps else []
        per_node = int(np.sum([sv.nfeatures() for sv in self._sieves]))
        node, feat = divmod(index, per_node) if level == "features" else (index, 0)
        # mixed radix over the ISS of the slice, first ISS most significant
        # (fruits/fruit.py:440-454 iterates them nested in that order)
        words, span = [], self.niteratedsums()
        for iss in self._iss:
            span //= iss.n_iterated_sums()
            words.append(short(iss.label((node // span) % iss.n_iterated_sums()), " : "))
        if words:
            parts.append(" -> ".join(words))
        if level == "iterated sums":
            if preps and not words:
                # (no ISS in the slice: the reference cuts four characters off
                # "<preparateurs> | ", fruit.py:671-672 -- kept for fidelity)
                return (parts[0] + " | ")[:-4]
            return " | ".join(parts) or "input"
        for sv in self._sieves:
            if feat < sv.nfeatures():
                return " | ".join(parts + [sv.label(feat)])
            feat -= sv.nfeatures()
        raise IndexError("Feature index out of range")
