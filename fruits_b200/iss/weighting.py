"""Exponential weightings of the iterated sums (reference:
``fruits/iss/weighting.py``).

A weighting provides the lookup ``g[n_series, length]``; summands that
combine time steps ``i < j`` are scaled by ``exp(alpha * (g(i) - g(j)))``
inside the ISS kernel (``csrc/lns.cuh``, weighted modes).  ``get_lookup_device``
returns ``(g, shared)``: ``shared=True`` means one row valid for all series
(Indices, Plateaus), so the kernel reads a single ``g[length]`` row.
"""
__all__ = ["Weighting", "L1", "L2", "Indices", "Plateaus", "Custom"]

from abc import ABC, abstractmethod
from typing import Callable, Optional

import numpy as np
import torch

from .. import _backend as be
from ..cache import CacheType, SharedSeedCache


def _nrm_scale_host(r: np.ndarray, scale: float) -> np.ndarray:
    """NRM of one row followed by ``* scale`` with numpy's operation order
    (reference: preparation/transform.py:184-198, weighting.py:107-110)."""
    lo, hi = np.min(r), np.max(r)
    out = np.zeros_like(r)   # keeps an integer dtype, exactly like the reference
    if lo != hi:
        out[...] = (r - lo) / (hi - lo)
    return np.asarray(out * scale, dtype=np.float64)


class Weighting(ABC):
    """Reference: weighting.py:12-39."""

    _cache: SharedSeedCache

    def __init__(self, total: bool = False) -> None:
        self.total = total

    @abstractmethod
    def get_lookup_device(self, X: torch.Tensor):
        """-> (g tensor, shared flag)"""

    def _memoized_row(self, X: torch.Tensor, build):
        """The shared row of a lookup that depends on the length only: built (and copied
        to the device) once per (length, device, parameters), not on every transform."""
        if getattr(self, "_transform", None) is not None:
            return build(X)                      # a Python transform may carry state
        key = (X.shape[2], X.device, tuple(sorted(
            (k, v) for k, v in self.__dict__.items() if k not in ("_row_memo", "_cache"))))
        memo = self.__dict__.get("_row_memo")
        if memo is None or memo[0] != key:
            self._row_memo = memo = (key, build(X))
        return memo[1]

    def get_lookup(self, X) -> np.ndarray:
        """``g[n, length]`` as a host array (reference API)."""
        Xd = be.to_device(X)
        g, shared = self.get_lookup_device(Xd)
        g = g.cpu().numpy()
        if shared:
            return np.ones((Xd.shape[0], Xd.shape[2])) * g
        return g


class Custom(Weighting):
    """``g = transform(X)`` evaluated on the host (reference :42-66)."""

    def __init__(self, transform: Callable[[np.ndarray], np.ndarray],
                 total: bool = False) -> None:
        super().__init__(total=total)
        self._transform = transform

    def get_lookup_device(self, X: torch.Tensor):
        g = np.ascontiguousarray(self._transform(X.cpu().numpy()), dtype=np.float64)
        return be.to_device(g), False


class Indices(Weighting):
    """``g(i) = i/N`` scaled to ``[0, scale]`` (reference :69-110)."""

    def __init__(self, relative: bool = True,
                 transform: Optional[Callable[[float], float]] = None,
                 scale: float = 50, total: bool = False) -> None:
        super().__init__(total=total)
        self._transform = transform
        self._relative = relative
        self._scale = scale

    def get_lookup_device(self, X: torch.Tensor):
        return self._memoized_row(X, self._shared_row)

    def _shared_row(self, X: torch.Tensor):
        length = X.shape[2]
        r = np.arange(1, length + 1)
        if self._relative:
            r = r / length
        if self._transform is not None:
            r = np.vectorize(self._transform)(r)
        g = _nrm_scale_host(r, self._scale)
        return be.to_device(np.ascontiguousarray(g)), True


class _IncrementSum(Weighting):
    _key = "L1"

    def __init__(self, on_prepared: bool = False, relative: bool = False,
                 transform: Optional[Callable[[float], float]] = None,
                 scale: float = 50, total: bool = False) -> None:
        super().__init__(total=total)
        self._on_prepared = on_prepared
        self._relative = relative
        self._transform = transform
        self._scale = scale

    def get_lookup_device(self, X: torch.Tensor):
        if not self._on_prepared:
            # raw-input cache, dimension 0 (reference cache.py:27, :97-112)
            r = self._cache.get_device(CacheType.ISS, self._key, X)
        else:
            r = SharedSeedCache._lsum(X.contiguous(), self._key == "L2")
        relative = int(self._relative)
        if self._transform is not None:
            # the caller's scalar Python function (np.vectorize, weighting.py:155-156):
            # the sums make one round trip through the host, the normalisation stays
            # on the GPU
            host = r.cpu().numpy()
            if self._relative:
                host = host / (host[:, -1:] + 1e-5)
            r = be.to_device(np.ascontiguousarray(np.vectorize(self._transform)(host),
                                                  dtype=np.float64))
            relative = 0
        out = torch.empty_like(r)
        be.check(be.lib().fb_nrm_scale(r.data_ptr(), out.data_ptr(), r.shape[0], r.shape[1],
                                       relative, float(self._scale), be.stream_ptr()))
        return out, False


class L1(_IncrementSum):
    """``g(i)`` = sum of absolute increments up to ``i`` (reference :113-160)."""
    _key = "L1"


class L2(_IncrementSum):
    """``g(i)`` = sum of squared increments up to ``i`` (reference :163-210)."""
    _key = "L2"


class Plateaus(Weighting):
    """Step function with ``n`` plateaus (reference :213-256)."""

    def __init__(self, n: int, reverse: bool = False, scale: float = 50,
                 total: bool = False) -> None:
        super().__init__(total=total)
        if n <= 1:
            raise ValueError(f"Number of plateaus ({n}) has to be > 1")
        self._nplateaus = n
        self._reverse = reverse
        self._scale = scale

    def get_lookup_device(self, X: torch.Tensor):
        return self._memoized_row(X, self._shared_row)

    def _shared_row(self, X: torch.Tensor):
        length = X.shape[2]
        r = np.ones(length)
        step = int(length / self._nplateaus)
        for i in range(self._nplateaus):
            r[i * step:(i + 1) * step] = i / (self._nplateaus - 1)
        if self._reverse:
            r = r[::-1]
        g = _nrm_scale_host(np.ascontiguousarray(r), self._scale)
        return be.to_device(np.ascontiguousarray(g)), True
