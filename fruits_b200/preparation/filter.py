"""Filtering preparateurs (reference: ``fruits/preparation/filter.py``): they
set parts of every series to zero -- random strips (``DIL``), everything
outside a per-series window (``WIN``), all but every n-th point (``DOT``),
evenly spaced strips (``PDD``).  All four are one kernel, ``fb_time_mask``: a
keep mask over the time axis and / or a window per series.  ``fit`` draws from
the global numpy RNG exactly like the reference.
"""
__all__ = ["DIL", "WIN", "DOT", "PDD"]

from typing import Any, Optional, Union

import numpy as np
import torch

from .. import _backend as be
from ..cache import CacheType
from .abstract import Preparateur


def _keep_mask(owner, X: torch.Tensor, state: tuple, build) -> torch.Tensor:
    """The uint8 keep mask of a fitted preparateur on the device, rebuilt only when the
    series length, the device or the fitted state changes (``build(t)`` is a host loop
    over up to ``t`` strips)."""
    key = (X.shape[2], X.device, state)
    memo = owner.__dict__.get("_keep_memo")
    if memo is None or memo[0] != key:
        mask = torch.from_numpy(np.ascontiguousarray(build(X.shape[2]), dtype=np.uint8))
        memo = owner._keep_memo = (key, mask.to(X.device))
    return memo[1]


def _masked(X: torch.Tensor, keep_d=None, lo=None, hi=None, lo_off: int = 0) -> torch.Tensor:
    X = X.contiguous()
    n, d, t = X.shape
    out = torch.empty_like(X)
    be.check(be.lib().fb_time_mask(X.data_ptr(), out.data_ptr(), n, d, t, be.ptr(keep_d),
                                   be.ptr(lo), be.ptr(hi), lo_off, be.stream_ptr()))
    return out


class DIL(Preparateur):
    """Dilation: random strips of every series are set to zero (reference:
    filter.py:11-70)."""

    def __init__(self, clusters: Optional[float] = None) -> None:
        self._clusters = clusters

    def _fit_device(self, X: torch.Tensor) -> None:
        t = X.shape[2]
        if self._clusters is not None:
            nclusters = int(self._clusters * t)
        else:
            upper_bound = int(np.floor(t / 10.0))
            nclusters = 1 if upper_bound <= 1 else np.random.randint(1, upper_bound)
        if nclusters >= t:
            self._indices = np.arange(t)
        else:
            self._indices = np.sort(np.random.choice(t, size=nclusters, replace=False))
        self._lengths = []
        for i in range(nclusters):       # one draw per strip, like filter.py:48-53
            if i == nclusters - 1:
                max_length = t - self._indices[i]
            else:
                max_length = self._indices[i + 1] - self._indices[i]
            self._lengths.append(np.random.randint(1, max_length + 1))

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_indices") or not hasattr(self, "_lengths"):
            raise RuntimeError("Missing call of self.fit()")
        def build(t):
            keep = np.ones(t, dtype=np.uint8)
            for i, index in enumerate(self._indices):
                keep[index:index + self._lengths[i]] = 0
            return keep
        state = (np.asarray(self._indices).tobytes(), tuple(int(x) for x in self._lengths))
        return _masked(X, _keep_mask(self, X, state, build))

    def _copy(self) -> "DIL":
        return DIL(self._clusters)

    def __str__(self) -> str:
        return f"DIL(clusters={self._clusters})"


class WIN(Preparateur):
    """Window: everything outside ``[start, end]`` -- measured as quantiles of
    the quadratic variation of the raw input -- is set to zero (reference:
    filter.py:73-121)."""

    def __init__(self, start: float, end: float) -> None:
        self._start = start
        self._end = end

    @property
    def requires_fitting(self) -> bool:
        return False

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        lo = self._cache.get_device(CacheType.COQUANTILE, str(self._start) + ":L2")
        hi = self._cache.get_device(CacheType.COQUANTILE, str(self._end) + ":L2")
        # the cache belongs to the RAW batch: row i of it serves row i of X, whatever
        # rows X holds (a fit sample included -- fruits/cache.py:97-112, filter.py:107-112)
        if lo.shape[0] < X.shape[0]:
            raise IndexError(f"index {lo.shape[0]} is out of bounds for the cached coquantiles")
        # X[i, j, coq_start[i]-1 : coq_end[i]] with Python's slice rules (a start of -1
        # is the last time step)
        return _masked(X, None, lo, hi, -1)

    def _needs_raw_cache(self) -> bool:
        return True

    def _copy(self) -> "WIN":
        return WIN(self._start, self._end)

    def __eq__(self, other) -> bool:
        if not isinstance(other, WIN):
            raise TypeError(f"Cannot compare WIN with type {type(other)}")
        return self._start == other._start and self._end == other._end

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"WIN(start={self._start}, end={self._end})"


class DOT(Preparateur):
    """Dotting: keeps every ``n``-th point from ``first`` on, zero elsewhere
    (reference: filter.py:124-200)."""

    def __init__(self, n: Union[int, float] = 2,
                 first: Optional[Union[int, float]] = None) -> None:
        if isinstance(n, float) and not 0 < n < 1:
            raise ValueError("If n is a float, it has to satisfy 0 < n < 1")
        if not isinstance(n, float) and not isinstance(n, int):
            raise TypeError("n has to be either a float or integer")
        self._n_given = n
        if isinstance(first, float) and not 0 < first < 1:
            raise ValueError("If first is a float,  it has to satisfy 0 < first < 1")
        if not isinstance(first, (float, int)) and first is not None:
            raise TypeError("first has to be either a float, integer or None")
        self._first_given = first

    def _fit_device(self, X: torch.Tensor) -> None:
        t = X.shape[2]
        if isinstance(self._n_given, float):
            self._n = max(int(self._n_given * t), 1)
        else:
            self._n = min(self._n_given, t)
        if isinstance(self._first_given, float):
            self._first = min(max(int(self._first_given * t), 1), t - 1)
        elif self._first_given is not None:
            self._first = min(self._first_given, t - 1)
        else:
            self._first = self._n - 1

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_n") or not hasattr(self, "_first"):
            raise RuntimeError("Missing call of self.fit()")
        def build(t):
            keep = np.zeros(t, dtype=np.uint8)
            keep[self._first::self._n] = 1
            return keep
        return _masked(X, _keep_mask(self, X, (self._first, self._n), build))

    def _copy(self) -> "DOT":
        return DOT(self._n_given, self._first_given)

    def __eq__(self, other: Any) -> bool:
        if not isinstance(other, DOT):
            raise TypeError(f"Cannot compare DOT with type {type(other)}")
        return self._n_given == other._n_given and self._first_given == other._first_given

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"DOT(n={self._n_given}, first={self._first_given})"


class PDD(Preparateur):
    """Proportion-density drop: evenly spaced strips are set to zero
    (reference: filter.py:203-270)."""

    def __init__(self, density: float = 0.1, proportion: float = 0.5) -> None:
        if not isinstance(density, float) or not 0.0 < density <= 1.0:
            raise ValueError("density has to be a float 0 < density <= 1")
        if not isinstance(proportion, float) or not 0.0 < proportion < 1.0:
            raise ValueError("proportion has to be a float 0 < proportion < 1")
        self._d_given = density
        self._p_given = proportion

    def _fit_device(self, X: torch.Tensor) -> None:
        t = X.shape[2]
        p = max(int(self._p_given * t), 1)
        points = max(int((1.0 - self._d_given) * t), 1)
        self._width = int(p / points)
        if points == t - self._width:
            points -= 1
        self._indices = np.linspace(0, t - self._width, points, dtype="int")

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_width") or not hasattr(self, "_indices"):
            raise RuntimeError("Missing call of self.fit()")
        def build(t):
            keep = np.ones(t, dtype=np.uint8)
            for index in self._indices:
                keep[index:index + self._width] = 0
            return keep
        state = (np.asarray(self._indices).tobytes(), int(self._width))
        return _masked(X, _keep_mask(self, X, state, build))

    def _copy(self) -> "PDD":
        return PDD(self._d_given, self._p_given)

    def __eq__(self, other) -> bool:
        if not isinstance(other, PDD):
            raise TypeError(f"Cannot compare PDD with type {type(other)}")
        return self._d_given == other._d_given and self._p_given == other._p_given

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"PDD(density={self._d_given}, proportion={self._p_given})"
