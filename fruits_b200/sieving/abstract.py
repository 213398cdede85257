"""Reference: ``fruits/sieving/abstract.py:8-34``."""
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import _backend as be
from ..seed import Seed


class FeatureSieve(Seed, ABC):
    """A sieve maps an iterated sum ``[n_series, length]`` to a few numbers
    per series."""

    @abstractmethod
    def _nfeatures(self) -> int:
        ...

    def nfeatures(self) -> int:
        return self._nfeatures()

    @abstractmethod
    def _summary(self) -> str:
        ...

    def summary(self) -> str:
        return self._summary()

    def _fused(self):
        """Description of this sieve for the fused kernel or None if it can
        only run on materialised iterated sums."""
        return None


def order_stats(V: torch.Tensor, k: int):
    """x_(k), x_(k+1) of every row of the contiguous ``V[P, M]`` on the GPU."""
    V = V.contiguous()
    P, M = V.shape
    lo, hi = be.empty((P,)), be.empty((P,))
    work = be.empty((be.lib().fb_order_stats_workspace(P),), dtype=torch.uint8)
    be.check(be.lib().fb_order_stats(V.data_ptr(), M, P, M, int(k), lo.data_ptr(),
                                     hi.data_ptr(), work.data_ptr(), be.stream_ptr()))
    return lo, hi


def _virtual_index(M: int, q: float):
    """numpy's "linear" quantile method: x_(k) + gamma * (x_(k+1) - x_(k))."""
    virtual = (M - 1) * float(q)
    if virtual >= M - 1:
        return M - 1, 0.0
    if virtual < 0:
        return 0, 0.0
    k = int(np.floor(virtual))
    return k, virtual - np.floor(virtual)


def _lerp(a: np.ndarray, b: np.ndarray, gamma: float) -> np.ndarray:
    """numpy's ``_lerp`` (numpy/lib/_function_base_impl.py) in float64."""
    gamma = np.float64(gamma)
    with np.errstate(invalid="ignore"):
        diff = b - a
        res = a + diff * gamma
        if gamma >= 0.5:
            res = b - diff * (1 - gamma)
    return res


def quantile_rows(V: torch.Tensor, q: float) -> np.ndarray:
    """``np.quantile(row, q)`` (method "linear") for every row of ``V[P, M]``:
    the two order statistics come from the GPU radix select, the interpolation
    is numpy's ``_lerp`` in float64."""
    P, M = V.shape
    k, gamma = _virtual_index(M, q)
    lo, hi = order_stats(V, k)
    a = lo.cpu().numpy()
    b = hi.cpu().numpy() if k < M - 1 else a
    return _lerp(a, b, gamma)


def quantile_multi(V: torch.Tensor, t: int, pairs: list) -> dict:
    """``{(inc, q): np.quantile(rows, q)}`` for every row of ``V[P, n*t]`` and
    every ``(inc, q)`` in ``pairs``: the quantile of the ``inc``-fold
    zero-padded increments (``inc`` in 0..2) of the ``t``-long series the row
    is made of.  Up to four selections share three reads of ``V``
    (``fb_order_stats_multi``); a value of ``None`` means "too many equal
    values for the candidate list", the caller then selects on the
    materialised increments."""
    import ctypes
    V = V.contiguous()
    P, M = V.shape
    out = {}
    L = be.lib()
    for g in range(0, len(pairs), 4):
        group = pairs[g:g + 4]
        S = len(group)
        kg = [_virtual_index(M, q) for _, q in group]
        incs = (ctypes.c_int32 * S)(*[int(i) for i, _ in group])
        ks = (ctypes.c_int64 * S)(*[int(k) for k, _ in kg])
        lo, hi = be.empty((P, S)), be.empty((P, S))
        done = be.empty((P, S), dtype=torch.int32)
        work = be.empty((L.fb_order_stats_multi_workspace(P, S),), dtype=torch.uint8)
        be.check(L.fb_order_stats_multi(V.data_ptr(), M, P, M, int(t), S, incs, ks, lo.data_ptr(),
                                        hi.data_ptr(), done.data_ptr(), work.data_ptr(),
                                        be.stream_ptr()))
        packed = torch.cat([lo, hi, done.to(torch.float64)], dim=1).cpu().numpy()
        for s, (pair, (k, gamma)) in enumerate(zip(group, kg)):
            if not packed[:, 2 * S + s].all():
                out[pair] = None
                continue
            a = packed[:, s]
            b = packed[:, S + s] if k < M - 1 else a
            out[pair] = _lerp(a, b, gamma)
    return out
