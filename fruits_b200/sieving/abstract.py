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


def quantile_rows(V: torch.Tensor, q: float) -> np.ndarray:
    """``np.quantile(row, q)`` (method "linear") for every row of ``V[P, M]``:
    the two order statistics come from the GPU radix select, the interpolation
    is numpy's ``_lerp`` (numpy/lib/_function_base_impl.py) in float64."""
    P, M = V.shape
    virtual = (M - 1) * float(q)
    if virtual >= M - 1:
        k, gamma = M - 1, 0.0
    elif virtual < 0:
        k, gamma = 0, 0.0
    else:
        k = int(np.floor(virtual))
        gamma = virtual - np.floor(virtual)
    lo, hi = order_stats(V, k)
    a = lo.cpu().numpy()
    b = hi.cpu().numpy() if k < M - 1 else a
    gamma = np.float64(gamma)
    with np.errstate(invalid="ignore"):
        diff = b - a
        res = a + diff * gamma
        if gamma >= 0.5:
            res = b - diff * (1 - gamma)
    return res
