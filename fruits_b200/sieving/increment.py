"""Increment sieves (reference: ``fruits/sieving/increment.py``): sieves that
look at the ``inc``-fold increments of an iterated sum -- ``NPI`` (:101-129),
``MPI`` (:132-163), ``XPI`` (:166-199), ``LPI`` (:202-239)."""
__all__ = ["NPI", "MPI", "XPI", "LPI"]

from collections.abc import Sequence
from typing import Literal, Optional, Union

import torch

from .. import _backend as be
from .segment import SegmentSieve


class IncrementSieve(SegmentSieve):
    """Args as :class:`SegmentSieve` (default ``q=(0, 1)``) plus
    ``inc``: ``inc > 0`` sieves the ``inc``-fold increments (the leading zeros
    count as values), ``inc < 0`` the ``-inc``-fold cumulative sums."""

    def __init__(self, cut: Union[Sequence[float], float] = -1,
                 q: Optional[Sequence[float]] = None, inc: int = 1,
                 coquantile_norm: Literal["L1", "L2"] = "L2") -> None:
        super().__init__(cut, q if q is not None else (0.0, 1.0), coquantile_norm)
        self._inc = inc

    def _pre_transform_device(self, X: torch.Tensor) -> torch.Tensor:
        # reference :63-71
        if self._inc == 0:
            return X
        X = X.contiguous()
        out = torch.empty_like(X)
        be.check(be.lib().fb_pretransform(X.data_ptr(), out.data_ptr(), X.shape[0],
                                          X.shape[1], int(self._inc), be.stream_ptr()))
        return out

    def _copy(self):
        # like the reference (:83-84) the copy drops coquantile_norm
        new = super()._copy()
        new._inc = self._inc
        return new

    def __str__(self) -> str:
        return f"{self.__class__.__name__}({self._cut}, {self._q}, {self._inc})"

    def _label(self, index: int) -> str:
        label = super()._label(index)
        return label[:3] + f"[inc={self._inc}]" + label[3:]


class NPI(IncrementSieve):
    """Number of increments in ``(q_k, q_{k+1}]`` (not normalised)."""
    _kind = be.SIEVE_NPI

    def _fused(self):
        ok = self._fusable_shape() and 0 <= self._inc <= 2
        return ("CNT", self._inc) if ok else None


class MPI(IncrementSieve):
    """Mean of the increments in ``(q_k, q_{k+1}]`` (0 if there are none)."""
    _kind = be.SIEVE_MPI

    def _fused(self):
        ok = self._fusable_shape() and 0 <= self._inc <= 2
        return ("AVG", self._inc) if ok else None


class XPI(IncrementSieve):
    """Mean index of the increments in ``(q_k, q_{k+1}]``."""
    _kind = be.SIEVE_XPI

    def _fused(self):
        ok = self._fusable_shape() and 0 <= self._inc <= 2
        return ("XPI", self._inc) if ok else None


class LPI(IncrementSieve):
    """Longest run of consecutive increments in ``(q_k, q_{k+1}]``."""
    _kind = be.SIEVE_LPI

    def _fused(self):
        ok = self._fusable_shape() and 0 <= self._inc <= 2
        return ("LPI", self._inc) if ok else None
