"""Segment sieves (reference: ``fruits/sieving/segment.py``): the common
base with cuts and quantile thresholds (:14-104) and ``MAX`` (:107-152),
``MIN`` (:155-200), ``END`` (:203-225), ``CUR`` (:228-274), ``AVG`` (:277-317),
``STD`` (:320-358)."""
__all__ = ["MAX", "MIN", "END", "CUR", "AVG", "STD"]

from abc import ABC
from collections.abc import Sequence
from typing import Literal, Optional, Union

import numpy as np
import torch

from .. import _backend as be
from ..cache import CacheType
from .abstract import FeatureSieve, quantile_rows


class SegmentSieve(FeatureSieve, ABC):
    """Args:
        cut: index (``X[:cut]``), float in [0, 1] (coquantile of the raw
            input) or a sequence of those; default ``-1`` = whole series.
        q: thresholds as probabilities; ``1.0`` means +inf, ``-1.0`` -inf and
            ``0.0`` the value 0, anything else is a quantile fitted on the
            data.  Features count / select values in ``(q_k, q_{k+1}]``.
    """

    _kind = -1
    _inc = 0

    def __init__(self, cut: Union[Sequence[float], float] = -1,
                 q: Optional[Sequence[float]] = None,
                 coquantile_norm: Literal["L1", "L2"] = "L2") -> None:
        self._cut = cut if isinstance(cut, Sequence) else (cut,)
        self._q = q if isinstance(q, Sequence) else (-1.0, 1.0)
        self._coquantile_norm = coquantile_norm

    @property
    def requires_fitting(self) -> bool:
        return any(q not in [-1, 0, 1] for q in self._q)

    # -- thresholds ---------------------------------------------------------
    def _pre_transform_device(self, X: torch.Tensor) -> torch.Tensor:
        return X

    def _set_quantiles(self, fitted: dict) -> None:
        """``fitted`` maps a probability to its fitted value."""
        qs = np.zeros(len(self._q))
        for i, q in enumerate(self._q):
            if q == 1.0:
                qs[i] = np.inf
            elif q == -1.0:
                qs[i] = -np.inf
            elif q != 0:
                qs[i] = fitted[q]
        self._quantiles = np.sort(qs)

    def _fit_device(self, X: torch.Tensor) -> None:
        # reference :66-75: np.quantile over the whole (pre-transformed) array
        fitted = {}
        need = [q for q in self._q if q not in (1.0, -1.0, 0)]
        if need:
            arr = self._pre_transform_device(X.contiguous()).reshape(1, -1)
            for q in need:
                fitted[q] = quantile_rows(arr, q)[0]
        self._set_quantiles(fitted)

    def _get_unfitted_quantiles(self) -> None:
        # reference :77-85 (not sorted there either)
        qs = np.zeros(len(self._q))
        for i, q in enumerate(self._q):
            if q == 1.0:
                qs[i] = np.inf
            elif q == -1.0:
                qs[i] = -np.inf
            elif q != 0:
                raise RuntimeError("Sieve has not been fitted properly")
        self._quantiles = qs

    # -- cuts ------------------------------------------------------------------
    def _default_cut(self) -> bool:
        return len(self._cut) == 1 and not isinstance(self._cut[0], float) \
            and self._cut[0] == -1

    def _cuts_device(self, n: int, t: int):
        """int64 ``[n, len(cut)+1]`` sorted cut indices (reference :51-64) or
        None for the default (whole series)."""
        if self._default_cut():
            return None
        cols = [torch.zeros(n, dtype=torch.float64, device=be.require_cuda())]
        for cut in self._cut:
            if isinstance(cut, float):
                c = self._cache.get_device(CacheType.COQUANTILE,
                                           str(cut) + ":" + self._coquantile_norm)
                if c.shape[0] < n:
                    raise ValueError("coquantile cache has fewer rows than the input")
                cols.append(c[:n].to(torch.float64))
            else:
                v = cut if cut >= 0 else t + cut + 1
                cols.append(torch.full((n,), float(v), dtype=torch.float64,
                                       device=be.require_cuda()))
        cuts = torch.sort(torch.stack(cols, dim=1), dim=1).values
        return cuts.to(torch.int64).contiguous()

    # -- transform ---------------------------------------------------------------
    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not self.requires_fitting:
            self._get_unfitted_quantiles()
        elif not hasattr(self, "_quantiles"):
            raise RuntimeError("Sieve has not been fitted properly")
        arr = self._pre_transform_device(X.contiguous())
        n, t = arr.shape
        out = be.empty((n, self.nfeatures()))
        self._apply(arr, out, 0)
        return out

    def _apply(self, arr: torch.Tensor, out: torch.Tensor, col0: int) -> None:
        """Sieve the pre-transformed ``arr[n, t]`` into ``out[:, col0:col0+nf]``."""
        n, t = arr.shape
        cuts = self._cuts_device(n, t)
        q = be.to_device(np.ascontiguousarray(self._quantiles, dtype=np.float64))
        be.check(be.lib().fb_segment_sieve(
            arr.data_ptr(), arr.stride(0), be.ptr(cuts), len(self._cut) + 1, q.data_ptr(),
            len(self._q), self._kind, out.data_ptr(), out.stride(0), col0, n, t,
            be.stream_ptr()))

    # -- bookkeeping -------------------------------------------------------------
    def _nfeatures(self) -> int:
        return len(self._cut) * (len(self._q) - 1)

    def _copy(self):
        # == self.__class__(self._cut, self._q) (the copy does not keep the
        # coquantile norm, as in the reference), without re-running the argument
        # checks: fit makes one copy of every sieve per iterated sum
        new = object.__new__(self.__class__)
        new._cut, new._q, new._coquantile_norm = self._cut, self._q, "L2"
        return new

    def __str__(self) -> str:
        return f"{self.__class__.__name__}({self._cut}, {self._q})"

    def _label(self, index: int) -> str:
        r, m = divmod(index, len(self._q) - 1)
        return (f"{self.__class__.__name__}"
                f"!{self._cut[r]}![{self._q[m]}, {self._q[m+1]}]")

    def _summary(self) -> str:
        string = f"{self.__class__.__name__} -> {self.nfeatures()}:"
        for x in self._cut:
            string += f"\n   > {x}"
        return string

    def _fusable_shape(self) -> bool:
        # one segment [0, cut) -- the whole series by default -- and one (lo, hi] interval
        return len(self._cut) == 1 and len(self._q) == 2

    def _cut_key(self):
        """None for the default cut, else what identifies the segment: sieves of a
        slice can share the generated kernel's per-series cut table only if equal."""
        if self._default_cut():
            return None
        c = self._cut[0]
        return (float(c), self._coquantile_norm) if isinstance(c, float) else (int(c), None)


class MAX(SegmentSieve):
    """Maximal value with ``q_k < x <= q_{k+1}`` per cut segment (0 if none)."""
    _kind = be.SIEVE_MAX

    def _fused(self):
        return ("MAX", 0) if self._fusable_shape() else None


class MIN(SegmentSieve):
    """Minimal value with ``q_k < x <= q_{k+1}`` per cut segment (0 if none)."""
    _kind = be.SIEVE_MIN

    def _fused(self):
        return ("MIN", 0) if self._fusable_shape() else None


class END(SegmentSieve):
    """Value at the end of every cut segment, ``X[:, cut-1]``."""
    _kind = be.SIEVE_END

    def _nfeatures(self) -> int:
        # like the reference: len(cut) * (len(q) - 1) with the default q
        return len(self._cut) * (len(self._q) - 1)

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        # reference :210-219: no quantiles involved
        arr = X.contiguous()
        n, t = arr.shape
        out = be.empty((n, len(self._cut)))
        self._apply(arr, out, 0)
        return out

    def _apply(self, arr, out, col0) -> None:
        n, t = arr.shape
        cuts = self._cuts_device(n, t)
        q = be.to_device(np.array([-np.inf, np.inf]))
        be.check(be.lib().fb_segment_sieve(
            arr.data_ptr(), arr.stride(0), be.ptr(cuts), len(self._cut) + 1, q.data_ptr(), 2,
            self._kind, out.data_ptr(), out.stride(0), col0, n, t, be.stream_ptr()))

    def _fused(self):
        return ("END", 0) if self._fusable_shape() else None


class CUR(SegmentSieve):
    """Curvature (reference: segment.py:228-274): sum of the squared
    second-order increments (zero-padded like ``_increments``) that lie in
    ``(q_k, q_{k+1}]``, per cut segment.  As in the reference the thresholds
    are fitted on the values themselves (``SegmentSieve._fit`` :66-75), not on
    the increments they are later compared with."""
    _kind = be.SIEVE_CUR

    def _apply(self, arr: torch.Tensor, out: torch.Tensor, col0: int) -> None:
        arr = arr.contiguous()
        inc2 = torch.empty_like(arr)
        be.check(be.lib().fb_pretransform(arr.data_ptr(), inc2.data_ptr(), arr.shape[0],
                                          arr.shape[1], 2, be.stream_ptr()))
        super()._apply(inc2, out, col0)

    def _fused(self):
        return ("CUR", 0) if self._fusable_shape() else None


class AVG(CUR):
    """Reference :277-317.  Its ``_transform`` (:303-307) calls
    ``CUR._backend``, so the reference's AVG *is* the curvature; a drop-in
    returns what the reference returns."""


class STD(CUR):
    """Reference :320-358; ``_transform`` (:346-350) calls ``CUR._backend``
    as well (see :class:`AVG`)."""
