"""Preparateurs that transform every dimension (reference:
``fruits/preparation/transform.py``).  On the hot path: ``INC`` (:15-89) and
``STD`` (:92-158); ``NRM`` (:161-209) is used by the weighting lookups.  The
remaining preparateurs of the reference (MAV, LAG, FFN, RIN, RDW, JLD, SPE,
RPE, CTS, QTC, FUN) are outside the accelerated path and raise.
"""
__all__ = ["INC", "STD", "NRM", "MAV", "LAG", "FFN", "RIN", "RDW", "JLD",
           "SPE", "RPE", "CTS", "QTC", "FUN"]

from typing import Any, Callable, Union

import numpy as np
import torch

from .. import _backend as be
from .abstract import Preparateur


def increments_device(X: torch.Tensor, k: int, pad_src=None) -> torch.Tensor:
    """``_increments(X, k)`` of the reference (cache.py:8-13) on the GPU."""
    X = X.contiguous()
    out = torch.empty_like(X)
    rows = X.numel() // X.shape[-1] if X.numel() else 0
    be.check(be.lib().fb_increments(X.data_ptr(), be.ptr(pad_src), out.data_ptr(), rows,
                                    X.shape[-1], int(k), be.stream_ptr()))
    return out


class INC(Preparateur):
    """Increments ``[0, x_2-x_1, ..., x_n-x_{n-1}]`` (reference:
    transform.py:15-89; same arguments)."""

    def __init__(self, shift: Union[int, float, Callable[[int], int]] = 1,
                 depth: int = 1, zero_padding: bool = True) -> None:
        self._shift = shift
        if depth < 1:
            raise ValueError("depth has to be a positive integer > 0")
        self._depth = depth
        self._zero_padding = zero_padding

    @property
    def requires_fitting(self) -> bool:
        return False

    def _resolve_shift(self, length: int) -> int:
        if isinstance(self._shift, int):
            return self._shift
        if isinstance(self._shift, float):
            return int(np.ceil(self._shift * length))
        if callable(self._shift):
            return int(self._shift(length))
        raise TypeError(f"Type {type(self._shift)} not supported for argument shift")

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        shift = self._resolve_shift(X.shape[2])
        X = X.contiguous()
        out = X
        for _ in range(self._depth):
            out = increments_device(out, shift, None if self._zero_padding else X)
        return out

    def _fusable(self):
        if (isinstance(self._shift, int) and self._shift == 1 and self._depth == 1
                and self._zero_padding):
            return "inc"
        return None

    def _copy(self) -> "INC":
        return INC(self._shift, self._depth, self._zero_padding)

    def __eq__(self, other) -> bool:
        return (isinstance(other, INC) and self._shift == other._shift
                and self._depth == other._depth
                and self._zero_padding == other._zero_padding)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"INC({self._shift}, {self._depth}, {self._zero_padding})"


class STD(Preparateur):
    """Standardisation (reference: transform.py:92-158).  ``separately=True``
    standardises every series and dimension on its own with numpy's pairwise
    mean / std; ``separately=False`` uses one global mean / std from fit."""

    def __init__(self, separately: bool = True, var: bool = True,
                 std_eps: float = 1e-5) -> None:
        self._separately = separately
        self._div_std = var
        self._mean = None
        self._std = None
        self._eps = std_eps

    def _fit_device(self, X: torch.Tensor) -> None:
        if not self._separately:
            # one global pairwise mean / std over the flattened array
            flat = X.contiguous().reshape(1, -1)
            stats = self._row_stats(flat, self._div_std, 0.0).cpu().numpy()
            self._mean = float(stats[0, 0])
            self._std = float(stats[0, 1]) if self._div_std else 1

    @staticmethod
    def _row_stats(rows2d: torch.Tensor, div_std: bool, eps: float) -> torch.Tensor:
        rows, t = rows2d.shape
        if t >= 65536 * 1024:
            raise NotImplementedError("row too long")
        stats = be.empty((rows, 2))
        be.check(be.lib().fb_row_stats(rows2d.data_ptr(), stats.data_ptr(), rows, t,
                                       int(div_std), float(eps), be.stream_ptr()))
        return stats

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        X = X.contiguous()
        n, d, t = X.shape
        if not self._separately:
            if self._mean is None or self._std is None:
                raise RuntimeError("Missing call of self.fit()")
            stats = torch.tensor([[self._mean, self._std + self._eps]],
                                 dtype=torch.float64, device=X.device).repeat(n * d, 1)
        else:
            stats = self._row_stats(X.reshape(n * d, t), self._div_std, self._eps)
        out = torch.empty_like(X)
        be.check(be.lib().fb_standardize(X.data_ptr(), stats.data_ptr(), out.data_ptr(),
                                         n * d, t, be.stream_ptr()))
        return out

    def _fusable(self):
        return "std" if self._separately else None

    def _row_independent_fit(self) -> bool:
        return bool(self._separately)      # else: one mean / std over the whole sample

    def _copy(self) -> "STD":
        # like the reference (transform.py:146-147) the copy drops std_eps
        return STD(self._separately, self._div_std)

    def __eq__(self, other: Any) -> bool:
        return (isinstance(other, STD) and self._separately == other._separately
                and self._div_std == other._div_std)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"STD({self._separately}, {self._div_std})"


class NRM(Preparateur):
    """Min-max normalisation per series and dimension (reference:
    transform.py:161-209); ``scale_dim=True`` is not accelerated."""

    def __init__(self, scale_dim: bool = False) -> None:
        if scale_dim:
            raise NotImplementedError("NRM(scale_dim=True) is outside the GPU hot path")
        self._scale_dim = scale_dim

    @property
    def requires_fitting(self) -> bool:
        return False

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        X = X.contiguous()
        n, d, t = X.shape
        out = torch.empty_like(X)
        be.check(be.lib().fb_nrm_scale(X.data_ptr(), out.data_ptr(), n * d, t, 0, 1.0,
                                       be.stream_ptr()))
        return out

    def _copy(self) -> "NRM":
        return NRM(scale_dim=self._scale_dim)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, NRM) and self._scale_dim == other._scale_dim

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"NRM({self._scale_dim})"


def _out_of_scope(name: str, ref: str):
    class _Unsupported(Preparateur):
        __doc__ = (f"{name} (reference: {ref}) is outside the accelerated hot "
                   "path (SURVEY.md section 2, row 11); there is no CPU fallback.")

        def __init__(self, *args, **kwargs) -> None:
            raise NotImplementedError(
                f"preparateur {name} is not part of the GPU hot path")

        def _transform_device(self, X):  # pragma: no cover
            raise NotImplementedError

        def _copy(self):  # pragma: no cover
            raise NotImplementedError

    _Unsupported.__name__ = _Unsupported.__qualname__ = name
    return _Unsupported


MAV = _out_of_scope("MAV", "transform.py:212-279")
LAG = _out_of_scope("LAG", "transform.py:282-343")
FFN = _out_of_scope("FFN", "transform.py:346-470")
RIN = _out_of_scope("RIN", "transform.py:473-582")
RDW = _out_of_scope("RDW", "transform.py:585-646")
JLD = _out_of_scope("JLD", "transform.py:649-709")
SPE = _out_of_scope("SPE", "transform.py:712-760")
RPE = _out_of_scope("RPE", "transform.py:763-840")
CTS = _out_of_scope("CTS", "transform.py:843-900")
QTC = _out_of_scope("QTC", "transform.py:903-980")
FUN = _out_of_scope("FUN", "transform.py:983-1048")
