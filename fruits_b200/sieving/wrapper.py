"""Sieve wrappers (reference: ``fruits/sieving/wrapper.py``): ``INC``
(:9-64) evaluates a sieve on the increments of its input, ``INT`` (:67-104)
on the cumulative sums.  Like the reference, the wrapped sieve is driven
through its public ``fit`` / ``transform``: it does not inherit the wrapper's
cache, so coquantile cuts of a wrapped sieve refer to the wrapped input."""
__all__ = ["INC", "INT"]

import torch

from .. import _backend as be
from ..preparation.transform import increments_device
from .abstract import FeatureSieve


class INC(FeatureSieve):
    """Args:
        sieve: the feature sieve to evaluate.
        depth: the reference recomputes the increments from the input in every
            round (wrapper.py:44-45, :50-51), so any ``depth >= 1`` means
            single increments and ``depth = 0`` none; kept as is.
        shift: lag of the increments, ``x[t] - x[t-shift]`` (zero padded).
    """

    def __init__(self, sieve: FeatureSieve, depth: int = 1, shift: int = 1) -> None:
        self._sieve = sieve
        self._shift = shift
        self._depth = depth

    @property
    def requires_fitting(self) -> bool:
        return self._sieve.requires_fitting

    def _nfeatures(self) -> int:
        return self._sieve.nfeatures()

    def _wrapped(self, X: torch.Tensor) -> torch.Tensor:
        if self._depth <= 0:
            return X
        return increments_device(X.contiguous(), int(self._shift))

    def _fit_device(self, X: torch.Tensor) -> None:
        self._sieve.fit(self._wrapped(X))

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        return self._sieve.transform(self._wrapped(X))

    def _copy(self) -> "INC":
        return INC(self._sieve.copy(), depth=self._depth, shift=self._shift)

    def _summary(self) -> str:
        return f"INC>{self._sieve.summary()}"

    def _label(self, index: int) -> str:
        return f"INC of {self._sieve._label(index)}"

    def __str__(self) -> str:
        return f"INC({str(self._sieve)}, {self._depth}, {self._shift})"


class INT(FeatureSieve):
    """Evaluates ``sieve`` on the cumulative sums of its input
    (``np.cumsum(X, axis=1)``: sequential additions)."""

    def __init__(self, sieve: FeatureSieve) -> None:
        self._sieve = sieve

    @property
    def requires_fitting(self) -> bool:
        return self._sieve.requires_fitting

    def _nfeatures(self) -> int:
        return self._sieve.nfeatures()

    @staticmethod
    def _wrapped(X: torch.Tensor) -> torch.Tensor:
        X = X.contiguous()
        out = torch.empty_like(X)
        be.check(be.lib().fb_pretransform(X.data_ptr(), out.data_ptr(), X.shape[0], X.shape[1],
                                          -1, be.stream_ptr()))
        return out

    def _fit_device(self, X: torch.Tensor) -> None:
        self._sieve.fit(self._wrapped(X))

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        return self._sieve.transform(self._wrapped(X))

    def _copy(self) -> "INT":
        return INT(self._sieve.copy())

    def _summary(self) -> str:
        return f"INT>{self._sieve.summary()}"

    def _label(self, index: int) -> str:
        return f"INT of {self._sieve._label(index)}"

    def __str__(self) -> str:
        return f"INT({str(self._sieve)})"
