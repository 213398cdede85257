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


class _Wrapped(FeatureSieve):
    """Everything but the transformation of the input is the wrapped sieve's."""

    _tag = ""

    def __init__(self, sieve: FeatureSieve) -> None:
        self._sieve = sieve

    def _wrapped(self, X: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    requires_fitting = property(lambda self: self._sieve.requires_fitting)

    def _nfeatures(self) -> int:
        return self._sieve.nfeatures()

    def _fit_device(self, X: torch.Tensor) -> None:
        self._sieve.fit(self._wrapped(X.contiguous()))

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        return self._sieve.transform(self._wrapped(X.contiguous()))

    def _summary(self) -> str:
        return f"{self._tag}>{self._sieve.summary()}"

    def _label(self, index: int) -> str:
        return f"{self._tag} of {self._sieve._label(index)}"


class INC(_Wrapped):
    """``sieve`` on the increments ``x[t] - x[t-shift]`` (zero padded).  The
    reference recomputes the increments from the input in every round
    (wrapper.py:44-45, :50-51), so any ``depth >= 1`` means single increments
    and ``depth = 0`` none; kept as is."""

    _tag = "INC"

    def __init__(self, sieve: FeatureSieve, depth: int = 1, shift: int = 1) -> None:
        super().__init__(sieve)
        self._depth, self._shift = depth, shift

    def _wrapped(self, X: torch.Tensor) -> torch.Tensor:
        return increments_device(X, int(self._shift)) if self._depth > 0 else X

    def _copy(self) -> "INC":
        return INC(self._sieve.copy(), depth=self._depth, shift=self._shift)

    def __str__(self) -> str:
        return f"INC({self._sieve}, {self._depth}, {self._shift})"


class INT(_Wrapped):
    """``sieve`` on the cumulative sums of its input (``np.cumsum(X, axis=1)``:
    sequential additions, ``fb_pretransform`` with a negative depth)."""

    _tag = "INT"

    def _wrapped(self, X: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(X)
        be.check(be.lib().fb_pretransform(X.data_ptr(), out.data_ptr(), X.shape[0], X.shape[1],
                                          -1, be.stream_ptr()))
        return out

    def _copy(self) -> "INT":
        return INT(self._sieve.copy())

    def __str__(self) -> str:
        return f"INT({self._sieve})"
