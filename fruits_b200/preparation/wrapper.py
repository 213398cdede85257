"""Preparateur wrappers (reference: ``fruits/preparation/wrapper.py``): ``DIM``
(:11-50) applies a preparateur to chosen dimensions, ``NEW`` (:53-103) appends
its result as additional dimensions.  In the fused kernels ``NEW(INC)`` costs
nothing: the added dimensions are the same rows read with the increment applied
at load time (``FruitSlice._fused_dims``)."""
__all__ = ["DIM", "NEW"]

from collections.abc import Sequence
from typing import Optional, Union

import torch

from .abstract import Preparateur


class _Carrier(Preparateur):
    """Holds the wrapped preparateur and hands it the shared cache."""

    def __init__(self, preparateur: Optional[Preparateur]) -> None:
        self._preparateur = preparateur

    @property
    def requires_fitting(self) -> bool:
        return self._preparateur is not None and self._preparateur.requires_fitting

    def _inner(self) -> Optional[Preparateur]:
        if self._preparateur is not None:
            self._preparateur._cache = self._cache
        return self._preparateur

    def _row_independent_fit(self) -> bool:
        return self._preparateur is None or self._preparateur._row_independent_fit()

    def _row_independent_transform(self) -> bool:
        return self._preparateur is None or self._preparateur._row_independent_transform()

    def _needs_raw_cache(self) -> bool:
        return self._preparateur is not None and self._preparateur._needs_raw_cache()


class DIM(_Carrier):
    """The untouched dimensions first, then ``preparateur`` applied to the
    dimensions ``dim`` (0-based index or indices)."""

    def __init__(self, preparateur: Preparateur, dim: Union[int, Sequence[int]]) -> None:
        super().__init__(preparateur)
        self._dim = (int(dim),) if isinstance(dim, int) else tuple(int(d) for d in dim)

    def _pick(self, X: torch.Tensor, dims) -> torch.Tensor:
        idx = torch.as_tensor(list(dims), device=X.device, dtype=torch.long)
        return X.index_select(1, idx).contiguous()

    def _fit_device(self, X: torch.Tensor) -> None:
        self._inner()._fit_device(self._pick(X, self._dim))

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        changed = self._inner()._transform_device(self._pick(X, self._dim))
        kept = self._pick(X, [d for d in range(X.shape[1]) if d not in self._dim])
        return torch.cat((kept, changed), dim=1).contiguous()

    def _copy(self) -> "DIM":
        return DIM(self._preparateur.copy(), self._dim)

    def __str__(self) -> str:
        return f"DIM({self._preparateur}, {self._dim})"


class NEW(_Carrier):
    """All dimensions followed by ``preparateur`` applied to all of them;
    ``NEW()`` without a preparateur duplicates the dimensions."""

    def __init__(self, preparateur: Optional[Preparateur] = None) -> None:
        super().__init__(preparateur)

    def _fit_device(self, X: torch.Tensor) -> None:
        if self._inner() is not None:
            self._preparateur._fit_device(X)

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        inner = self._inner()
        added = X if inner is None else inner._transform_device(X)
        return torch.cat((X, added), dim=1).contiguous()

    def _copy(self) -> "NEW":
        return NEW(None if self._preparateur is None else self._preparateur.copy())

    def __str__(self) -> str:
        return f"NEW({self._preparateur})"
