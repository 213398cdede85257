"""Preparateur wrappers (reference: ``fruits/preparation/wrapper.py``)."""
__all__ = ["DIM", "NEW"]

from collections.abc import Sequence
from typing import Optional, Union

import numpy as np
import torch

from .abstract import Preparateur


class DIM(Preparateur):
    """Applies a preparateur to some dimensions only and appends the result
    behind the untouched dimensions (reference: wrapper.py:11-50)."""

    def __init__(self, preparateur: Preparateur, dim: Union[int, Sequence[int]]) -> None:
        self._preparateur = preparateur
        self._dim = np.array([dim]) if isinstance(dim, int) else np.array(dim)

    @property
    def requires_fitting(self) -> bool:
        return self._preparateur.requires_fitting

    def _sel(self, X):
        return torch.as_tensor(self._dim, device=X.device, dtype=torch.long)

    def _fit_device(self, X: torch.Tensor) -> None:
        self._preparateur._cache = self._cache
        self._preparateur._fit_device(X.index_select(1, self._sel(X)).contiguous())

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        self._preparateur._cache = self._cache
        picked = X.index_select(1, self._sel(X)).contiguous()
        transformed = self._preparateur._transform_device(picked)
        keep = [d for d in range(X.shape[1]) if d not in set(int(x) for x in self._dim)]
        rest = X.index_select(1, torch.as_tensor(keep, device=X.device, dtype=torch.long))
        return torch.cat((rest, transformed), dim=1).contiguous()

    def _copy(self) -> "DIM":
        return DIM(self._preparateur.copy(), tuple(int(x) for x in self._dim))

    def __str__(self) -> str:
        return f"DIM({str(self._preparateur)}, {tuple(int(x) for x in self._dim)})"


class NEW(Preparateur):
    """Appends the result of a preparateur as new dimensions (reference:
    wrapper.py:53-103); ``NEW()`` duplicates all dimensions."""

    def __init__(self, preparateur: Optional[Preparateur] = None) -> None:
        self._preparateur = preparateur

    @property
    def requires_fitting(self) -> bool:
        if self._preparateur is None:
            return False
        return self._preparateur.requires_fitting

    def _fit_device(self, X: torch.Tensor) -> None:
        if self._preparateur is not None:
            self._preparateur._cache = self._cache
            self._preparateur._fit_device(X)

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if self._preparateur is None:
            return torch.cat((X, X), dim=1).contiguous()
        self._preparateur._cache = self._cache
        transformed = self._preparateur._transform_device(X)
        return torch.cat((X, transformed), dim=1).contiguous()

    def _copy(self) -> "NEW":
        if self._preparateur is None:
            return NEW()
        return NEW(self._preparateur.copy())

    def __str__(self) -> str:
        return f"NEW({str(self._preparateur)})"
