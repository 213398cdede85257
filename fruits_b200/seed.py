"""Base class of everything that can be added to a fruit.

API of the reference's ``fruits/seed.py`` (:11-82).  Public ``fit`` /
``transform`` accept numpy arrays (copied to the GPU, result copied back) or
CUDA ``torch`` tensors (used in place, result stays on the GPU); subclasses
implement ``_fit_device`` / ``_transform_device`` on device tensors.
"""
from abc import ABC, abstractmethod
from typing import TypeVar

import torch

from . import _backend as be
from .cache import SharedSeedCache

TCopy = TypeVar("TCopy", bound="Seed")


def _raw3d(X: torch.Tensor) -> torch.Tensor:
    return X[:, None, :] if X.dim() == 2 else X


class Seed(ABC):

    _cache: SharedSeedCache

    @property
    def requires_fitting(self) -> bool:
        return True

    # -- device-side hooks -------------------------------------------------
    def _fit_device(self, X: torch.Tensor) -> None:
        pass

    @abstractmethod
    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        ...

    # -- reference-style private names kept as aliases ----------------------
    def _fit(self, X) -> None:
        self._fit_device(be.to_device(X))

    def _transform(self, X):
        return self._transform_device(be.to_device(X))

    # -- public API ----------------------------------------------------------
    def fit(self, X) -> None:
        """Fits the seed to the given data (reference: seed.py:26-35)."""
        Xd = be.to_device(X)
        has_cache = hasattr(self, "_cache")
        if not has_cache:
            self._cache = SharedSeedCache(_raw3d(Xd))
        try:
            self._fit_device(Xd)
        finally:
            if not has_cache:
                del self._cache

    def transform(self, X):
        """Transforms the given data (reference: seed.py:41-51)."""
        Xd = be.to_device(X)
        has_cache = hasattr(self, "_cache")
        if not has_cache:
            self._cache = SharedSeedCache(_raw3d(Xd))
        try:
            result = self._transform_device(Xd)
        finally:
            if not has_cache:
                del self._cache
        if isinstance(X, torch.Tensor):
            return result
        return result.cpu().numpy()

    def fit_transform(self, X):
        self.fit(X)
        return self.transform(X)

    @abstractmethod
    def _copy(self: TCopy) -> TCopy:
        ...

    def copy(self: TCopy) -> TCopy:
        return self._copy()

    def _label(self, index: int = 0) -> str:
        return str(self)

    def label(self, index: int = 0) -> str:
        return self._label(index)

    def __str__(self) -> str:
        return self.__class__.__name__
