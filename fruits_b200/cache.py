"""Raw-input cache shared by all seeds of one ``Fruit.fit`` / ``transform``
call (reference: ``fruits/cache.py:51-135``).

Holds the raw input on the GPU and memoises the two quantities derived from
it: the L1/L2 increment sums of dimension 0 (exponential weightings) and the
coquantile cut indices (float ``cut`` arguments of the sieves).
"""
from enum import Enum, auto
from typing import Optional

import numpy as np
import torch

from . import _backend as be


class CacheType(Enum):
    COQUANTILE = auto()
    ISS = auto()


class SharedSeedCache:

    def __init__(self, X=None) -> None:
        self._cache = {CacheType.COQUANTILE: {}, CacheType.ISS: {}}
        # uploaded on first use: most pipelines never ask the cache for anything,
        # and fit only needs the fit sample on the device
        self._source = X
        self._uploaded = None

    @property
    def _input(self):
        if self._uploaded is None and self._source is not None:
            self._uploaded = self._as3d(be.to_device(self._source))
        return self._uploaded

    @staticmethod
    def _as3d(X: torch.Tensor) -> torch.Tensor:
        if X.dim() == 1:
            return X[None, None, :].contiguous()
        if X.dim() == 2:
            return X[:, None, :].contiguous()
        return X

    @staticmethod
    def _lsum(X: torch.Tensor, l2: bool) -> torch.Tensor:
        n, d, t = X.shape
        out = be.empty((n, t))
        be.check(be.lib().fb_lsum(X.data_ptr(), out.data_ptr(), n, d, t, int(l2),
                                  be.stream_ptr()))
        return out

    def get_device(self, cache_id: CacheType, key: str, X=None) -> torch.Tensor:
        store = self._cache[cache_id]
        if store.get(key) is None:
            src = self._input
            if src is None:
                if X is None:
                    raise RuntimeError("No input for cache given")
                src = self._as3d(be.to_device(X))
            if cache_id == CacheType.COQUANTILE:
                c, norm = key.split(":")
                if norm not in ("L1", "L2"):
                    raise ValueError(f"unknown coquantile norm {norm!r}")
                s = self._lsum(src, norm == "L2")
                out = be.empty((s.shape[0],), dtype=torch.int64)
                be.check(be.lib().fb_coquantile(s.data_ptr(), out.data_ptr(), s.shape[0],
                                                s.shape[1], float(c), be.stream_ptr()))
                store[key] = out
            else:
                if key not in ("L1", "L2"):
                    raise ValueError(f"unknown cache key {key!r}")
                store[key] = self._lsum(src, key == "L2")
        return store[key]

    def get(self, cache_id: CacheType, key: str, X: Optional[np.ndarray] = None) -> np.ndarray:
        """Same contract as the reference's ``SharedSeedCache.get``; returns a
        host copy."""
        return self.get_device(cache_id, key, X).cpu().numpy()
