"""Callback hooks of ``Fruit.transform`` (reference: ``fruits/callback.py``).

Callbacks receive the prepared data, every iterated sum and every block of
sieved features as host arrays.  That is incompatible with keeping the
iterated sums in registers, so a transform with callbacks runs on the
materialising (non-fused) route: every hook costs a device-to-host copy of
what it is shown.
"""
from abc import ABC

import numpy as np


class AbstractCallback(ABC):
    """Subclass and override what you need; every hook defaults to a no-op.
    The arrays are fresh host copies (float64), safe to keep."""

    def on_next_slice(self) -> None:
        """A new ``FruitSlice`` starts (before its preparateurs run)."""

    def on_preparateur(self, X: np.ndarray) -> None:
        """``X[n, d', t]``: the data after one more preparateur of the slice."""

    def on_preparation_end(self, X: np.ndarray) -> None:
        """``X[n, d', t]``: the fully prepared input of the slice's ISS."""

    def on_iterated_sum(self, X: np.ndarray) -> None:
        """``X[n, t]``: one iterated sum, in emission order."""

    def on_sieve(self, X: np.ndarray) -> None:
        """The feature columns one sieve produced for the current iterated sum."""

    def on_sieving_end(self, X: np.ndarray) -> None:
        """``X[n, nfeatures]``: all features of the slice."""
