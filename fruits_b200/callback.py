"""Callback hooks of ``Fruit.transform`` (reference: ``fruits/callback.py``).

Callbacks receive the prepared data, every iterated sum and every block of
sieved features as host arrays.  That is incompatible with keeping the
iterated sums in registers, so a transform with callbacks runs on the
materialising (non-fused) route.
"""
from abc import ABC

import numpy as np


class AbstractCallback(ABC):

    def on_next_slice(self) -> None:
        """Called every time the next FruitSlice starts."""

    def on_preparateur(self, X: np.ndarray) -> None:
        """Called after each preparateur with the prepared data."""

    def on_preparation_end(self, X: np.ndarray) -> None:
        """Called once after the last preparateur."""

    def on_iterated_sum(self, X: np.ndarray) -> None:
        """Called for every iterated sum."""

    def on_sieve(self, X: np.ndarray) -> None:
        """Called after each use of a feature sieve."""

    def on_sieving_end(self, X: np.ndarray) -> None:
        """Called once at the end of the feature calculation."""
