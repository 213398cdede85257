"""Iterated-sums signature: the ``ISS`` seed, its modes, semirings, weightings,
the cosine weighted variant and the prefix cache plan."""
from . import semiring, weighting
from .cache import CachePlan
from .cos import CosWISS
from .iss import ISS, ISSMode

__all__ = ["ISS", "ISSMode", "CosWISS", "CachePlan", "semiring", "weighting"]
