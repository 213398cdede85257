"""Semirings of the iterated sums (reference: ``fruits/iss/semiring.py``).

On the GPU a semiring is only a tag that selects the instantiation of the ISS
kernel (``csrc/lns.cuh``): ``Reals`` (:161-231, sum / product, the standard
iterated sums) and ``Arctic`` (:341-457, max / plus); ``Bayesian``
(:461-601, max / times) has its own scan kernel (``csrc/bayes.cu``), and so
has ``Arctic(argmax=True)`` (``csrc/arctic_argmax.cu``).
"""
from abc import ABC

from .. import _backend as be


class Semiring(ABC):
    _code: int = -1

    def __str__(self) -> str:
        return self.__class__.__name__


class Reals(Semiring):
    """Field of real numbers with the usual sum and product (default)."""
    _code = be.SEMIRING_REALS


class Arctic(Semiring):
    """Max-plus semiring: "sum" is the maximum, "product" the addition.

    ``argmax=True`` additionally returns the positions of all involved maxima
    (reference :234-279): per word of ``p`` letters ``p + p(p+1)/2`` rows, in
    ``ISSMode.EXTENDED`` only."""
    _code = be.SEMIRING_ARCTIC

    def __init__(self, argmax: bool = False) -> None:
        self._argmax = bool(argmax)


class Bayesian(Semiring):
    """Max-times semiring on [0, 1] (reference :461-601): "sum" is the
    maximum, "product" the multiplication.  Evaluated by a parallel running
    maximum (``csrc/bayes.cu``); slices over this semiring are sieved on the
    materialised iterated sums."""
    _code = be.SEMIRING_BAYESIAN
