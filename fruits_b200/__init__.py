"""fruits_b200 -- B200-native implementation of the FRUITS hot path.

Drop-in for the Python API of irkri/fruits 1.0.0 on the path
preparateur -> iterated-sums signature -> sieves (``Fruit`` / ``FruitSlice``
``fit`` / ``transform``, ``ISS`` with ``SimpleWord`` words or words over Python
letters, ``ISSMode``, the ``Reals`` / ``Arctic`` / ``Bayesian`` semirings,
exponential weightings, ``CosWISS``, all preparateurs and sieves of the
reference).  All arithmetic runs in hand-written sm_100a CUDA
(``fruits_b200/csrc``) behind the C ABI of ``include/fruits_b200.h``.
"""
from . import cache, callback, iss, preparation, seed, sieving
from .fruit import Fruit, FruitSlice
from .iss import semiring, words
from .iss.cos import CosWISS
from .iss.iss import ISS, ISSMode

__version__ = "0.1.0"
