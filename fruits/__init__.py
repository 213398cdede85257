"""Drop-in alias: ``import fruits`` resolves to the B200-native
implementation ``fruits_b200`` (same public names as irkri/fruits 1.0.0).

Scope: ``Fruit`` / ``FruitSlice`` / ``ISS`` / ``CosWISS`` (plain and randomised)
with ``SimpleWord`` words and words over Python letters, the ``Reals`` /
``Arctic`` / ``Bayesian`` semirings, the ``Indices`` / ``L1`` / ``L2`` /
``Plateaus`` / ``Custom`` weightings (with or without a Python ``transform``),
all preparateurs (``INC, STD, NRM, MAV, LAG, FFN, RIN, RDW, JLD, SPE, RPE, CTS, QTC,
FUN, DIL, WIN, DOT, PDD`` and the wrappers ``NEW``, ``DIM``) and every sieve of the
reference.  All arithmetic runs on the GPU; user-supplied Python (``FUN``, letter
functions, lookup transforms) is called on the host with numpy arrays, as in the
reference.  There is no CPU fallback for anything else, so what is not built
raises ``NotImplementedError``: generic words with ``Arctic(argmax=True)``, more
than four distinct ``alpha`` values in one word and letters with more than 15
occurrences ("Limits")."""
import sys as _sys

import fruits_b200 as _impl
from fruits_b200 import *  # noqa: F401,F403
from fruits_b200 import (CosWISS, Fruit, FruitSlice, ISS, ISSMode, cache,  # noqa: F401
                         callback, iss, preparation, seed, semiring, sieving, words)

for _name, _mod in list(_sys.modules.items()):
    if _name == "fruits_b200" or _name.startswith("fruits_b200."):
        _sys.modules.setdefault("fruits" + _name[len("fruits_b200"):], _mod)
