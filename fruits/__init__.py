"""Drop-in alias: ``import fruits`` resolves to the B200-native
implementation ``fruits_b200`` (same public names as irkri/fruits 1.0.0).

Scope: the alias is a drop-in for the *hot path* -- ``Fruit`` / ``FruitSlice`` /
``ISS`` / ``CosWISS`` with ``SimpleWord`` words, the ``Reals`` / ``Arctic`` /
``Bayesian`` semirings, the ``Indices`` / ``L1`` / ``L2`` / ``Plateaus`` weightings,
the preparateurs ``INC``, ``STD``, ``NRM``, ``NEW``, ``DIM`` and every sieve of
the reference.  Everything else of the reference's surface exists by name but
raises ``NotImplementedError`` when constructed, because there is no CPU
fallback to run it on: the preparateurs ``MAV, LAG, FFN, RIN, RDW, JLD, SPE, RPE,
CTS, QTC, FUN, DIL, WIN, DOT, PDD``, ``NRM(scale_dim=True)``, weightings with a
Python ``transform``, words with callable letters, ``Arctic(argmax=True)``, the
randomised ``CosWISS`` variants, more than four distinct ``alpha`` values per
ISS and letters with more than 15 occurrences ("Limits")."""
import sys as _sys

import fruits_b200 as _impl
from fruits_b200 import *  # noqa: F401,F403
from fruits_b200 import (CosWISS, Fruit, FruitSlice, ISS, ISSMode, cache,  # noqa: F401
                         callback, iss, preparation, seed, semiring, sieving, words)

for _name, _mod in list(_sys.modules.items()):
    if _name == "fruits_b200" or _name.startswith("fruits_b200."):
        _sys.modules.setdefault("fruits" + _name[len("fruits_b200"):], _mod)
