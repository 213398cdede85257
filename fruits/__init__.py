"""Drop-in alias: ``import fruits`` resolves to the B200-native
implementation ``fruits_b200`` (same public names as irkri/fruits 1.0.0)."""
import sys as _sys

import fruits_b200 as _impl
from fruits_b200 import *  # noqa: F401,F403
from fruits_b200 import (CosWISS, Fruit, FruitSlice, ISS, ISSMode, cache,  # noqa: F401
                         callback, iss, preparation, seed, semiring, sieving, words)

for _name, _mod in list(_sys.modules.items()):
    if _name == "fruits_b200" or _name.startswith("fruits_b200."):
        _sys.modules.setdefault("fruits" + _name[len("fruits_b200"):], _mod)
