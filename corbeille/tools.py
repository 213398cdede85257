"""Index helper of the harness (reference:
``experiments/corbeille/corbeille/tools.py:8-52``)."""
from typing import Literal

import numpy as np

import fruits


def split_index(fruit: fruits.Fruit, index: int,
                level: Literal["prepared", "iterated sums", "features"] = "features") -> tuple:
    """Position of a flat ``index`` inside the fruit: ``(slice,)`` for the
    prepared data, ``(slice, iterated sum)``, or ``(slice, iterated sum, sieve,
    feature of that sieve)`` -- the counting order of ``Fruit.transform``'s
    columns.  ``ValueError`` beyond the end or for an unknown level."""
    if level == "prepared":
        if 0 <= index < len(fruit):
            return (index,)
    elif level in ("iterated sums", "features"):
        for s, slc in enumerate(fruit):
            n_sums = int(np.prod([iss.n_iterated_sums() for iss in slc.get_iss()]))
            widths = [sieve.nfeatures() for sieve in slc.get_sieves()]
            per_sum = 1 if level == "iterated sums" else sum(widths)
            if index >= n_sums * per_sum:
                index -= n_sums * per_sum
                continue
            if index < 0:
                break
            word, rest = divmod(index, per_sum)
            if level == "iterated sums":
                return (s, word)
            for k, width in enumerate(widths):
                if rest < width:
                    return (s, word, k, rest)
                rest -= width
    raise ValueError("Index out of range or unknown level")
