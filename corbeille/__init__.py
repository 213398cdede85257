"""Caller side of the hot path: the timing
harness and the UCR ``.txt`` loader of the reference's ``corbeille`` extension
(``experiments/corbeille/corbeille``), so that published-style experiments run
against the GPU path::

    import corbeille
    from experiments.fruit_reduced import fruit
    data = corbeille.data.load("path/to/UCR/Chinatown")
    seconds, accuracy = corbeille.fruitify(data, fruit)

Mirrored: ``fruitify`` / ``fruitify_all`` / ``decide_which_fruit``, the whole
``data`` module (``.txt`` and ``.arff`` readers, ``multisine``, the resampling
helpers) and ``tools.split_index``; the analysis class ``Fruitalyser`` (matplotlib
plots) is outside the path.
"""
from . import data, tools
from .fruitifier import decide_which_fruit, fruitify, fruitify_all

__all__ = ["data", "tools", "fruitify", "fruitify_all", "decide_which_fruit"]
