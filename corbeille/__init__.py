"""Caller side of the hot path: the timing
harness and the UCR ``.txt`` loader of the reference's ``corbeille`` extension
(``experiments/corbeille/corbeille``), so that published-style experiments run
against the GPU path::

    import corbeille
    from experiments.fruit_reduced import fruit
    data = corbeille.data.load("path/to/UCR/Chinatown")
    seconds, accuracy = corbeille.fruitify(data, fruit)

Only ``fruitify`` / ``fruitify_all`` and ``data.load`` / ``data.load_all`` /
``data.replace_nan`` are mirrored; the analysis classes (``Fruitalyser``,
plots) are outside the accelerated path.
"""
from . import data
from .fruitifier import fruitify, fruitify_all

__all__ = ["data", "fruitify", "fruitify_all"]
