"""Letters of generic words (reference: ``fruits/iss/words/letters.py``).

A letter is user code: a Python function ``f(X, i)`` that maps one
multidimensional series ``X[n_dims, length]`` (and a dimension index) to one
row ``[length]``.  An :class:`ExtendedLetter` bundles letters whose rows are
combined by the semiring's operation.  The GPU path evaluates every distinct
extended letter once per batch on the host -- it is the caller's Python, like
the function of ``FUN`` -- and hands the rows to the ISS kernels as additional
input dimensions (``ISS._lettered``), so words over letters run through the
same trie / fused kernels as a ``SimpleWord``.
"""
__all__ = ["ExtendedLetter", "get_available", "letter"]

import functools
from typing import Callable, Optional

import numpy as np

# name -> unbound letter: ``unbound(dim)`` returns the function of one series
_REGISTRY: dict = {}


def _register(name: str, unbound: Callable) -> None:
    if name in _REGISTRY:
        raise RuntimeError(f"Letter with name '{name}' already exists")
    _REGISTRY[name] = unbound


def _lookup(name: str) -> Callable:
    try:
        return _REGISTRY[name]
    except KeyError:
        raise RuntimeError(f"Letter with name '{name}' does not exist") from None


def get_available() -> list:
    """Names usable in an :class:`ExtendedLetter` (reference :120-124)."""
    return list(_REGISTRY)


def _bind(func: Callable) -> Callable:
    """``func(X, i)`` -> ``unbound(i)(X)``, the calling convention of a word."""
    @functools.wraps(func)
    def unbound(i: int):
        def bound(X: np.ndarray) -> np.ndarray:
            return func(X, i)
        return bound
    return unbound


def letter(*args, name: Optional[str] = None):
    """Decorator that registers ``func(X, i) -> row`` as a letter, under its
    own name (``@letter``) or a given one (``@letter(name="ReLU")``)
    (reference :137-206)."""
    if len(args) > 1:
        raise RuntimeError("Too many arguments")
    if name is None:
        if len(args) == 1 and callable(args[0]):
            unbound = _bind(args[0])
            _register(args[0].__name__, unbound)
            return unbound
        raise ValueError("Please either specify the 'name' argument or use this "
                         "decorator without calling it.")

    def decorate(func: Callable):
        unbound = _bind(func)
        _register(name, unbound)
        return unbound
    return decorate


# the two predefined letters (reference :95-110)
_register("DIM", _bind(lambda X, i: X[i, :]))
_register("ABS", _bind(lambda X, i: np.abs(X[i, :])))


class ExtendedLetter:
    """Letters (by name) with the dimension each one is bound to; written like
    ``DIM(1)ABS(2)`` with 1-based dimensions (reference :12-92)."""

    def __init__(self, letter_string: str = "") -> None:
        self._letters: list = []
        self._dimensions: list = []
        self._string_repr = ""
        self._pos = -1
        for part in letter_string.split(")")[:-1]:
            name, dim = part.split("(")
            self.append(name, int(dim) - 1)

    def append(self, letter: str, dim: int = 0) -> None:
        """Adds the registered letter ``letter`` bound to the 0-based ``dim``."""
        self._letters.append(_lookup(letter))
        self._dimensions.append(dim)
        self._string_repr += f"{letter}({dim + 1})"

    def copy(self) -> "ExtendedLetter":
        twin = ExtendedLetter()
        twin._letters = list(self._letters)
        twin._dimensions = list(self._dimensions)
        twin._string_repr = self._string_repr
        return twin

    def __len__(self) -> int:
        return len(self._letters)

    def __getitem__(self, i: int) -> Callable:
        return self._letters[i](self._dimensions[i])

    def __iter__(self) -> "ExtendedLetter":
        self._pos = -1
        return self

    def __next__(self) -> Callable:
        if self._pos + 1 >= len(self._letters):
            raise StopIteration()
        self._pos += 1
        return self[self._pos]

    def __str__(self) -> str:
        return f"[{self._string_repr}]"
