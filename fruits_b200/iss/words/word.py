"""Words: host-side metadata of the iterated sums.

Mirrors the public behaviour of the reference's ``fruits/iss/words/word.py``
(``SimpleWord`` :128-268, ``Word`` :9-125).  A ``SimpleWord`` is a sequence of
extended letters; each extended letter is stored as its exponent vector over
the input dimensions (``"[112][2]"`` -> ``[[2, 1], [0, 1]]``), negative
exponents meaning division.  A generic ``Word`` is a sequence of
:class:`~fruits_b200.iss.words.letters.ExtendedLetter` objects (Python
functions of one series); ``ISS`` evaluates those on the host and feeds their
rows to the same kernels as extra input dimensions.
"""
import re
from typing import Optional, Sequence

import numpy as np

_WORD_PATTERN = r"(\[(-?\d|\(-?\d+\))+\])+"
_TOKEN = re.compile(r"\((-?\d+)\)|(-\d?)|(\d)")


class Word:
    """A sequence of extended letters, e.g. ``Word("[ABS(1)DIM(2)][DIM(1)]")``
    (reference: word.py:9-125)."""

    def __init__(self, word_string: Optional[str] = None) -> None:
        self._extended_letters: list = []
        self._alpha: Optional[np.ndarray] = None
        self._iter = -1
        if word_string is not None:
            self.multiply(word_string)

    @property
    def alpha(self) -> np.ndarray:
        """Alpha values of a weighted iterated sum; ones by default
        (reference: word.py:71-76)."""
        if self._alpha is None:
            return np.ones((len(self),), dtype=np.float32)
        return self._alpha

    @alpha.setter
    def alpha(self, alpha: Sequence[float]) -> None:
        if len(alpha) != len(self):
            raise ValueError("Size of alpha array does not match word length")
        self._alpha = np.array(alpha, dtype=np.float32)

    def multiply(self, other) -> None:
        """Appends an extended letter, the letters of another word or of a
        string like ``"[DIM(1)][ABS(2)DIM(1)]"`` (reference: word.py:84-98)."""
        from .letters import ExtendedLetter
        if isinstance(other, ExtendedLetter):
            self._extended_letters.append(other)
        elif isinstance(other, Word):
            self._extended_letters.extend(other._extended_letters)
        elif isinstance(other, str):
            for part in other.split("]")[:-1]:
                self._extended_letters.append(ExtendedLetter(part[1:]))
        else:
            raise TypeError(f"Cannot multiply Word with {type(other)}")

    def copy(self) -> "Word":
        twin = Word()
        twin._extended_letters = [el.copy() for el in self._extended_letters]
        return twin

    def __len__(self) -> int:
        return len(self._extended_letters)

    def __iter__(self):
        self._iter = -1
        return self

    def __next__(self):
        if self._iter < len(self._extended_letters) - 1:
            self._iter += 1
            return self._extended_letters[self._iter]
        raise StopIteration()

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, Word):
            raise NotImplementedError
        return False

    def __str__(self) -> str:
        return "".join(str(el) for el in self._extended_letters)


class SimpleWord(Word):
    """Word whose letters pick single input dimensions, written like
    ``"[11][122]"`` (dimensions are 1-based, ``(10)`` for two-digit
    dimensions, a leading ``-`` for a negative exponent).  The string has to
    match ``(\\[(-?\\d|\\(-?\\d+\\))+\\])+`` (reference: word.py:189-206)."""

    def __init__(self, string: str) -> None:
        super().__init__()
        self._max_dim = 0
        self._name = ""
        self.multiply(string)

    def multiply(self, other) -> None:
        if not isinstance(other, str):
            raise NotImplementedError
        if not re.fullmatch(_WORD_PATTERN, other):
            raise ValueError("SimpleWord can only be multiplied with a "
                             "string matching the regular expression "
                             r"'(\[(-?\d|\(-?\d+\))+\])+'")
        self._name += other
        parsed = []
        for chunk in other.split("]")[:-1]:
            letters = []
            for par, neg, pos in _TOKEN.findall(chunk[1:]):
                if pos:
                    letters.append(int(pos))
                elif neg:
                    letters.append(-1 if neg == "-" else int(neg))
                else:
                    letters.append(int(par))
            parsed.append(letters)
        width = max(abs(x) for el in parsed for x in el)
        if width > self._max_dim:
            for el in self._extended_letters:
                el.extend([0] * (width - self._max_dim))
            self._max_dim = width
        for letters in parsed:
            expo = [0] * self._max_dim
            for x in letters:
                expo[abs(x) - 1] += 1 if x > 0 else -1
            self._extended_letters.append(expo)

    def copy(self) -> "SimpleWord":
        sw = SimpleWord(self._name)
        sw._extended_letters = [list(el) for el in self._extended_letters]
        return sw

    def exponents(self) -> np.ndarray:
        """``int32[p, max_dim]`` exponent matrix (what the reference passes to
        its kernels, semiring.py:30)."""
        return np.array(self._extended_letters, dtype=np.int32)

    def __eq__(self, other: object) -> bool:
        if not isinstance(other, SimpleWord):
            raise NotImplementedError
        return list(self._extended_letters) == list(other._extended_letters)

    def __str__(self) -> str:
        return self._name
