"""Word generators (reference: fruits/iss/words/creation.py).

The enumeration ORDER is part of the contract -- it fixes the column order of
the feature matrix -- so ``of_weight`` walks partitions and their
permutations exactly like the reference does (creation.py:9-13, :44-45),
including the use of a ``set`` of permutations whose iteration order is
CPython's hash order for tuples of small ints.
"""
import itertools
from collections.abc import Sequence

from .letters import ExtendedLetter
from .word import SimpleWord, Word


def _partitions(n: int, smallest: int = 1):
    yield (n,)
    for head in range(smallest, n // 2 + 1):
        for tail in _partitions(n - head, head):
            yield (head,) + tail


def _letters_of_weight(w: int, dim: int) -> list:
    out = []
    for combo in itertools.combinations_with_replacement(range(1, dim + 1), w):
        out.append("[" + "".join(str(x) if x < 10 else f"({x})" for x in combo) + "]")
    return out


def of_weight(w: int, dim: int = 1) -> tuple:
    """All words with exactly ``w`` letters over ``dim`` dimensions
    (reference: creation.py:26-50)."""
    by_weight = [_letters_of_weight(i, dim) for i in range(1, w + 1)]
    words = []
    for partition in _partitions(w):
        for order in set(itertools.permutations(partition)):
            for combo in itertools.product(*(by_weight[k - 1] for k in order)):
                words.append(SimpleWord("".join(combo)))
    return tuple(words)


def alternate_sign(words: Sequence[SimpleWord]) -> list:
    """For every word two words with alternating signs of the exponents
    (reference: creation.py:86-103)."""
    out = []
    for word in words:
        first, second = "", ""
        for i, expo in enumerate(word):
            neg = "".join(int(c) * f"-{d + 1}" for d, c in enumerate(expo))
            pos = neg.replace("-", "")
            first += f"[{neg if i % 2 == 0 else pos}]"
            second += f"[{pos if i % 2 == 0 else neg}]"
        out.append(SimpleWord(first))
        out.append(SimpleWord(second))
    return out


def replace_letters(word: SimpleWord, letter_gen) -> Word:
    """A generic word with the shape of ``word`` whose letters are the names
    ``letter_gen`` yields, one per letter occurrence in ascending dimension
    order; ``DIM`` once the iterator is exhausted (reference: creation.py:53-83)."""
    out = Word()
    for expo in word:
        el = ExtendedLetter()
        for dim, count in enumerate(expo):
            for _ in range(count):
                el.append(next(letter_gen, "DIM"), dim)
        out.multiply(el)
    return out
