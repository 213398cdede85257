"""Cache plan: which prefixes of which word are emitted in EXTENDED mode.

Same contract as the reference's ``fruits/iss/cache.py`` (:6-81): word ``i``
contributes its last ``plan[i]`` prefixes, where ``plan[i]`` is the number of
its prefixes (compared as bracket strings) that no earlier word starts with.
The GPU path turns this into a prefix trie (``fruits_b200/_plan.py``) so that
every prefix is computed once.
"""
from typing import Optional, Sequence

from .words.word import Word


class CachePlan:

    def __init__(self, words: Sequence[Word]) -> None:
        self._words = words
        strings = [str(w) for w in words]
        # first word index at which every prefix string appears
        first_seen: dict = {}
        self._plan: list = []
        for i, s in enumerate(strings):
            letters = s.split("]")[:-1]
            new = 0
            prefix = ""
            for el in letters:
                prefix += el + "]"
                if prefix not in first_seen:
                    first_seen[prefix] = i
                    new += 1
            self._plan.append(new)

    def unique_el_depth(self, index: int) -> int:
        """Number of iterated sums emitted for the word ``index``."""
        return self._plan[index]

    def get_word_index(self, is_index: int) -> int:
        """Word that emits the iterated sum ``is_index``."""
        for i, depth in enumerate(self._plan):
            is_index -= depth
            if is_index < 0:
                return i
        raise IndexError("Not enough iterated sums in cache plan")

    def get_word_string(self, is_index: int) -> str:
        """Prefix string of the iterated sum ``is_index``."""
        for i, depth in enumerate(self._plan):
            is_index -= depth
            if is_index < 0:
                letters = str(self._words[i]).split("]")
                return "]".join(letters[:int(is_index)]) + "]"
        raise IndexError("Not enough iterated sums in cache plan")

    def n_iterated_sums(self, word_indices: Optional[Sequence[int]] = None) -> int:
        if word_indices is None:
            return sum(self._plan)
        return sum(self._plan[i] for i in word_indices)
