"""Words: the index patterns of the iterated sums (``SimpleWord`` strings such
as ``"[1][12][-2]"``), their enumeration by weight and helper constructors."""
from . import letters
from .creation import alternate_sign, of_weight, replace_letters
from .letters import ExtendedLetter, letter
from .word import SimpleWord, Word

__all__ = ["letters", "SimpleWord", "Word", "ExtendedLetter", "letter", "of_weight",
           "alternate_sign", "replace_letters"]
