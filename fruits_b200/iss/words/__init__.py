from . import letters
from .creation import alternate_sign, of_weight, replace_letters
from .letters import ExtendedLetter, letter
from .word import SimpleWord, Word
