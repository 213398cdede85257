"""Filtering preparateurs of the reference (``fruits/preparation/filter.py``:
DIL, WIN, DOT, PDD) are outside the accelerated hot path and raise."""
__all__ = ["DIL", "WIN", "DOT", "PDD"]

from .transform import _out_of_scope

DIL = _out_of_scope("DIL", "filter.py:13-73")
WIN = _out_of_scope("WIN", "filter.py:76-134")
DOT = _out_of_scope("DOT", "filter.py:137-196")
PDD = _out_of_scope("PDD", "filter.py:199-270")
