"""Sieve wrappers of the reference (``fruits/sieving/wrapper.py``: INC, INT)
are outside the accelerated hot path."""
__all__ = []
