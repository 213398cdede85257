"""Generic letters (reference: fruits/iss/words/letters.py) are Python
callables evaluated per time series; they cannot be compiled for the device
and are outside the accelerated hot path."""

__all__ = ["ExtendedLetter", "get_available", "letter"]


class ExtendedLetter:
    def __init__(self, letter_string: str = "") -> None:
        raise NotImplementedError(
            "ExtendedLetter holds Python letter functions; only SimpleWord is "
            "supported by the GPU path")


def letter(*args, **kwargs):
    raise NotImplementedError(
        "custom letters are Python callables; only SimpleWord is supported "
        "by the GPU path")


def get_available() -> list:
    return ["DIM"]
