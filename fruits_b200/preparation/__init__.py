"""Preparateurs: transformations of the input series ahead of the ISS.

``INC``, ``STD``, ``NRM`` and the wrappers ``NEW`` / ``DIM`` run on the GPU
(``csrc/prep.cu``; ``INC`` and ``STD`` are folded into the loads of the fused
kernels); the other names of the reference exist and raise
``NotImplementedError`` -- they are outside the accelerated path.
"""
from .abstract import Preparateur
from .filter import DIL, DOT, PDD, WIN
from .transform import (CTS, FFN, FUN, INC, JLD, LAG, MAV, NRM, QTC, RDW, RIN, RPE, SPE, STD)
from .wrapper import DIM, NEW

__all__ = ["Preparateur", "INC", "STD", "NRM", "NEW", "DIM", "DIL", "WIN", "DOT", "PDD", "MAV",
           "LAG", "FFN", "RIN", "RDW", "JLD", "SPE", "RPE", "CTS", "QTC", "FUN"]
