"""Preparateurs: transformations of the input series ahead of the ISS.

All twenty preparateurs of the reference run on the GPU.  ``INC`` and ``STD``
(and ``NEW(INC)``) are folded into the loads of the fused kernels; the others
write a prepared copy with the streaming kernels of ``csrc/prep.cu`` /
``csrc/prep_more.cu``, after which the slice still takes a fused kernel.
``FUN`` calls the user's function on the host.
"""
from .abstract import Preparateur
from .filter import DIL, DOT, PDD, WIN
from .transform import (CTS, FFN, FUN, INC, JLD, LAG, MAV, NRM, QTC, RDW, RIN, RPE, SPE, STD)
from .wrapper import DIM, NEW

__all__ = ["Preparateur", "INC", "STD", "NRM", "NEW", "DIM", "DIL", "WIN", "DOT", "PDD", "MAV",
           "LAG", "FFN", "RIN", "RDW", "JLD", "SPE", "RPE", "CTS", "QTC", "FUN"]
