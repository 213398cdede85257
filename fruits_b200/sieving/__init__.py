"""Feature sieves: numbers extracted from one iterated sum per series.

Segment sieves (cuts + quantile intervals), increment sieves (on the
zero-padded increments), implicit sieves (fitted quantiles) and the two
wrappers; all evaluated by ``csrc/sieve.cu`` on materialised sums, the common
ones also inside the fused kernels.
"""
from .abstract import FeatureSieve
from .implicit import CPV, PPV
from .increment import LPI, MPI, NPI, XPI
from .segment import AVG, CUR, END, MAX, MIN, STD
from .wrapper import INC, INT

__all__ = ["FeatureSieve", "NPI", "MPI", "XPI", "LPI", "MAX", "MIN", "END", "CUR", "AVG", "STD",
           "PPV", "CPV", "INC", "INT"]
