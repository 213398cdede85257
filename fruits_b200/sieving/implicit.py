"""Implicit sieves (reference: ``fruits/sieving/implicit.py``): ``PPV``
(:11-153) and ``CPV`` (:156-213)."""
__all__ = ["PPV", "CPV"]

from typing import Union

import numpy as np
import torch

from .. import _backend as be
from .abstract import FeatureSieve, quantile_rows


class PPV(FeatureSieve):
    """Proportion of values ``>=`` a fitted quantile (deprecated in the
    reference in favour of NPI, kept for drop-in compatibility).

    Args as in the reference (implicit.py:25-53): ``quantile`` (value or
    probability, or a list), ``constant`` (interpret as value), ``sample_size``
    (fraction of series used for the quantile), ``segments``."""

    _mode = 0   # bit 1 of the kernel's mode word: connected components (CPV)

    def __init__(self, quantile: Union[list, float] = 0.5,
                 constant: Union[list, bool] = False, sample_size: float = 1.0,
                 segments: bool = False) -> None:
        many = isinstance(quantile, list)
        qs = list(quantile) if many else [quantile]
        if isinstance(constant, list):
            if many and len(constant) != len(qs):
                raise ValueError("'quantile' and 'constant' must be lists of the same length "
                                 "(or 'constant' a single boolean)")
            if not many and len(constant) > 1:
                raise ValueError("a single 'quantile' takes a single boolean 'constant'")
            cs = list(constant)
        else:
            cs = [constant] * len(qs)
        if many and any(not c and not 0 <= q <= 1 for q, c in zip(qs, cs)):
            raise ValueError("a 'quantile' that is not 'constant' is a probability in [0, 1]")
        if not 0 < sample_size <= 1:
            raise ValueError("'sample_size' has to be a float in (0, 1]")
        if segments and len(qs) == 1:
            raise ValueError("'segments' needs a list of at least two quantiles")
        # segments: distinct quantiles in ascending order, paired with the flags
        # in their given order (reference :84-86)
        self._q_c_input = (sorted(zip(list(set(qs)), cs), key=lambda qc: qc[0]) if segments
                           else list(zip(qs, cs)))
        self._sample_size = sample_size
        self._segments = segments

    def _nfeatures(self) -> int:
        if self._segments:
            return len(self._q_c_input) - 1
        return len(self._q_c_input)

    def _draw(self, n_rows: int):
        """Consume the global numpy RNG exactly like the reference does
        (implicit.py:103-108): one ``choice`` per non-constant quantile."""
        draws = []
        for q, const in self._q_c_input:
            if const:
                draws.append(None)
            else:
                size = max(int(self._sample_size * n_rows), 1)
                draws.append(np.random.choice(np.arange(n_rows), size=size, replace=False))
        return draws

    def _fit_device(self, X: torch.Tensor) -> None:
        X = X.contiguous()
        draws = self._draw(X.shape[0])
        self._q = [x[0] for x in self._q_c_input]
        for i, (q, const) in enumerate(self._q_c_input):
            if const:
                continue
            sel = draws[i]
            if len(sel) == X.shape[0]:
                rows = X   # a permutation of all rows: same multiset of values
            else:
                idx = torch.as_tensor(sel, device=X.device, dtype=torch.long)
                rows = X.index_select(0, idx)
            self._q[i] = quantile_rows(rows.reshape(1, -1), q)[0]

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_q"):
            raise RuntimeError("Missing call of PPV.fit()")
        arr = X.contiguous()
        out = be.empty((arr.shape[0], self.nfeatures()))
        self._apply(arr, out, 0)
        return out

    def _apply(self, arr: torch.Tensor, out: torch.Tensor, col0: int) -> None:
        n, t = arr.shape
        q = be.to_device(np.array(self._q, dtype=np.float64))
        be.check(be.lib().fb_ppv(arr.data_ptr(), arr.stride(0), q.data_ptr(), len(self._q),
                                 int(self._segments) | self._mode, out.data_ptr(), out.stride(0), col0,
                                 n, t, be.stream_ptr()))

    def _fused(self):
        if not self._segments and len(self._q_c_input) == 1:
            return ("PPV", 0)
        return None

    def _copy(self) -> "PPV":
        return PPV(quantile=[x[0] for x in self._q_c_input],
                   constant=[x[1] for x in self._q_c_input],
                   sample_size=self._sample_size, segments=self._segments)

    def _summary(self) -> str:
        string = f"PPV [sampling={self._sample_size}"
        if self._segments:
            string += ", segments"
        string += f"] -> {self.nfeatures()}:"
        for x in self._q_c_input:
            string += f"\n   > {x[0]} | {x[1]}"
        return string

    def __str__(self) -> str:
        return ("PPV("
                f"quantile={[x[0] for x in self._q_c_input]}, "
                f"constant={[x[1] for x in self._q_c_input]}, "
                f"sample_size={self._sample_size}, "
                f"segments={self._segments})")


class CPV(PPV):
    """Proportion of connected components of values ``>=`` a fitted quantile
    (reference :156-213): the number of rising edges of the indicator
    ``X >= q`` (a series that starts above ``q`` does not open one -- the
    increments are zero-padded), times two, over the even-rounded length.
    Arguments as :class:`PPV`."""
    _mode = 2

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_q"):
            raise RuntimeError("Missing call of CPV.fit()")
        return super()._transform_device(X)

    def _fused(self):
        if not self._segments and len(self._q_c_input) == 1:
            return ("CPV", 0)
        return None

    def _copy(self) -> "CPV":
        return CPV([x[0] for x in self._q_c_input], [x[1] for x in self._q_c_input],
                   self._sample_size, self._segments)

    def _summary(self) -> str:
        return "C" + super()._summary()[1:]

    def __str__(self) -> str:
        return "C" + super().__str__()[1:]
