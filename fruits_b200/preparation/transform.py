"""Preparateurs that transform every dimension (reference:
``fruits/preparation/transform.py``).  On the hot path: ``INC`` (:15-89) and
``STD`` (:92-158), which the fused kernels apply while they load the input;
``NRM`` (:161-209) is used by the weighting lookups.  The remaining
preparateurs of the reference (MAV, LAG, FFN, RIN, RDW, JLD, SPE, RPE, CTS,
QTC, FUN) write a prepared copy with the streaming kernels of
``csrc/prep_more.cu``; their ``fit`` draws from the global numpy RNG with the
reference's own calls, in the reference's order, so a seeded fit ends in the
same weights and the same generator state.
"""
__all__ = ["INC", "STD", "NRM", "MAV", "LAG", "FFN", "RIN", "RDW", "JLD",
           "SPE", "RPE", "CTS", "QTC", "FUN"]

from typing import Any, Callable, Literal, Optional, Union

import numpy as np
import torch

from .. import _backend as be
from .abstract import Preparateur


def increments_device(X: torch.Tensor, k: int, pad_src=None) -> torch.Tensor:
    """``_increments(X, k)`` of the reference (cache.py:8-13) on the GPU."""
    X = X.contiguous()
    out = torch.empty_like(X)
    rows = X.numel() // X.shape[-1] if X.numel() else 0
    be.check(be.lib().fb_increments(X.data_ptr(), be.ptr(pad_src), out.data_ptr(), rows,
                                    X.shape[-1], int(k), be.stream_ptr()))
    return out


class INC(Preparateur):
    """Increments ``[0, x_2-x_1, ..., x_n-x_{n-1}]`` (reference:
    transform.py:15-89; same arguments)."""

    def __init__(self, shift: Union[int, float, Callable[[int], int]] = 1,
                 depth: int = 1, zero_padding: bool = True) -> None:
        self._shift = shift
        if depth < 1:
            raise ValueError("depth has to be a positive integer > 0")
        self._depth = depth
        self._zero_padding = zero_padding

    @property
    def requires_fitting(self) -> bool:
        return False

    def _resolve_shift(self, length: int) -> int:
        if isinstance(self._shift, int):
            return self._shift
        if isinstance(self._shift, float):
            return int(np.ceil(self._shift * length))
        if callable(self._shift):
            return int(self._shift(length))
        raise TypeError(f"Type {type(self._shift)} not supported for argument shift")

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        shift = self._resolve_shift(X.shape[2])
        X = X.contiguous()
        out = X
        for _ in range(self._depth):
            out = increments_device(out, shift, None if self._zero_padding else X)
        return out

    def _fusable(self):
        if (isinstance(self._shift, int) and self._shift == 1 and self._depth == 1
                and self._zero_padding):
            return "inc"
        return None

    def _copy(self) -> "INC":
        return INC(self._shift, self._depth, self._zero_padding)

    def __eq__(self, other) -> bool:
        return (isinstance(other, INC) and self._shift == other._shift
                and self._depth == other._depth
                and self._zero_padding == other._zero_padding)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"INC({self._shift}, {self._depth}, {self._zero_padding})"


class STD(Preparateur):
    """Standardisation (reference: transform.py:92-158).  ``separately=True``
    standardises every series and dimension on its own with numpy's pairwise
    mean / std; ``separately=False`` uses one global mean / std from fit."""

    def __init__(self, separately: bool = True, var: bool = True,
                 std_eps: float = 1e-5) -> None:
        self._separately = separately
        self._div_std = var
        self._mean = None
        self._std = None
        self._eps = std_eps

    def _fit_device(self, X: torch.Tensor) -> None:
        if not self._separately:
            # one global pairwise mean / std over the flattened array
            flat = X.contiguous().reshape(1, -1)
            stats = self._row_stats(flat, self._div_std, 0.0).cpu().numpy()
            self._mean = float(stats[0, 0])
            self._std = float(stats[0, 1]) if self._div_std else 1

    @staticmethod
    def _row_stats(rows2d: torch.Tensor, div_std: bool, eps: float) -> torch.Tensor:
        rows, t = rows2d.shape
        if t >= 65536 * 1024:
            raise NotImplementedError("row too long")
        stats = be.empty((rows, 2))
        be.check(be.lib().fb_row_stats(rows2d.data_ptr(), stats.data_ptr(), rows, t,
                                       int(div_std), float(eps), be.stream_ptr()))
        return stats

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        X = X.contiguous()
        n, d, t = X.shape
        if not self._separately:
            if self._mean is None or self._std is None:
                raise RuntimeError("Missing call of self.fit()")
            stats = torch.tensor([[self._mean, self._std + self._eps]],
                                 dtype=torch.float64, device=X.device).repeat(n * d, 1)
        else:
            stats = self._row_stats(X.reshape(n * d, t), self._div_std, self._eps)
        out = torch.empty_like(X)
        be.check(be.lib().fb_standardize(X.data_ptr(), stats.data_ptr(), out.data_ptr(),
                                         n * d, t, be.stream_ptr()))
        return out

    def _fusable(self):
        return "std" if self._separately else None

    def _row_independent_fit(self) -> bool:
        return bool(self._separately)      # else: one mean / std over the whole sample

    def _copy(self) -> "STD":
        # like the reference (transform.py:146-147) the copy drops std_eps
        return STD(self._separately, self._div_std)

    def __eq__(self, other: Any) -> bool:
        return (isinstance(other, STD) and self._separately == other._separately
                and self._div_std == other._div_std)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"STD({self._separately}, {self._div_std})"


class NRM(Preparateur):
    """Min-max normalisation per series and dimension, or per series over all
    dimensions with ``scale_dim=True`` (reference: transform.py:161-209)."""

    def __init__(self, scale_dim: bool = False) -> None:
        self._scale_dim = scale_dim

    @property
    def requires_fitting(self) -> bool:
        return False

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        X = X.contiguous()
        n, d, t = X.shape
        out = torch.empty_like(X)
        # scale_dim: one minimum / maximum per series over all dimensions and time
        # steps (transform.py:187-189) -- the same kernel on rows of d * t values
        rows, length = (n, d * t) if self._scale_dim else (n * d, t)
        be.check(be.lib().fb_nrm_scale(X.data_ptr(), out.data_ptr(), rows, length, 0, 1.0,
                                       be.stream_ptr()))
        return out

    def _copy(self) -> "NRM":
        return NRM(scale_dim=self._scale_dim)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, NRM) and self._scale_dim == other._scale_dim

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"NRM({self._scale_dim})"


def _dev(a, dtype) -> torch.Tensor:
    """Small fitted host array -> device tensor (weights, index lists)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(be.require_cuda())


def _split_dims(n_in: int, n_out: int) -> np.ndarray:
    """Input dimensions per output dimension, as equal as possible
    (transform.py:500-504, :699-703)."""
    quotient, remainder = divmod(n_in, n_out)
    return np.array([quotient + 1] * remainder + [quotient] * (n_out - remainder),
                    dtype=np.int32)


class MAV(Preparateur):
    """Moving average over ``width`` time steps, zeros in front (reference:
    transform.py:212-274).  ``width=-1`` (the average over the dimensions of the
    reference's docstring) never gets past the "Missing call of self.fit()"
    check in the reference either (:250-262: ``fit`` sets no width for it)."""

    def __init__(self, width: Union[int, float] = 5) -> None:
        if isinstance(width, float) and not 0.0 < width < 1.0:
            raise ValueError("If width is a float, it has to be in (0,1)")
        self._w_given = width

    def _fit_device(self, X: torch.Tensor) -> None:
        if isinstance(self._w_given, float):
            self._w = max(int(self._w_given * X.shape[2]), 1)
        elif self._w_given > 0:
            self._w = self._w_given

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_w"):
            raise RuntimeError("Missing call of self.fit()")
        X = X.contiguous()
        n, d, t = X.shape
        out = torch.empty_like(X)
        be.check(be.lib().fb_moving_average(X.data_ptr(), out.data_ptr(), n * d, t, int(self._w),
                                            be.stream_ptr()))
        return out

    def _copy(self) -> "MAV":
        return MAV(self._w_given)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, MAV) and self._w_given == other._w_given

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"MAV({self._w_given})"


class LAG(Preparateur):
    """Lead-lag transform: every dimension becomes a lead and a lag dimension
    of length ``2T - 1`` (reference: transform.py:277-309)."""

    @property
    def requires_fitting(self) -> bool:
        return False

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        X = X.contiguous()
        n, d, t = X.shape
        out = be.empty((n, 2 * d, 2 * t - 1))
        be.check(be.lib().fb_lead_lag(X.data_ptr(), out.data_ptr(), n * d, t, be.stream_ptr()))
        return out

    def _copy(self) -> "LAG":
        return LAG()

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, LAG)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return "LAG()"


class FFN(Preparateur):
    """Two-layer network with gaussian weights applied to every time step
    (reference: transform.py:312-388)."""

    def __init__(self, d_out: int = 1, d_hidden: Optional[int] = None, center: bool = True,
                 relu_out: bool = False) -> None:
        self._d_hidden = d_hidden
        self._d_out = d_out
        self._center = center
        self._relu_out = relu_out

    def _fit_device(self, X: torch.Tensor) -> None:
        d = X.shape[1]
        d_hidden = 2 * d if self._d_hidden is None else self._d_hidden
        # the reference's three draws, in its order (transform.py:346-360)
        self._weights1 = np.random.normal(loc=0, scale=1.0, size=(d_hidden, d))
        self._biases = np.random.normal(loc=0, scale=1.0, size=(d_hidden,))
        self._weights2 = np.random.normal(loc=0, scale=1.0, size=(self._d_out, d_hidden))

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_weights1"):
            raise RuntimeError("FFN was not fitted")
        X = X.contiguous()
        n, d, t = X.shape
        if self._weights1.shape[1] != d:
            raise ValueError(f"FFN was fitted on {self._weights1.shape[1]} dimensions, got {d}")
        # (uploaded per call: the fitted arrays are public attributes a caller may replace)
        w1, b1, w2 = (_dev(w, np.float64) for w in (self._weights1, self._biases, self._weights2))
        mean = STD._row_stats(X.reshape(n * d, t), False, 0.0) if self._center and n else None
        out = be.empty((n, self._d_out, t))
        be.check(be.lib().fb_ffn(X.data_ptr(), be.ptr(mean), w1.data_ptr(), b1.data_ptr(),
                                 w2.data_ptr(), out.data_ptr(), n, d, t, w1.shape[0],
                                 self._d_out, int(self._relu_out), be.stream_ptr()))
        return out

    def _copy(self) -> "FFN":
        return FFN(d_out=self._d_out, d_hidden=self._d_hidden, center=self._center,
                   relu_out=self._relu_out)

    def __str__(self) -> str:
        return f"FFN({self._d_out}, {self._d_hidden}, {self._center}, {self._relu_out})"


class RIN(Preparateur):
    """Random increments ``y_i = x_i - (k_w x_{i-1} + ... + k_1 x_{i-w})`` with a
    kernel drawn in ``fit`` (reference: transform.py:391-568)."""

    def __init__(self, width: Union[int, Callable[[int], int]] = 1,
                 adaptive_width: bool = False, out_dim: int = -1, force_sum_one: bool = False,
                 kernel: Optional[np.ndarray] = None) -> None:
        self._width = width
        self._adaptive_width = adaptive_width
        self._out_dim = out_dim
        self._force_sum_one = force_sum_one
        self._const_kernel = kernel

    def _fit_device(self, X: torch.Tensor) -> None:
        d, t = X.shape[1], X.shape[2]
        if self._const_kernel is not None:
            self._kernel = self._const_kernel.copy()
            self._ndim_per_kernel = np.ones((d,), dtype=np.int32)
            self._dims_per_kernel = np.arange(d, dtype=np.int32)
            return
        width = self._width(t) if callable(self._width) else min(self._width, t - 1)
        out_dim = self._out_dim if self._out_dim > 0 else d
        if out_dim > d:
            raise ValueError(f"Output dimensions ({out_dim}) should be <= input dimensions ({d})")
        self._ndim_per_kernel = _split_dims(d, out_dim)
        self._dims_per_kernel = np.random.choice(d, size=d, replace=False).astype(np.int32)
        if self._force_sum_one:
            while True:      # transform.py:509-520
                self._kernel = np.random.uniform(-1., 1., size=(d, width))
                change = 1.0 - np.sum(self._kernel, axis=1)
                diff = 1.0 - np.abs(self._kernel)
                diffsum = np.sum(diff, axis=1)
                if np.sum(diffsum < 1e-5) > 0:
                    continue
                self._kernel += diff * (change / diffsum)[:, np.newaxis]
                break
        else:
            self._kernel = np.random.normal(size=(d, width))
            self._kernel -= np.mean(self._kernel, axis=1)[:, np.newaxis]

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_kernel"):
            raise RuntimeError("RIN preparateur misses a .fit() call")
        X = X.contiguous()
        n, d, t = X.shape
        kern = np.ascontiguousarray(self._kernel, dtype=np.float64)
        if kern.ndim != 2 or kern.shape[0] < int(self._ndim_per_kernel.sum()):
            raise ValueError("RIN kernel needs one row per input dimension")
        if int(self._dims_per_kernel.max(initial=0)) >= d or len(self._dims_per_kernel) > d:
            raise IndexError(f"RIN was fitted on {len(self._dims_per_kernel)} dimensions, got {d}")
        k_d, ndim_d, dims_d = (_dev(kern, np.float64), _dev(self._ndim_per_kernel, np.int32),
                               _dev(self._dims_per_kernel, np.int32))
        n_out, w = len(self._ndim_per_kernel), kern.shape[1]
        out = be.empty((n, n_out, t))
        be.check(be.lib().fb_random_increments(
            X.data_ptr(), k_d.data_ptr(), ndim_d.data_ptr(), dims_d.data_ptr(), out.data_ptr(),
            n, d, t, n_out, w, w if self._adaptive_width else 0, be.stream_ptr()))
        return out

    def _copy(self) -> "RIN":
        return RIN(width=self._width, adaptive_width=self._adaptive_width,
                   out_dim=self._out_dim, force_sum_one=self._force_sum_one,
                   kernel=self._const_kernel)

    def __eq__(self, other: Any) -> bool:
        return bool(isinstance(other, RIN) and self._width == other._width
                    and self._adaptive_width == other._adaptive_width
                    and self._out_dim == other._out_dim
                    and self._force_sum_one == other._force_sum_one
                    and self._const_kernel == other._const_kernel)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return (f"RIN({self._width}, {self._adaptive_width}, "
                f"{self._out_dim}, {self._force_sum_one}, {self._const_kernel})")


class RDW(Preparateur):
    """Every dimension raised to a random exponent (reference:
    transform.py:571-613)."""

    def __init__(self, dist: Literal["dirichlet", "uniform"] = "dirichlet") -> None:
        self._dist = dist

    def _fit_device(self, X: torch.Tensor) -> None:
        X = X.contiguous()
        n, d, t = X.shape
        if self._dist == "dirichlet":
            # alphas = np.max(np.mean(np.abs(X), axis=0), axis=1) on the GPU, the draw on
            # the host (transform.py:592-596)
            a = be.empty((d,))
            be.check(be.lib().fb_abs_mean_max(X.data_ptr(), a.data_ptr(), n, d, t,
                                              be.stream_ptr()))
            alphas = a.cpu().numpy()
            alphas[alphas != 0] = alphas[alphas != 0] / np.max(alphas[alphas != 0])
            if np.sum(alphas == 0) >= 1:
                alphas += 1e-5
            self._weights = np.random.dirichlet(alphas)
        else:
            self._weights = np.random.random(d)
            self._weights = self._weights / np.sum(self._weights)

    def _row_independent_fit(self) -> bool:
        return self._dist != "dirichlet"      # the dirichlet parameters look at the whole sample

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_weights"):
            raise RuntimeError("Missing call of self.fit()")
        X = X.contiguous()
        n, d, t = X.shape
        if len(self._weights) != d:
            raise ValueError(f"RDW was fitted on {len(self._weights)} dimensions, got {d}")
        out = torch.empty_like(X)
        be.check(be.lib().fb_dim_pow(X.data_ptr(), _dev(self._weights, np.float64).data_ptr(),
                                     out.data_ptr(), n, d, t, be.stream_ptr()))
        return out

    def _copy(self) -> "RDW":
        return RDW(self._dist)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, RDW) and other._dist == self._dist

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"RDW({self._dist!r})"


class JLD(Preparateur):
    """Johnson-Lindenstrauss projection of the dimensions with gaussian
    weights (reference: transform.py:616-746)."""

    def __init__(self, dim: Union[int, float] = 0.99, distribute: bool = False,
                 bias: bool = False) -> None:
        if isinstance(dim, float) and not (0 < dim < 1):
            raise ValueError("'dim' has to be an integer or a float in (0, 1)")
        self._d = dim
        self._distribute = distribute
        self._bias = bias

    def _fit_device(self, X: torch.Tensor) -> None:
        d = X.shape[1]
        if isinstance(self._d, float):
            div_ = 3 * self._d**2 - 2 * self._d**3
            out_dim = int(24 * np.log(d) / div_) + 1
        else:
            out_dim = self._d
        if self._distribute:
            if out_dim > d:
                raise ValueError(
                    f"Output dimensions ({out_dim}) should be <= input dimensions ({d})")
            self._ndim_per_kernel = _split_dims(d, out_dim)
            self._dims_per_kernel = np.random.choice(d, size=d, replace=False).astype(np.int32)
        else:
            self._ndim_per_kernel = np.array(out_dim * [d], dtype=np.int32)
            self._dims_per_kernel = np.array(out_dim * list(range(d)), dtype=np.int32)
        self._kernel = np.random.standard_normal(d if self._distribute else d * out_dim)
        if self._bias:
            self._bias_weights = np.random.standard_normal(out_dim)
        else:
            self._bias_weights = np.zeros(out_dim, dtype=np.float64)

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_kernel"):
            raise RuntimeError("Missing call of self.fit()")
        X = X.contiguous()
        n, d, t = X.shape
        if int(self._dims_per_kernel.max(initial=0)) >= d:
            raise IndexError(f"JLD was fitted on more dimensions than the {d} given")
        k_d, b_d, ndim_d, dims_d = (_dev(self._kernel, np.float64),
                                    _dev(self._bias_weights, np.float64),
                                    _dev(self._ndim_per_kernel, np.int32),
                                    _dev(self._dims_per_kernel, np.int32))
        n_out = len(self._ndim_per_kernel)
        out = be.empty((n, n_out, t))
        be.check(be.lib().fb_dim_project(X.data_ptr(), k_d.data_ptr(), b_d.data_ptr(),
                                         ndim_d.data_ptr(), dims_d.data_ptr(), out.data_ptr(),
                                         n, d, t, n_out, be.stream_ptr()))
        return out

    def _copy(self) -> "JLD":
        return JLD(dim=self._d, distribute=self._distribute, bias=self._bias)

    def __eq__(self, other: Any) -> bool:
        return (isinstance(other, JLD) and self._d == other._d
                and self._distribute == other._distribute and self._bias == other._bias)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"JLD({self._d}, {self._distribute}, {self._bias})"


class SPE(Preparateur):
    """Sinusoidal positional embedding ``x_t * sin(t / T**f)`` (or ``+``), the
    position optionally measured by the L1 / L2 increment sums of the raw input
    (reference: transform.py:749-835).  A custom ``function`` is user code: it is
    called on the host with the positions, its result goes back to the GPU."""

    def __init__(self, freq: float,
                 operation: Literal["additive", "multiplicative"] = "multiplicative",
                 function: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                 step_transform: Optional[Literal["L1", "L2"]] = None,
                 max_length: Optional[int] = None) -> None:
        self._freq = freq
        self._operation = operation
        self._function = function
        self._step_transform = step_transform
        self._max_length = max_length

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        from ..cache import CacheType
        if self._operation not in ("multiplicative", "additive"):
            raise ValueError(f"Unknown operation given: {self._operation}")
        X = X.contiguous()
        n, d, t = X.shape
        use_sin = int(self._function is None)
        if self._step_transform is None:
            T = t if self._max_length is None else self._max_length
            wave = be.empty((1, t))
            be.check(be.lib().fb_spe_range(0, wave.data_ptr(), 1, t, float(T**self._freq),
                                           float(self._freq), 0, use_sin, be.stream_ptr()))
        else:
            # the sums come from the cache of the RAW batch (all its rows, whatever rows X
            # holds -- fruits/cache.py:97-112); numpy broadcasting decides below
            src = self._cache.get_device(CacheType.ISS, self._step_transform, X).contiguous()
            if src.shape[1] != t:
                raise ValueError("the cached increment sums do not match the input")
            wave = be.empty(tuple(src.shape))
            per_row = self._max_length is None       # T = the last value of every row
            be.check(be.lib().fb_spe_range(
                src.data_ptr(), wave.data_ptr(), src.shape[0], t,
                1.0 if per_row else float(self._max_length**self._freq), float(self._freq),
                int(per_row), use_sin, be.stream_ptr()))
        if self._function is not None:
            host = wave.cpu().numpy()
            wave = be.to_device(np.asarray(
                self._function(host[0] if self._step_transform is None else host),
                dtype=np.float64).reshape(wave.shape))
        rows = wave.shape[0]
        if rows != n and rows != 1 and n != 1:
            raise ValueError(f"operands could not be broadcast together with shapes "
                             f"{tuple(X.shape)} {(rows, 1, t)}")
        out = be.empty((max(n, rows) if n else 0, d, t))
        be.check(be.lib().fb_wave_embed(X.data_ptr(), wave.data_ptr(), out.data_ptr(), n, rows,
                                        d, t, int(self._operation == "additive"),
                                        be.stream_ptr()))
        return out

    def _needs_raw_cache(self) -> bool:
        return self._step_transform is not None

    def _copy(self) -> "SPE":
        return SPE(freq=self._freq, operation=self._operation, function=self._function,
                   step_transform=self._step_transform, max_length=self._max_length)

    def __eq__(self, other: Any) -> bool:
        return (isinstance(other, SPE) and self._freq == other._freq
                and self._operation == other._operation and self._function == other._function
                and self._step_transform == other._step_transform
                and self._max_length == other._max_length)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return (f"SPE({self._freq}, {self._operation}, {self._function}, "
                f"{self._step_transform}, {self._max_length})")


class RPE(Preparateur):
    """Rotational positional embedding of a two-dimensional series (reference:
    transform.py:838-907)."""

    def __init__(self, freq: float, max_length: Optional[int] = None) -> None:
        self._freq = freq
        self._max_length = max_length

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if X.shape[1] != 2:
            raise ValueError(f"RPE input has to have 2 dimensions, got {X.shape[1]}")
        X = X.contiguous()
        n, _, t = X.shape
        T = t if self._max_length is None else self._max_length
        out = torch.empty_like(X)
        be.check(be.lib().fb_rotate2(X.data_ptr(), out.data_ptr(), n, t,
                                     float(float(T)**self._freq), be.stream_ptr()))
        return out

    def _copy(self) -> "RPE":
        return RPE(freq=self._freq, max_length=self._max_length)

    def __eq__(self, other: Any) -> bool:
        return (isinstance(other, RPE) and self._freq == other._freq
                and self._max_length == other._max_length)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"RPE({self._freq}, {self._max_length})"


class CTS(Preparateur):
    """Constant time shift to the left, the last value repeated behind; or,
    with ``pseudo_shift``, the first ``s`` values set to zero (reference:
    transform.py:910-958)."""

    def __init__(self, s: Union[float, int], pseudo_shift: bool = False) -> None:
        self._s = s
        self._pseudo_shift = pseudo_shift

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        X = X.contiguous()
        n, d, t = X.shape
        shift = max(1, int(self._s * t)) if 0 < self._s < 1 else int(self._s)
        out = torch.empty_like(X)
        if self._pseudo_shift:
            # Y[:, :, :shift] = 0 with Python's slice rules for any integer
            from .filter import _keep_mask

            def build(length):
                keep = np.ones(length, dtype=np.uint8)
                keep[:shift] = 0
                return keep
            be.check(be.lib().fb_time_mask(X.data_ptr(), out.data_ptr(), n, d, t,
                                           _keep_mask(self, X, (shift,), build).data_ptr(),
                                           0, 0, 0, be.stream_ptr()))
            return out
        if shift < 1:
            # the reference's slice assignment Y[:, :, :-s] = Y[:, :, s:] only has matching
            # shapes for s >= 1
            raise ValueError("CTS needs a shift of at least one time step")
        be.check(be.lib().fb_time_shift(X.data_ptr(), out.data_ptr(), n * d, t, shift,
                                        be.stream_ptr()))
        return out

    def _copy(self) -> "CTS":
        return CTS(s=self._s, pseudo_shift=self._pseudo_shift)

    def __eq__(self, other: Any) -> bool:
        return (isinstance(other, CTS) and self._s == other._s
                and self._pseudo_shift == other._pseudo_shift)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"CTS({self._s}, {self._pseudo_shift})"


class QTC(Preparateur):
    """Quantile cut ``min(q, x_i)`` (``max`` with ``lower``), ``q`` a quantile of
    all values of the fit sample (reference: transform.py:961-1015)."""

    def __init__(self, q: float, lower: bool = False, bound: Optional[float] = None) -> None:
        self._q = q
        self._lower = lower
        self._bound = bound

    def _fit_device(self, X: torch.Tensor) -> None:
        from ..sieving.abstract import quantile_rows
        # np.quantile over the flattened sample: GPU radix select + numpy's _lerp
        self._quantile = np.float64(quantile_rows(X.contiguous().reshape(1, -1), self._q)[0])

    def _row_independent_fit(self) -> bool:
        return False

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        if not hasattr(self, "_quantile"):
            raise RuntimeError("Missing call of self.fit()")
        X = X.contiguous()
        out = torch.empty_like(X)
        bound = self._quantile if self._bound is None else self._bound
        be.check(be.lib().fb_clip_where(X.data_ptr(), out.data_ptr(), X.numel(),
                                        float(self._quantile), float(bound), int(self._lower),
                                        be.stream_ptr()))
        return out

    def _copy(self) -> "QTC":
        return QTC(q=self._q, lower=self._lower, bound=self._bound)

    def __eq__(self, other: Any) -> bool:
        return (isinstance(other, QTC) and self._q == other._q
                and self._lower == other._lower and self._bound == other._bound)

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"QTC({self._q}, {self._lower}, {self._bound})"


class FUN(Preparateur):
    """Applies a user-supplied Python function to the dataset (reference:
    transform.py:1018-1048).  The function is user code on numpy arrays, so the
    input makes one round trip through host memory; it is given the whole batch
    at once (``Fruit.transform`` does not stream row chunks through it)."""

    def __init__(self, f: Callable[[np.ndarray], np.ndarray]) -> None:
        self._function = f

    @property
    def requires_fitting(self) -> bool:
        return False

    def _row_independent_fit(self) -> bool:
        return False

    def _row_independent_transform(self) -> bool:
        return False

    def _transform_device(self, X: torch.Tensor) -> torch.Tensor:
        return be.to_device(np.asarray(self._function(X.cpu().numpy())))

    def _copy(self) -> "FUN":
        return FUN(self._function)

    def __eq__(self, other: Any) -> bool:
        return False

    __hash__ = object.__hash__

    def __str__(self) -> str:
        return f"FUN({self._function})"
