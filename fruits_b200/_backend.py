"""ctypes binding of ``libfruits_b200.so`` (C ABI in ``include/fruits_b200.h``).

PyTorch is used for plumbing only: device memory (``torch.empty(...,
device="cuda")``), host<->device copies and streams.  Every numeric step of the
hot path runs in the hand-written CUDA library; if the library or a CUDA
device is missing the calls raise -- there is no CPU fallback.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfruits_b200.so")

FB_MAX_ROWS = 16
FB_MAX_USED_DIMS = 7
FB_RING = 256
FB_MAX_ALPHAS = 4
FB_MAX_FEATS = 16
FB_NTHR = 16

SEMIRING_REALS, SEMIRING_ARCTIC, SEMIRING_BAYESIAN = 0, 1, 2
WEIGHT_NONE, WEIGHT_TOTAL, WEIGHT_NONTOTAL = 0, 1, 2
FEAT_CNT, FEAT_AVG, FEAT_PPV, FEAT_MAX, FEAT_MIN, FEAT_END = range(6)
FEAT_XPI, FEAT_LPI, FEAT_CUR, FEAT_CPV = range(6, 10)
SIEVE_NPI, SIEVE_MPI, SIEVE_MAX, SIEVE_MIN, SIEVE_XPI, SIEVE_LPI, SIEVE_END, SIEVE_CUR = range(8)
POLICY_MAT = 0

# numpy view of `struct fb_slot` (16 bytes)
SLOT_DTYPE = np.dtype([
    ("letter_lo", "<u4"), ("letter_hi", "<u4"), ("parent", "<i2"),
    ("emit", "<i2"), ("depth", "u1"), ("aidx", "u1"), ("weight", "u1"),
    ("flags", "u1"),
])
assert SLOT_DTYPE.itemsize == 16


class FbDim(ctypes.Structure):
    _fields_ = [("raw_dim", ctypes.c_int32), ("inc", ctypes.c_int32),
                ("std", ctypes.c_int32), ("pad", ctypes.c_int32)]


class FbIssPlan(ctypes.Structure):
    _fields_ = [
        ("semiring", ctypes.c_int32), ("weight_mode", ctypes.c_int32),
        ("n_blocks", ctypes.c_int32), ("n_rows", ctypes.c_int32),
        ("n_emit", ctypes.c_int32), ("n_used_dims", ctypes.c_int32),
        ("n_alphas", ctypes.c_int32), ("max_depth", ctypes.c_int32),
        ("alphas", ctypes.c_float * FB_MAX_ALPHAS),
        ("dims", FbDim * FB_MAX_USED_DIMS),
        ("slots", ctypes.c_void_p), ("row_pub", ctypes.c_void_p),
        ("row_weight", ctypes.c_void_p),
    ]


class FbBatch(ctypes.Structure):
    _fields_ = [("X", ctypes.c_void_p), ("n", ctypes.c_int64),
                ("d", ctypes.c_int64), ("t", ctypes.c_int64),
                ("g", ctypes.c_void_p), ("g_ld", ctypes.c_int64),
                ("stats", ctypes.c_void_p)]


class FbSievePlan(ctypes.Structure):
    _fields_ = [("n_feats", ctypes.c_int32),
                ("kind", ctypes.c_int32 * FB_MAX_FEATS),
                ("arg", ctypes.c_int32 * FB_MAX_FEATS),
                ("thresholds", ctypes.c_void_p)]


_lib = None


def lib() -> ctypes.CDLL:
    """The CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m fruits_b200.build` "
            "(there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    ip = ctypes.POINTER(ctypes.c_int)
    sig = {
        "fb_abi_version": ([], i32),
        "fb_last_error": ([], ctypes.c_char_p),
        "fb_device_info": ([ip, ip, ip, ip], i32),
        "fb_slice_policy": ([ctypes.POINTER(FbSievePlan), i32, i32], i32),
        "fb_slice_rows": ([i32], i32),
        "fb_slice_features_ex": ([ctypes.POINTER(FbIssPlan), ctypes.POINTER(FbBatch),
                                  ctypes.POINTER(FbSievePlan), vp, i64, i64, i32, i32, vp], i32),
        "fb_slice_features": ([ctypes.POINTER(FbIssPlan), ctypes.POINTER(FbBatch),
                               ctypes.POINTER(FbSievePlan), vp, i64, i64, vp], i32),
        "fb_iss_materialize": ([ctypes.POINTER(FbIssPlan), ctypes.POINTER(FbBatch), vp, vp], i32),
        "fb_increments": ([vp, vp, vp, i64, i64, i64, vp], i32),
        "fb_row_stats": ([vp, vp, i64, i64, i32, dbl, vp], i32),
        "fb_standardize": ([vp, vp, vp, i64, i64, vp], i32),
        "fb_lsum": ([vp, vp, i64, i64, i64, i32, vp], i32),
        "fb_nrm_scale": ([vp, vp, i64, i64, i32, dbl, vp], i32),
        "fb_coquantile": ([vp, vp, i64, i64, dbl, vp], i32),
        "fb_pretransform": ([vp, vp, i64, i64, i32, vp], i32),
        "fb_segment_sieve": ([vp, i64, vp, i32, vp, i32, i32, vp, i64, i64, i64, i64, vp], i32),
        "fb_ppv": ([vp, i64, vp, i32, i32, vp, i64, i64, i64, i64, vp], i32),
        "fb_nan_to_num": ([vp, i64, vp], i32),
        "fb_multimem_copy": ([vp, i64, vp, i64, i64, i64, vp], i32),
        "fb_order_stats_workspace": ([i64], i64),
        "fb_order_stats": ([vp, i64, i64, i64, i64, vp, vp, vp, vp], i32),
        "fb_order_stats_multi_workspace": ([i64, i32], i64),
        "fb_order_stats_multi": ([vp, i64, i64, i64, i64, i32, ctypes.POINTER(ctypes.c_int32),
                                  ctypes.POINTER(ctypes.c_int64), vp, vp, vp, vp, vp], i32),
        "fb_order_stats_dist_layout": ([i64, i32, ctypes.POINTER(ctypes.c_int64)], i32),
        "fb_order_stats_dist": ([i32, vp, i64, i64, i64, i64, i64, i32,
                                 ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int64),
                                 vp, vp, vp, vp, vp], i32),
        "fb_order_stats_dist8_layout": ([i64, ctypes.POINTER(ctypes.c_int64)], i32),
        "fb_order_stats_dist8": ([i32, vp, i64, i64, i64, i64, i64, vp, vp, vp, vp], i32),
        "fb_fp64_peak": ([vp, i32, i32, vp], i32),
        "fb_jit_compile": ([ctypes.c_char_p, ctypes.c_char_p, i32, i32, ctypes.POINTER(vp),
                            ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t], i32),
        "fb_jit_link": ([ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t), i32,
                         ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p,
                         ctypes.c_size_t], i32),
        "fb_jit_free": ([vp], None),
        "fb_jit_load": ([ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(vp)], i32),
        "fb_jit_unload": ([vp], i32),
        "fb_jit_slice_features": ([vp, vp, ctypes.POINTER(FbBatch), vp, i64, vp, i64, vp, i64, i64,
                                   i32, vp], i32),
        "fb_jit_slice_features_cut": ([vp, vp, ctypes.POINTER(FbBatch), vp, i64, vp, i64, vp, vp,
                                       i64, i64, i32, vp], i32),
        "fb_jit_chain_features": ([vp, vp, ctypes.POINTER(FbBatch), vp, i64, vp, i64, i64, i32,
                                   vp], i32),
        "fb_cos_trig": ([vp, i32, i64, vp, vp], i32),
        "fb_cos_rows": ([vp, i32, i64, vp, i32, vp, vp], i32),
        "fb_coswiss_sep_word": ([vp, i64, i64, i64, vp, i32, i32, vp, vp, i32, i32, i32, vp, vp], i32),
        "fb_coswiss_word": ([vp, i64, i64, i64, vp, i32, i32, i32, i32, vp, i32, vp, i32, i32, vp, vp], i32),
        "fb_bayes_word": ([vp, i64, i64, i64, vp, i32, i32, vp, vp, i64, i32, i32, vp, vp], i32),
        "fb_arctic_word": ([vp, i64, i64, i64, vp, i32, i32, vp, vp, i64, i32, i32, vp, vp], i32),
        "fb_exp_rows": ([vp, vp, i64, i64, ctypes.POINTER(ctypes.c_float), i32, vp], i32),
        "fb_arctic_argmax_workspace": ([i64, i64, i32], i64),
        "fb_arctic_argmax_word": ([vp, i64, i64, i64, vp, i32, i32, vp, vp, i64, vp, vp, vp], i32),
        # csrc/prep_more.cu: the preparateurs beside INC / STD / NRM
        "fb_time_mask": ([vp, vp, i64, i64, i64, vp, vp, vp, i64, vp], i32),
        "fb_time_shift": ([vp, vp, i64, i64, i64, vp], i32),
        "fb_lead_lag": ([vp, vp, i64, i64, vp], i32),
        "fb_moving_average": ([vp, vp, i64, i64, i64, vp], i32),
        "fb_random_increments": ([vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp], i32),
        "fb_dim_project": ([vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, vp], i32),
        "fb_ffn": ([vp, vp, vp, vp, vp, vp, i64, i64, i64, i32, i32, i32, vp], i32),
        "fb_dim_pow": ([vp, vp, vp, i64, i64, i64, vp], i32),
        "fb_abs_mean_max": ([vp, vp, i64, i64, i64, vp], i32),
        "fb_rotate2": ([vp, vp, i64, i64, dbl, vp], i32),
        "fb_spe_range": ([vp, vp, i64, i64, dbl, dbl, i32, i32, vp], i32),
        "fb_wave_embed": ([vp, vp, vp, i64, i64, i64, i64, i32, vp], i32),
        "fb_clip_where": ([vp, vp, i64, dbl, dbl, i32, vp], i32),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    if L.fb_abi_version() != 1:
        raise RuntimeError("libfruits_b200.so ABI version mismatch")
    _lib = L
    return L


EXPORTED = [
    "fb_abi_version", "fb_last_error", "fb_device_info", "fb_slice_policy",
    "fb_slice_rows", "fb_slice_features_ex", "fb_slice_features",
    "fb_iss_materialize", "fb_increments", "fb_row_stats", "fb_standardize",
    "fb_lsum", "fb_nrm_scale", "fb_coquantile", "fb_pretransform",
    "fb_segment_sieve", "fb_ppv", "fb_nan_to_num", "fb_multimem_copy", "fb_order_stats_workspace",
    "fb_order_stats", "fb_fp64_peak", "fb_jit_compile", "fb_jit_free", "fb_jit_load",
    "fb_jit_unload", "fb_jit_slice_features", "fb_jit_slice_features_cut", "fb_jit_chain_features", "fb_jit_link", "fb_exp_rows", "fb_cos_trig", "fb_cos_rows", "fb_coswiss_sep_word", "fb_coswiss_word",
    "fb_bayes_word", "fb_arctic_word", "fb_order_stats_multi_workspace", "fb_order_stats_multi",
    "fb_order_stats_dist_layout", "fb_order_stats_dist", "fb_order_stats_dist8_layout",
    "fb_order_stats_dist8",
    "fb_arctic_argmax_workspace", "fb_arctic_argmax_word",
    "fb_time_mask", "fb_time_shift", "fb_lead_lag", "fb_moving_average",
    "fb_random_increments", "fb_dim_project", "fb_ffn", "fb_dim_pow", "fb_abs_mean_max",
    "fb_rotate2", "fb_spe_range", "fb_wave_embed", "fb_clip_where",
]


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().fb_last_error().decode(errors="replace")
        if rc == -2:
            raise NotImplementedError(msg)
        if rc == -1:
            raise ValueError(msg)
        raise RuntimeError(f"CUDA error {rc}: {msg}")


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "fruits_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """``cudaStream_t`` of torch's current stream on the current device."""
    if _raw_stream is not None:         # (no Stream object per call: this sits on every launch)
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def to_device(a, dtype=torch.float64) -> torch.Tensor:
    """numpy / torch input -> contiguous CUDA tensor (no copy if already one)."""
    dev = require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.asarray(a)
        if dtype == torch.float64 and a.dtype != np.float64:
            # the reference's numba signatures only accept float64
            # (fruits/cache.py:8, fruits/iss/semiring.py:169)
            raise TypeError(f"input must be float64, got {a.dtype}")
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype != dtype:
        raise TypeError(f"input must be {dtype}, got {t.dtype}")
    return t.to(dev, non_blocking=True).contiguous()


def empty(shape, dtype=torch.float64) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device=require_cuda())


def zeros(shape, dtype=torch.float64) -> torch.Tensor:
    return torch.zeros(shape, dtype=dtype, device=require_cuda())


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()
