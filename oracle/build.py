"""Build and load the C part of the oracle (test infrastructure only).

``build_oracle()`` compiles ``fruits_oracle.c`` with gcc into
``oracle/_build/libfruits_oracle.so``.  ``-ffp-contract=off`` keeps the real
semiring free of FMA contraction (the reference rounds after every array
statement); the arctic kernels call ``fma()`` explicitly.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "fruits_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "libfruits_oracle.so")

_lib = None


def build_oracle(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(_OUT)
            and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC)):
        return _OUT
    cmd = [
        "gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-mavx2", "-mfma",
        "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden",
        "-o", _OUT, _SRC, "-lm",
    ]
    subprocess.run(cmd, check=True)
    return _OUT


def load_oracle() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = build_oracle()
    lib = ctypes.CDLL(path)
    i64, i32, dbl = ctypes.c_int64, ctypes.c_int, ctypes.c_double
    ptr = ctypes.c_void_p
    lib.fo_increments.argtypes = [ptr, ptr, i64, i64, i64, i64]
    lib.fo_increments.restype = None
    lib.fo_lsum.argtypes = [ptr, ptr, i64, i64, i64, i32]
    lib.fo_lsum.restype = None
    lib.fo_coquantile.argtypes = [ptr, ptr, i64, i64, dbl]
    lib.fo_coquantile.restype = None
    lib.fo_iterated_sums.argtypes = [ptr, ptr, ptr, ptr, ptr, i64, i64, i64,
                                     i64, i64, i64, i32, i32]
    lib.fo_iterated_sums.restype = None
    lib.fo_segment_sieve.argtypes = [ptr, ptr, ptr, ptr, i64, i64, i64, i64,
                                     i32]
    lib.fo_segment_sieve.restype = None
    lib.fo_ppv.argtypes = [ptr, ptr, ptr, i64, i64, i64, i32]
    lib.fo_ppv.restype = None
    lib.fo_num_threads.argtypes = []
    lib.fo_num_threads.restype = i32
    _lib = lib
    return lib
