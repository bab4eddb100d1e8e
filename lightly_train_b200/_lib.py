"""ctypes loader for libb200dino.so (the C-ABI in include/b200dino.h).

The library is the product path: if it is missing or fails to load this module raises -- there is no CPU or
PyTorch fallback.  Call `lightly_train_b200._build.build()` (or `python -m lightly_train_b200._build`) first.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libb200dino.so"

ERRORS = {
    -1: "B200_ERR_INVALID_ARG",
    -2: "B200_ERR_UNSUPPORTED",
    -3: "B200_ERR_CUDA",
    -4: "B200_ERR_DRIVER",
}

EPI_BF16, EPI_F32, EPI_F32_ATOMIC, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_DGELU, EPI_BIAS_GELU_DG, EPI_MUL_AUX = range(8)


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_longlong), ("a_mn", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_mn", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("splits", C.c_int), ("epi", C.c_int), ("block_n", C.c_int), ("ws_mode", C.c_int),
        ("alpha", C.c_float),
        ("C", C.c_void_p), ("ldc", C.c_longlong),
        ("C2", C.c_void_p), ("ldc2", C.c_longlong),
        ("aux", C.c_void_p), ("ldaux", C.c_longlong),
        ("bias", C.c_void_p), ("gamma", C.c_void_p),
        ("rowscale", C.c_void_p), ("rows_per_scale", C.c_int),
    ]


class LnArgs(C.Structure):
    """b200_ln_args (include/b200dino.h)."""
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_longlong),
        ("weight", C.c_void_p), ("bias", C.c_void_p), ("eps", C.c_float),
        ("xn_out", C.c_void_p), ("ld_xn", C.c_longlong),
        ("mean", C.c_void_p), ("rstd", C.c_void_p),
    ]


class B200Error(RuntimeError):
    pass


_lib = None
LAUNCHES = 0  # number of C-ABI kernel launches issued (bench.py reports it as gpu_launches)


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise B200Error(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no fallback path."
            )
        _lib = C.CDLL(str(LIB_PATH))
        _lib.b200_version.restype = C.c_char_p
        _declare(_lib)
    return _lib


def _declare(L: C.CDLL) -> None:
    from . import _sigs

    for name, argtypes in _sigs.SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int


def check(rc: int, what: str) -> None:
    global LAUNCHES
    LAUNCHES += 1
    if rc != 0:
        code = rc
        extra = ""
        if rc <= -3 - 16:
            code = -3
            extra = f" (cudaError_t={(-(rc + 3)) // 16})"
        raise B200Error(f"{what} failed: {ERRORS.get(code, rc)}{extra}")
