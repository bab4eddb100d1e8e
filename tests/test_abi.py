"""CPU checks of the drop-in boundary: libb200dino.so loads without a GPU and exports every symbol that
include/b200dino.h declares (no compute calls here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from lightly_train_b200 import _build, _lib
    _build.build()
    return _lib.lib()


def test_library_loads_and_reports_version(lib):
    assert b"sm_100a" in lib.b200_version()


def test_every_declared_symbol_is_exported(lib):
    text = (ROOT / "include" / "b200dino.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(b200_\w+)\s*\(", text))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in b200dino.h but not exported"


def test_header_signatures_parse():
    from lightly_train_b200 import _sigs
    assert _sigs.SIGNATURES["b200_gemm"] == [ctypes.c_void_p, ctypes.c_void_p]
    assert len(_sigs.SIGNATURES["b200_layernorm_fwd"]) == 13


def test_ops_refuse_cpu_tensors():
    """The product path has no CPU fallback: CPU tensors must raise, not silently compute."""
    import torch
    from lightly_train_b200 import ops
    from lightly_train_b200._lib import B200Error
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(B200Error):
        ops.gemm(a, a, torch.zeros(8, 8, dtype=torch.bfloat16))


def test_product_package_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under lightly_train_b200/ may import, call or mention it."""
    import pathlib
    root = pathlib.Path(__file__).resolve().parents[1] / "lightly_train_b200"
    offenders = [str(p) for p in root.rglob("*.py") if "oracle" in p.read_text()]
    offenders += [str(p) for p in (root / "csrc").glob("*.cu*") if "oracle" in p.read_text()]
    assert not offenders, offenders


def test_error_convention_on_invalid_arguments(lib):
    """entry points return a negative int on bad arguments and never throw, exit or touch the device --
    these calls are rejected by the argument checks before any CUDA API is reached, so they run without a GPU."""
    from lightly_train_b200._lib import GemmArgs
    INVALID, UNSUPPORTED = -1, -2
    g = GemmArgs()  # all-zero: null operands
    assert lib.b200_gemm(ctypes.byref(g), None) == INVALID
    dummy = ctypes.c_void_p(0x1000)  # never dereferenced: the shape checks fail first
    g.A = g.B = g.C = 0x1000
    g.M, g.N, g.K, g.lda, g.ldb, g.ldc = 128, 100, 64, 64, 64, 100  # N not a multiple of 8
    assert lib.b200_gemm(ctypes.byref(g), None) == UNSUPPORTED
    g.N, g.ldc, g.epi = 128, 128, 99  # unknown epilogue
    assert lib.b200_gemm(ctypes.byref(g), None) == INVALID
    # attention kernels are specialised for head_dim 64; sequences above 272 tokens are not supported
    assert lib.b200_attention_fwd(dummy, 192, 1, 16, 1, 32, 0.125, dummy, 64, None, None) == UNSUPPORTED
    assert lib.b200_attention_fwd(dummy, 192, 1, 300, 1, 64, 0.125, dummy, 64, None, None) == UNSUPPORTED
    assert lib.b200_attention_fwd(dummy, 192, 1, 16, 1, 64, -1.0, dummy, 64, None, None) == INVALID
    # SwiGLU gate: H must be a multiple of 8
    assert lib.b200_swiglu_fwd(dummy, 24, 4, 12, dummy, 12, None) == UNSUPPORTED
    assert lib.b200_swiglu_fwd(None, 16, 4, 8, dummy, 8, None) == INVALID
