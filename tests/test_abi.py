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
