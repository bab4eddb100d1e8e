"""ctypes signatures of the C-ABI, parsed from include/b200dino.h (single source of truth)."""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

HEADER = Path(__file__).resolve().parents[1] / "include" / "b200dino.h"

_SCALARS = {"int": C.c_int, "long long": C.c_longlong, "float": C.c_float}


def _ctype(decl: str):
    decl = decl.strip()
    if "*" in decl:
        return C.c_void_p
    # strip the parameter name
    parts = decl.split()
    base = " ".join(parts[:-1]) if len(parts) > 1 else parts[0]
    base = base.replace("const ", "").strip()
    if base in _SCALARS:
        return _SCALARS[base]
    raise ValueError(f"unknown C type in header: {decl!r}")


def parse_header(path: Path = HEADER) -> dict[str, list]:
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    sigs: dict[str, list] = {}
    for m in re.finditer(r"\bint\s+(b200_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2)
        params = " ".join(params.split())
        sigs[name] = [] if params in ("", "void") else [_ctype(p) for p in params.split(",")]
    return sigs


SIGNATURES = parse_header()
