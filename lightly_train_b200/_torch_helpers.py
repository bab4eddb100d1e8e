"""update_momentum with the reference signature (LT/_torch_helpers.py:75-96) on arena-backed modules.

The reference walks both parameter lists with two foreach passes.  Here the parameters of a mirror module are views
into one flat fp32 arena per side, so the EMA of a whole module is ONE coalesced sweep (`b200_ema`) over the arena
span that the module's parameters cover (padding between parameters is zero on both sides and stays zero), which also
refreshes the teacher's bf16 GEMM shadow.  Modules that are not arena-backed fall back to one launch per parameter.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn

from . import ops
from ._arena import Arena


def _arena_of(module: nn.Module) -> Optional[Arena]:
    for m in module.modules():
        a = getattr(m, "arena", None)
        if isinstance(a, Arena):
            return a
    return None


def _span(params: List[torch.Tensor], arena: Arena) -> Optional[Tuple[int, int]]:
    base = arena.fp32.data_ptr()
    lo, hi = None, None
    for p in params:
        off = (p.data_ptr() - base) // 4
        if off < 0 or off + p.numel() > arena.total:
            return None
        lo = off if lo is None else min(lo, off)
        hi = off + p.numel() if hi is None else max(hi, off + p.numel())
    if lo is None:
        return None
    lo = lo // 4 * 4  # 16-byte aligned sweep
    return lo, hi


@torch.no_grad()
def update_momentum(model: nn.Module, model_ema: nn.Module, m: float) -> None:
    """Updates parameters of `model_ema` with the EMA of `model`: p_ema = p_ema * m + p * (1 - m)."""
    ps, pt = list(model.parameters()), list(model_ema.parameters())
    if len(ps) != len(pt):
        raise ValueError(f"update_momentum: parameter lists differ in length ({len(ps)} vs {len(pt)})")
    if not pt:
        return
    sa, ta = _arena_of(model), _arena_of(model_ema)
    if sa is not None and ta is not None and sa.total == ta.total:
        s_span, t_span = _span(ps, sa), _span(pt, ta)
        if s_span is not None and s_span == t_span:
            lo, hi = s_span
            ops.ema(ta.fp32[lo:hi], sa.fp32[lo:hi], m, ta.bf16[lo:hi])
            return
    for a, b in zip(ps, pt):
        ops.ema(b.view(-1), a.detach().view(-1), m)
    if ta is not None:
        ta.bf16_valid = False
