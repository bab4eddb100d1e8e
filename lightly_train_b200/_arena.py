"""Flat parameter arenas.

All parameters of one "side" (student or teacher: backbone + heads) live in ONE contiguous fp32 buffer, every
parameter padded to a multiple of CHUNK elements.  nn.Parameters handed to the mirror modules are views into
it, so state_dict()/load_state_dict() see the reference's names and shapes while
  * the optimizer / EMA / gradient all-reduce / grad-norm each touch memory in a single coalesced sweep, and
  * per-parameter hyper-parameters (layer-wise lr decay, weight-decay mask, freeze flags) become small
    per-chunk tables (LT/_methods/dinov2/utils.py:191-273 as data instead of ~25 optimizer groups).
A bf16 shadow of the arena feeds the tcgen05 GEMMs and is refreshed by the fused optimizer sweep.
"""
from __future__ import annotations

from typing import Dict, Iterable, Tuple

import torch

from . import ops

CHUNK = 1024


class Arena:
    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device, with_grad: bool, with_optim_state: bool):
        self.shapes = dict(shapes)
        self.offsets: Dict[str, Tuple[int, int]] = {}
        off = 0
        for name, shape in shapes.items():
            n = 1
            for s in shape:
                n *= s
            self.offsets[name] = (off, n)
            off += (n + CHUNK - 1) // CHUNK * CHUNK
        self.total = off
        self.n_chunks = off // CHUNK
        self.device = device
        self.fp32 = torch.zeros(off, device=device, dtype=torch.float32)
        self.bf16 = torch.zeros(off, device=device, dtype=torch.bfloat16)
        self.grad = torch.zeros(off, device=device, dtype=torch.float32) if with_grad else None
        self.exp_avg = torch.zeros(off, device=device, dtype=torch.float32) if with_optim_state else None
        self.exp_avg_sq = torch.zeros(off, device=device, dtype=torch.float32) if with_optim_state else None
        self.bf16_valid = False

    # ------------------------------------------------------------------ views
    def _v(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        off, n = self.offsets[name]
        return buf[off:off + n].view(self.shapes[name])

    def p(self, name: str) -> torch.Tensor:
        return self._v(self.fp32, name)

    def w(self, name: str) -> torch.Tensor:
        """bf16 shadow view (GEMM operand)."""
        return self._v(self.bf16, name)

    def g(self, name: str) -> torch.Tensor:
        return self._v(self.grad, name)

    def names(self) -> Iterable[str]:
        return self.offsets.keys()

    # ------------------------------------------------------------------ maintenance
    def refresh_bf16(self) -> None:
        ops.cast_bf16(self.fp32, self.bf16)
        self.bf16_valid = True

    def zero_grad(self) -> None:
        ops.fill_f32(self.grad, 0.0)

    def chunk_table(self, per_param: Dict[str, float], dtype=torch.float32) -> torch.Tensor:
        """Expand a per-parameter value to the per-chunk table the sweep kernels index."""
        t = torch.zeros(self.n_chunks, dtype=dtype)
        for name, (off, n) in self.offsets.items():
            c0, c1 = off // CHUNK, (off + n + CHUNK - 1) // CHUNK
            t[c0:c1] = per_param[name]
        return t.to(self.device)

    def load_from(self, state: Dict[str, torch.Tensor]) -> None:
        for name in self.offsets:
            self.p(name).copy_(state[name].to(self.device, torch.float32).view(self.shapes[name]))
        self.bf16_valid = False
