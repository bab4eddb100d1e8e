"""DistillationV3Loss on B200 kernels.

Mirror of LT/_methods/distillationv3/distillationv3_loss.py:16-117 (same class name, constructor and forward arguments,
returns (global_loss, local_loss)).  The two softmax / log_softmax / KLDivLoss(batchmean) chains are ONE fused row
kernel each (`b200_kl_rows`: forward value + analytic gradient wrt the student logits in a single pass), wrapped in a
torch.autograd.Function because the student (a torchvision ResNet on cuDNN) trains under torch autograd; the similarity
matmuls stay torch ops for the same reason (the gradient has to flow into the student features).
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor, nn

from ... import ops


class _KLRows(torch.autograd.Function):
    """mean over rows of KL(softmax(t / T) || softmax(s / T)); gradient wrt s from the same kernel pass."""

    @staticmethod
    def forward(ctx, s: Tensor, t: Tensor, inv_temp: float) -> Tensor:  # type: ignore[override]
        s2 = s.reshape(-1, s.shape[-1]).float().contiguous()
        t2 = t.reshape(-1, t.shape[-1]).float().contiguous()
        rows = torch.empty(s2.shape[0], device=s2.device, dtype=torch.float32)
        ds = torch.empty_like(s2)
        ops.kl_rows(s2, t2, inv_temp, rows, ds, gscale=1.0 / s2.shape[0])
        ctx.save_for_backward(ds)
        ctx.shape, ctx.dtype = s.shape, s.dtype
        return rows.sum() / s2.shape[0]

    @staticmethod
    def backward(ctx, g: Tensor):  # type: ignore[override]
        (ds,) = ctx.saved_tensors
        return (ds * g).reshape(ctx.shape).to(ctx.dtype), None, None


class DistillationV3Loss(nn.Module):
    def __init__(self, temperature_global: float, temperature_local: float) -> None:
        super().__init__()
        self.temperature_global = temperature_global
        self.temperature_local = temperature_local

    def forward(self, teacher_features_global: Tensor, teacher_features_local: Tensor, student_features_global: Tensor,
                student_features_local: Tensor, queue: Tensor) -> Tuple[Tensor, Tensor]:
        """All inputs L2-normalised: *_global [B, D], *_local [B, M, D], queue [C, D]  (:35-58)."""
        s_q = torch.einsum("b d, c d -> b c", student_features_global, queue)  # :60-67
        t_q = torch.einsum("b d, c d -> b c", teacher_features_global, queue)
        global_loss = _KLRows.apply(s_q, t_q, 1.0 / self.temperature_global)  # :69-84
        t_tt = torch.einsum("b m d, b n d -> b m n", teacher_features_local, teacher_features_local).flatten(0, 1)  # :86-92
        s_ss = torch.einsum("b m d, b n d -> b m n", student_features_local, student_features_local).flatten(0, 1)  # :94-100
        local_loss = _KLRows.apply(s_ss, t_tt, 1.0 / self.temperature_local)  # :102-115
        return global_loss, local_loss
