"""DistillationV2Loss: MSE between un-normalised teacher and student token features.

Mirror of LT/_methods/distillationv2/distillationv2_loss.py:12-44.  One fused kernel (`b200_mse`): value and the gradient
wrt the student features in a single pass over both tensors (the reference's MSELoss + autograd reads them twice)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from ... import ops


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher: Tensor, student: Tensor) -> Tensor:  # type: ignore[override]
        t = teacher.reshape(-1).float().contiguous()
        s = student.reshape(-1).float().contiguous()
        out = torch.zeros(1, device=s.device, dtype=torch.float32)
        ds = torch.empty_like(s)
        ops.mse(t, s, out, ds)
        ctx.save_for_backward(ds)
        ctx.shape, ctx.dtype = student.shape, student.dtype
        return out[0]

    @staticmethod
    def backward(ctx, g: Tensor):  # type: ignore[override]
        (ds,) = ctx.saved_tensors
        return None, (ds * g).reshape(ctx.shape).to(ctx.dtype)


class DistillationV2Loss(nn.Module):
    def forward(self, teacher_features: Tensor, student_features: Tensor) -> Tensor:
        """teacher / student [B, n_features, D]; mean over all elements of (teacher - student)^2  (:27-44)."""
        return _MSE.apply(teacher_features, student_features)
