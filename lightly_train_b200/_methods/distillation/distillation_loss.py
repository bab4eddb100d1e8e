"""DistillationLoss (distillation v1) on the fused KL row kernel.

Mirror of LT/_methods/distillation/distillation_loss.py:14-75: KLDivLoss(batchmean) between the softmax distributions of the
student's and the teacher's similarities to a queue of teacher features, at one temperature -- the global term of
DistillationV3Loss, so it shares `b200_kl_rows` (forward value + gradient wrt the student logits in one pass)."""
from __future__ import annotations

import torch
from torch import Tensor, nn

from ..distillationv3.distillationv3_loss import _KLRows


class DistillationLoss(nn.Module):
    def __init__(self, temperature: float) -> None:
        super().__init__()
        self.temperature = temperature

    def forward(self, teacher_features: Tensor, student_features: Tensor, queue: Tensor) -> Tensor:
        """All inputs L2-normalised: teacher / student [B, D], queue [C, D]  (:33-75)."""
        s_q = torch.einsum("b d, c d -> b c", student_features, queue)
        t_q = torch.einsum("b d, c d -> b c", teacher_features, queue)
        return _KLRows.apply(s_q, t_q, 1.0 / self.temperature)
