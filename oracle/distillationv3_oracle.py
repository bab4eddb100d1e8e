"""CPU restatement of the DistillationV3 loss -- TEST INFRASTRUCTURE ONLY, like dinov2_oracle.py:
only tests/ (and later smoke()/bench's CPU arm) may import it; the product path never does.

Follows LT/_methods/distillationv3/distillationv3_loss.py:35-117 (DistillationV3Loss.forward): two KL(batchmean)
terms between softmax distributions over
  * global : similarities of the (L2-normalised) global teacher / student features to a queue of teacher features,
  * local  : token-token similarities inside each image, teacher vs student,
each at its own temperature.  Pinned against the imported reference module (tests/golden/distill_v3_loss.pt:
values and gradients wrt the student features) and the reference's own property tests (zero when teacher == student,
non-negative).  The CUDA path for the distillation method (cfg4) is not built yet.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor


def _kl_batchmean(student_logits: Tensor, teacher_logits: Tensor) -> Tensor:
    """KLDivLoss(reduction="batchmean", log_target=False)(log_softmax(s), softmax(t)) with rows = all leading dims
    flattened by the caller: sum_rows sum_k p_t * (log p_t - log p_s) / rows  (distillationv3_loss.py:33, 82-84)."""
    log_ps = torch.log_softmax(student_logits, dim=-1)
    log_pt = torch.log_softmax(teacher_logits, dim=-1)
    pt = log_pt.exp()
    return (pt * (log_pt - log_ps)).sum() / student_logits.shape[0]


def distillation_v3_loss(teacher_global: Tensor, teacher_local: Tensor, student_global: Tensor, student_local: Tensor,
                         queue: Tensor, temperature_global: float, temperature_local: float) -> Tuple[Tensor, Tensor]:
    """teacher/student_global [B, D], *_local [B, M, D], queue [C, D]; all L2-normalised by the caller (:44-58)."""
    # global: similarities to the queue of teacher features (:60-84)
    s_q = student_global @ queue.t()
    t_q = teacher_global @ queue.t()
    loss_global = _kl_batchmean(s_q / temperature_global, t_q / temperature_global)
    # local: token-token similarities inside each image, rows flattened over (image, token) (:86-115)
    t_tt = torch.bmm(teacher_local, teacher_local.transpose(1, 2)).flatten(0, 1)
    s_ss = torch.bmm(student_local, student_local.transpose(1, 2)).flatten(0, 1)
    loss_local = _kl_batchmean(s_ss / temperature_local, t_tt / temperature_local)
    return loss_global, loss_local
