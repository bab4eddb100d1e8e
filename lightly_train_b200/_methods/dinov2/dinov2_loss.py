"""DINOLoss / IBOTPatchLoss on B200 kernels.

Mirrors LT/_methods/dinov2/dinov2_loss.py:61-297: same class names, constructor arguments, `center` buffers
([1,K] / [1,1,K] fp32, registered under the same names so checkpoints stay compatible) and the same call
sequence (softmax_center_teacher | sinkhorn_knopp_teacher -> update_center -> forward / forward_masked).

Difference in representation, not in arithmetic: teacher probabilities are never materialised as a [rows, K]
fp32 tensor.  The teacher methods return a `TeacherProbs` handle (bf16 logits + per-prototype `colterm` +
per-row `rowterm`) from which p[b,k] = exp(t[b,k]/T + colterm[k] + rowterm[b]); the loss methods consume the
handle with the fused CE kernel (forward value + analytic gradient in one pass).  `TeacherProbs.materialize()`
produces the dense tensor for inspection / tests.

Distributed semantics follow the reference: center batch sums and Sinkhorn prototype sums are all-reduced
(`torch.distributed`, NCCL) -- the only collectives on the loss path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor, nn

from ... import ops


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


@dataclass
class TeacherProbs:
    logits: Tensor      # bf16 [R, K]
    colterm: Tensor     # f32 [K]
    rowterm: Tensor     # f32 [R]
    t_scale: float      # 1 / teacher_temp

    def materialize(self) -> Tensor:
        z = self.logits.float() * self.t_scale + self.colterm + self.rowterm[:, None]
        return torch.exp(z)


def _as_bf16_2d(x: Tensor) -> Tensor:
    x = x.reshape(-1, x.shape[-1])
    if x.dtype != torch.bfloat16:
        y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        ops.cast_bf16(x.contiguous().float(), y)
        return y
    return x.contiguous()


def sinkhorn_colterm(logits: Tensor, t_scale: float, n_iterations: int = 3, scale_dev: Optional[Tensor] = None,
                     row_mask: Optional[Tensor] = None) -> Tensor:
    """log u of the Sinkhorn-Knopp diagonal scaling Q = diag(u) exp(t/T)^T diag(v)  (dinov2_loss.py:84-115).

    Each iteration = one weighted column reduction (prototype sums, all-reduced over ranks) + one row LSE.
    Scalar normalisations of the reference (sum_Q, /K, /B, *B) cancel in the final column normalisation, which
    the caller performs as rowterm = -LSE_k(t/T + log u).
    scale_dev: 1/T read from device memory (CUDA-graph replay).  row_mask (f32 [R], 0 for real rows, -1e30 for the padding
    rows of a static-shape replay): added to log v so that padding rows carry no mass in the prototype sums.
    """
    R, K = logits.shape
    dev = logits.device
    logv = torch.zeros(R, device=dev, dtype=torch.float32) if row_mask is None else row_mask.clone()
    logu = torch.empty(K, device=dev, dtype=torch.float32)
    sums = torch.empty(K, device=dev, dtype=torch.float32)
    for it in range(n_iterations):
        ops.fill_f32(sums, 0.0)
        ops.col_reduce(logits, sums, rowvec=logv, scale=t_scale, mode=1, scale_dev=scale_dev)
        if _world() > 1:
            dist.all_reduce(sums)
        ops.vec_op(logu, sums, 0.0, 0.0, 2)  # logu = -log(sums)
        if it + 1 < n_iterations:
            ops.row_lse(logits, logu, t_scale, logv, scale_dev=scale_dev)  # log v = -LSE_k(t/T + log u)
            if row_mask is not None:
                logv.add_(row_mask)
    return logu


class _CenteredLoss(nn.Module):
    center: Tensor

    def __init__(self, out_dim: int, student_temp: float, center_momentum: float, center_shape) -> None:
        super().__init__()
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.register_buffer("center", torch.zeros(*center_shape))
        self.updated = True
        self.reduce_handle = None
        self.async_batch_center: Optional[Tensor] = None
        self._len = 1

    # ---- centering ---------------------------------------------------------------------------
    @torch.no_grad()
    def softmax_center_teacher(self, teacher_output: Tensor, teacher_temp: float) -> TeacherProbs:
        self.apply_center_update()
        t = _as_bf16_2d(teacher_output)
        t_scale = 1.0 / teacher_temp
        colterm = torch.empty(t.shape[1], device=t.device, dtype=torch.float32)
        ops.vec_op(colterm, self.center.view(-1), t_scale, 0.0, 1)  # -center / T
        rowterm = torch.empty(t.shape[0], device=t.device, dtype=torch.float32)
        ops.row_lse(t, colterm, t_scale, rowterm)
        return TeacherProbs(t, colterm, rowterm, t_scale)

    @torch.no_grad()
    def sinkhorn_knopp_teacher(self, teacher_output: Tensor, teacher_temp: float,
                               n_masked_patches_tensor: Optional[Tensor] = None, n_iterations: int = 3) -> TeacherProbs:
        t = _as_bf16_2d(teacher_output)
        t_scale = 1.0 / teacher_temp
        logu = sinkhorn_colterm(t, t_scale, n_iterations)
        rowterm = torch.empty(t.shape[0], device=t.device, dtype=torch.float32)
        ops.row_lse(t, logu, t_scale, rowterm)
        return TeacherProbs(t, logu, rowterm, t_scale)

    @torch.no_grad()
    def update_center(self, teacher_output: Tensor) -> None:
        self.reduce_center_update(teacher_output)

    @torch.no_grad()
    def _launch_reduce(self, batch_sum: Tensor, length: int) -> None:
        self.updated = False
        self._len = length
        self.async_batch_center = batch_sum
        if _world() > 1:
            self.reduce_handle = dist.all_reduce(self.async_batch_center, async_op=True)

    @torch.no_grad()
    def apply_center_update(self) -> None:
        if self.updated is False:
            if self.reduce_handle is not None:
                self.reduce_handle.wait()
                self.reduce_handle = None
            m = self.center_momentum
            # center = center*m + (sum / (len*world)) * (1-m)      (dinov2_loss.py:148-160)
            ops.vec_op(self.center.view(-1), self.async_batch_center.view(-1), (1.0 - m) / (self._len * _world()), m, 0)
            self.updated = True


class DINOLoss(_CenteredLoss):
    def __init__(self, out_dim: int, student_temp: float = 0.1, center_momentum: float = 0.9) -> None:
        super().__init__(out_dim, student_temp, center_momentum, (1, out_dim))

    @torch.no_grad()
    def reduce_center_update(self, teacher_output: Tensor) -> None:
        t = _as_bf16_2d(teacher_output)
        s = torch.zeros(t.shape[1], device=t.device, dtype=torch.float32)
        ops.col_reduce(t, s)
        self._launch_reduce(s, t.shape[0])

    def forward(self, student_output_list: Sequence[Tensor], teacher_out_softmaxed_centered_list) -> Tensor:
        """-sum_s sum_t mean_rows(sum_k t*log_softmax(s/T_s)) (dinov2_loss.py:117-133); forward value only.
        Each teacher entry is a TeacherProbs handle (or a list of handles sharing logits storage)."""
        teachers: List[TeacherProbs] = (list(teacher_out_softmaxed_centered_list)
                                        if isinstance(teacher_out_softmaxed_centered_list, (list, tuple))
                                        else [teacher_out_softmaxed_centered_list])
        total = None
        for s in student_output_list:
            s2 = _as_bf16_2d(s)
            R = s2.shape[0]
            for tp in teachers:
                idx = torch.arange(R, device=s2.device, dtype=torch.int32)
                rows = torch.empty(R, device=s2.device, dtype=torch.float32)
                ops.dino_ce(s2, tp.logits, tp.colterm, tp.rowterm, idx, None, None, 1.0 / self.student_temp, tp.t_scale, rows)
                val = rows.sum() / R
                total = val if total is None else total + val
        return total


class IBOTPatchLoss(_CenteredLoss):
    def __init__(self, patch_out_dim: int, student_temp: float = 0.1, center_momentum: float = 0.9) -> None:
        super().__init__(patch_out_dim, student_temp, center_momentum, (1, 1, patch_out_dim))

    @torch.no_grad()
    def reduce_center_update(self, teacher_patch_tokens: Tensor) -> None:
        # sum over dim0 of mean over dim1 of [1, M, K]  (dinov2_loss.py:274-282)
        length = teacher_patch_tokens.shape[0] if teacher_patch_tokens.dim() == 3 else 1
        t = _as_bf16_2d(teacher_patch_tokens)
        per_group = t.shape[0] // length
        s = torch.zeros(t.shape[1], device=t.device, dtype=torch.float32)
        w = torch.full((t.shape[0],), 1.0 / per_group, device=t.device, dtype=torch.float32)
        ops.col_reduce(t, s, rowvec=w)
        self._launch_reduce(s, length)

    def forward_masked(self, student_patch_tokens_masked: Tensor, teacher_patch_tokens_masked: TeacherProbs,
                       student_masks_flat: Tensor, n_masked_patches: Optional[int] = None,
                       masks_weight: Optional[Tensor] = None) -> Tensor:
        """dinov2_loss.py:246-268; forward value only."""
        s2 = _as_bf16_2d(student_patch_tokens_masked)
        tp = teacher_patch_tokens_masked
        if masks_weight is None:
            masks_weight = (1 / student_masks_flat.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(student_masks_flat)[student_masks_flat]
        R = s2.shape[0] if n_masked_patches is None else n_masked_patches
        idx = torch.arange(R, device=s2.device, dtype=torch.int32)
        rows = torch.empty(R, device=s2.device, dtype=torch.float32)
        ops.dino_ce(s2[:R], tp.logits, tp.colterm, tp.rowterm, idx, None, masks_weight.float().contiguous(),
                    1.0 / self.student_temp, tp.t_scale, rows)
        return rows.sum() / student_masks_flat.shape[0]
