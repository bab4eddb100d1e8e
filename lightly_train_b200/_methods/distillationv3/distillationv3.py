"""DistillationV3 method (BASELINE cfg4: DINOv3 ViT teacher -> convolutional / any student) on the B200 path.

Mirror of LT/_methods/distillationv3/distillationv3.py:168-374: same class / argument names, `teacher_queue` buffer,
`student_projection_head_{global,local}` parameters, `training_step_impl`, `_forward_teacher`, `_forward_student`,
`_update_queue`, `_mixup_data`, the `train_loss/{local,global}_loss` log keys.

What runs where:
  * teacher (frozen DINOv3 ViT, the dominant FLOPs of the step): hand-written sm_100a kernels through libb200dino.so
    (lightly_train_b200/_models/dinov3_vit.py: tcgen05 GEMMs with fused epilogues, RoPE kernel, attention kernels);
  * loss: the two KL terms as fused forward+gradient row kernels (distillationv3_loss.py);
  * student backbone (e.g. torchvision ResNet-50), the two linear projection heads, the bilinear resize and the
    L2 normalisations: torch ops under torch.autograd / bf16 autocast (cuDNN / cuBLAS), because the student's backward is
    torch's -- stated as such in DESIGN.md; only the teacher and the loss are this package's kernels.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .distillationv3_loss import DistillationV3Loss


@dataclass
class DistillationV3Args:
    """LT/_methods/distillationv3/distillationv3.py:85-110 (same field names and defaults; queue_size "auto" resolves to 8192,
    the largest bucket of resolve_auto :118-137)."""

    queue_size: int = 8192
    temperature_global: float = 0.07
    temperature_local: float = 0.07
    teacher: Any = "dinov3/vitb16"
    teacher_args: Optional[Dict[str, Any]] = None
    lr_scale_method: str = "sqrt"
    reference_batch_size: int = 1536
    loss_local_weight: float = 1.0


@dataclass
class TrainingStepResult:
    loss: Tensor
    log_dict: Dict[str, Any] = field(default_factory=dict)


class DistillationV3(nn.Module):
    def __init__(self, method_args: DistillationV3Args, optimizer_args: Any, embedding_model: nn.Module,
                 global_batch_size: int, num_input_channels: int = 3, *, teacher_embedding_model: Optional[nn.Module] = None) -> None:
        """embedding_model: the student -- an object with `.wrapped_model.forward_features / forward_pool` and `.embed_dim`
        (reference EmbeddingModel surface).  teacher_embedding_model: a DINOv3ViTModelWrapper of this package (what
        get_teacher(...) resolves "dinov3/vitb16" to in the reference, :49-82)."""
        super().__init__()
        if teacher_embedding_model is None:
            raise ValueError("pass teacher_embedding_model (lightly_train_b200._models.dinov3_vit.DINOv3ViTModelWrapper)")
        self.method_args = method_args
        self.optimizer_args = optimizer_args
        self.global_batch_size = global_batch_size
        self.teacher_embedding_model = teacher_embedding_model
        for p in self.teacher_embedding_model.parameters():
            p.requires_grad_(False)
        self.student_embedding_model = embedding_model
        self.flatten = nn.Flatten(start_dim=1)
        self.teacher_embedding_dim: int = self.teacher_embedding_model.feature_dim()
        self.student_projection_head_global = nn.Linear(embedding_model.embed_dim, self.teacher_embedding_dim)
        self.student_projection_head_local = nn.Linear(embedding_model.embed_dim, self.teacher_embedding_dim)
        nn.init.trunc_normal_(self.student_projection_head_global.weight, std=0.02)
        nn.init.trunc_normal_(self.student_projection_head_local.weight, std=0.02)
        self.criterion = DistillationV3Loss(temperature_global=method_args.temperature_global,
                                            temperature_local=method_args.temperature_local)
        self.teacher_queue: Tensor
        self.register_buffer("teacher_queue", torch.zeros([method_args.queue_size, self.teacher_embedding_dim]))

    # ------------------------------------------------------------------ the step (:235-273)
    def training_step_impl(self, batch: Dict[str, Any], batch_idx: int = 0) -> TrainingStepResult:
        views = batch["views"][0]
        views = self._mixup_data(views)
        x_teacher_global, x_teacher_local, (th, tw) = self._forward_teacher(views)
        x_student_global, x_student_local = self._forward_student(views, th, tw)
        self._update_queue(x_teacher=x_teacher_global)
        global_loss, local_loss = self.criterion(teacher_features_global=x_teacher_global, teacher_features_local=x_teacher_local,
                                                 student_features_global=x_student_global,
                                                 student_features_local=x_student_local, queue=self.teacher_queue)
        loss = global_loss + self.method_args.loss_local_weight * local_loss
        return TrainingStepResult(loss=loss, log_dict={"train_loss/local_loss": local_loss.detach(),
                                                       "train_loss/global_loss": global_loss.detach()})

    @torch.no_grad()
    def _update_queue(self, x_teacher: Tensor) -> None:
        """:275-290 -- FIFO of teacher features, per GPU."""
        B = x_teacher.size(0)
        queue_size = self.teacher_queue.size(0)
        if B >= queue_size:
            self.teacher_queue = x_teacher[:queue_size].clone()
        else:
            self.teacher_queue[B:] = self.teacher_queue[:-B].clone()
            self.teacher_queue[:B] = x_teacher

    @torch.no_grad()
    def _forward_teacher(self, x: Tensor) -> Tuple[Tensor, Tensor, Tuple[int, int]]:
        """:292-319 -- the DINOv3 ViT forward on this package's kernels, then L2 normalisation."""
        output = self.teacher_embedding_model.forward_features(x)
        x_local = output["features"]
        x_global = self.teacher_embedding_model.forward_pool(output)["pooled_features"].flatten(1)
        th, tw = x_local.shape[-2:]
        x_local = x_local.permute(0, 2, 3, 1).flatten(start_dim=1, end_dim=2)
        x_local = F.normalize(x_local, dim=-1, p=2)
        x_global = F.normalize(x_global, dim=-1, p=2)
        return x_global, x_local, (th, tw)

    def _forward_student(self, x: Tensor, teacher_features_h: int, teacher_features_w: int) -> Tuple[Tensor, Tensor]:
        """:321-361 -- student backbone + linear heads + bilinear resize + L2 normalisation (torch autograd)."""
        x_global_local = self.student_embedding_model.wrapped_model.forward_features(x)
        x_global = self.student_embedding_model.wrapped_model.forward_pool(x_global_local)["pooled_features"]
        x_local = x_global_local["features"]
        x_global = self.flatten(x_global)
        x_local = x_local.permute(0, 2, 3, 1)
        x_global = self.student_projection_head_global(x_global)
        x_local = self.student_projection_head_local(x_local)
        x_local = x_local.permute(0, 3, 1, 2)
        x_local = F.interpolate(x_local, size=(teacher_features_h, teacher_features_w), mode="bilinear", align_corners=False)
        x_local = x_local.permute(0, 2, 3, 1).flatten(start_dim=1, end_dim=2)
        x_global = F.normalize(x_global, dim=-1, p=2)
        x_local = F.normalize(x_local, dim=-1, p=2)
        return x_global, x_local

    @staticmethod
    def _mixup_data(x: Tensor) -> Tensor:
        """:363-374 -- lambda ~ U(0, 1) and the permutation are drawn on the host like the reference does."""
        lambda_ = torch.empty(1).uniform_(0.0, 1.0).item()
        index = torch.randperm(x.size(0))
        return lambda_ * x + (1.0 - lambda_) * x[index.to(x.device), :]

    def trainable_modules(self):
        return [self.student_embedding_model, self.student_projection_head_global, self.student_projection_head_local]
