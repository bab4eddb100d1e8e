"""Host-side helpers of the DINOv2 method: block-wise mask generation and optimizer hyper-parameter tables.

Mirrors LT/_methods/dinov2/utils.py (MaskingGenerator :41-113, create_collated_masks :116-152,
get_vit_lr_decay_rate :155-186, get_optimizer_with_decay :191-250).  The mask generator consumes the python
`random` stream in the same order as the reference, so seeding `random` reproduces the reference's masks
bit for bit (pinned by tests/test_host_logic.py against tests/golden/masks.pt).
"""
from __future__ import annotations

import math
import random
from typing import Dict

import numpy as np
import torch


class MaskingGenerator:
    def __init__(self, input_size, max_num_patches: int, min_num_patches: int = 4, min_aspect: float = 0.3,
                 max_aspect: float | None = None) -> None:
        if not isinstance(input_size, tuple):
            input_size = (input_size, input_size)
        self.height, self.width = input_size
        self.num_patches = self.height * self.width
        self.min_num_patches = min_num_patches
        self.max_num_patches = max_num_patches
        hi = max_aspect or 1 / min_aspect
        self.log_aspect_ratio = (math.log(min_aspect), math.log(hi))

    def get_shape(self):
        return self.height, self.width

    def _try_block(self, mask: np.ndarray, budget: int) -> int:
        """Up to 10 proposals of a random rectangle; accept the first that adds 1..budget new patches."""
        for _ in range(10):
            area = random.uniform(self.min_num_patches, budget)
            ratio = math.exp(random.uniform(*self.log_aspect_ratio))
            bh = int(round(math.sqrt(area * ratio)))
            bw = int(round(math.sqrt(area / ratio)))
            if bw >= self.width or bh >= self.height:
                continue
            top = random.randint(0, self.height - bh)
            left = random.randint(0, self.width - bw)
            window = mask[top:top + bh, left:left + bw]
            fresh = bh * bw - int(window.sum())
            if 0 < fresh <= budget:
                window[...] = True
                return fresh
        return 0

    def __call__(self, num_masking_patches: int = 0) -> np.ndarray:
        mask = np.zeros((self.height, self.width), dtype=bool)
        count = 0
        while count < num_masking_patches:
            budget = min(num_masking_patches - count, self.max_num_patches)
            added = self._try_block(mask, budget)
            if added == 0:
                break
            count += added
        return mask


def create_collated_masks(mask_ratio_min: float, mask_ratio_max: float, n_masked_crops: int, n_crops: int,
                          mask_generator: MaskingGenerator) -> Dict[str, torch.Tensor]:
    n_tokens = mask_generator.num_patches
    edges = np.linspace(mask_ratio_min, mask_ratio_max, n_masked_crops + 1)
    masks = [torch.from_numpy(mask_generator(int(n_tokens * random.uniform(edges[i], edges[i + 1]))))
             for i in range(n_masked_crops)]
    masks += [torch.from_numpy(mask_generator(0)) for _ in range(n_masked_crops, n_crops)]
    random.shuffle(masks)
    collated = torch.stack(masks).flatten(1)
    indices = collated.flatten().nonzero().flatten()
    # == (1 / collated.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(collated)[collated] of the reference (:149),
    # in numpy: torch's boolean gather over the stride-0 expansion (and repeat_interleave) spin up the intra-op thread
    # pool and cost tens of ms per step on the host, more than the whole device step
    counts = collated.sum(-1).numpy()
    weight = torch.from_numpy(np.repeat(np.float32(1) / np.maximum(counts, 1).astype(np.float32), counts))
    return {"collated_masks": collated, "mask_indices_list": indices, "masks_weight": weight}


def get_vit_lr_decay_rate(name: str, lr_decay_rate: float, num_layers: int = 12, chunked_blocks: bool = False) -> float:
    layer_id = num_layers + 1
    if any(k in name for k in ("pos_embed", "patch_embed", "mask_token", "cls_token", "register_tokens")):
        layer_id = 0
    elif ".blocks." in name and ".residual." not in name:
        layer_id = int(name[name.find(".blocks."):].split(".")[2]) + 1
    elif chunked_blocks and "blocks." in name and "residual." not in name:
        layer_id = int(name[name.find("blocks."):].split(".")[2]) + 1
    elif "blocks." in name and "residual." not in name:
        layer_id = int(name[name.find("blocks."):].split(".")[1]) + 1
    return lr_decay_rate ** (num_layers + 1 - layer_id)


def param_group_settings(name: str, is_backbone: bool, num_layers: int, layerwise_decay: float,
                         patch_embed_lr_multiplier: float) -> Dict[str, float]:
    """lr multiplier / weight-decay switch of one parameter, as get_optimizer_with_decay assigns them.

    `name` is the parameter name inside its module (backbone or head container).  Returns
    {"lr_scale", "wd_scale", "last_layer", "head"}; the last two feed the lr-freeze logic of
    DINOv2.on_before_optimizer_step (dinov2.py:620-634), which keys on the *group name* = name of the first
    parameter of the fused group -- every parameter whose own name contains the substring falls in such a
    group because get_fused_param_groups (:253-273) splits groups on exactly these two flags.
    """
    lr_scale = get_vit_lr_decay_rate(name, layerwise_decay, num_layers) if is_backbone else 1.0
    wd_scale = 0.0 if (name.endswith(".bias") or "norm" in name or "gamma" in name) else 1.0
    if "patch_embed" in name:
        lr_scale *= patch_embed_lr_multiplier
    return {"lr_scale": lr_scale, "wd_scale": wd_scale, "last_layer": float("last_layer" in name),
            "head": float("head" in name)}
