"""CPU restatement of the DINOv3 ViT *teacher forward* (distillation cfg4) -- TEST INFRASTRUCTURE
ONLY (same rule as dinov2_oracle.py: tests/, smoke() and bench's CPU arm may import it; the product path never does).

Follows LT/_models/dinov3/dinov3_src:
  models/vision_transformer.py:220-311   prepare_tokens_with_masks / forward_features_list (eval mode: no RoPE jitter)
  layers/rope_position_encoding.py:62-136 axial RoPE sin/cos tables (periods = base^(2i/(D_head/2)), coords in [-1, 1])
  layers/attention.py:21-133             rope_apply on q, k of the PATCH tokens only (cls + storage tokens untouched),
                                         LinearKMaskedBias (the k third of the qkv bias is multiplied by 0), SDPA
  layers/block.py:103-141                pre-norm residual block with LayerScale
  layers/ffn_layers.py:28-53             Mlp (fc1, GELU, fc2)
Weights are a plain state_dict with the reference's names.  Pinned against the imported reference module
(tests/golden/dinov3_tiny.pt).  The CUDA path for this model is not built yet.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


@dataclass
class Dinov3Config:
    """Constructor arguments of dinov3 DinoVisionTransformer that change the arithmetic (hub ViT-B/16:
    embed 768, depth 12, heads 12, layerscale 1e-5, 4 storage tokens, masked k bias, norm "layernormbf16")."""

    embed_dim: int = 768
    depth: int = 12
    num_heads: int = 12
    patch_size: int = 16
    ffn_ratio: float = 4.0
    layerscale_init: Optional[float] = 1e-5
    n_storage_tokens: int = 4
    mask_k_bias: bool = True
    ln_eps: float = 1e-5            # norm_layer="layernormbf16" (vision_transformer.py:43-47); "layernorm" = 1e-6
    rope_base: float = 100.0
    rope_normalize_coords: str = "separate"


def rope_sincos(cfg: Dinov3Config, H: int, W: int) -> Tuple[Tensor, Tensor]:
    """rope_position_encoding.py:62-117 in eval mode (no shift / jitter / rescale), fp32.  Returns 2 x [H*W, D_head]."""
    d_head = cfg.embed_dim // cfg.num_heads
    periods = cfg.rope_base ** (2 * torch.arange(d_head // 4, dtype=torch.float32) / (d_head // 2))  # :120-126
    if cfg.rope_normalize_coords == "separate":
        ch, cw = torch.arange(0.5, H, dtype=torch.float32) / H, torch.arange(0.5, W, dtype=torch.float32) / W
    elif cfg.rope_normalize_coords == "max":
        m = max(H, W)
        ch, cw = torch.arange(0.5, H, dtype=torch.float32) / m, torch.arange(0.5, W, dtype=torch.float32) / m
    elif cfg.rope_normalize_coords == "min":
        m = min(H, W)
        ch, cw = torch.arange(0.5, H, dtype=torch.float32) / m, torch.arange(0.5, W, dtype=torch.float32) / m
    else:
        raise ValueError(cfg.rope_normalize_coords)
    coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).flatten(0, 1)  # [HW, 2] (row, col)
    coords = 2.0 * coords - 1.0
    angles = 2 * math.pi * coords[:, :, None] / periods[None, None, :]  # [HW, 2, D_head/4]
    angles = angles.flatten(1, 2)
    angles = torch.cat((angles, angles), dim=-1)  # [HW, D_head]
    return torch.sin(angles), torch.cos(angles)


def _rope_apply(x: Tensor, sin: Tensor, cos: Tensor) -> Tensor:
    """attention.py:21-33: x*cos + rotate_half(x)*sin with rotate_half([a, b]) = [-b, a]."""
    a, b = x.chunk(2, dim=-1)
    return x * cos + torch.cat((-b, a), dim=-1) * sin


def attention(sd: Dict[str, Tensor], pre: str, cfg: Dinov3Config, x: Tensor, rope: Tuple[Tensor, Tensor]) -> Tensor:
    B, N, D = x.shape
    h = cfg.num_heads
    bias = sd[pre + "qkv.bias"]
    if cfg.mask_k_bias:  # LinearKMaskedBias: bias * bias_mask, mask = 1 | 0 (k third) | 1  (vision_transformer.py:61-64)
        mask = torch.ones_like(bias)
        mask[D:2 * D] = 0
        bias = bias * mask
    qkv = F.linear(x, sd[pre + "qkv.weight"], bias).reshape(B, N, 3, h, D // h)
    q, k, v = (t.transpose(1, 2) for t in qkv.unbind(2))  # [B, h, N, d]
    sin, cos = rope
    prefix = N - sin.shape[-2]  # cls + storage tokens keep their q, k (attention.py:79-100)
    q = torch.cat((q[:, :, :prefix], _rope_apply(q[:, :, prefix:], sin, cos)), dim=2)
    k = torch.cat((k[:, :, :prefix], _rope_apply(k[:, :, prefix:], sin, cos)), dim=2)
    s = (q @ k.transpose(-1, -2)) * (D // h) ** -0.5
    o = (s.softmax(-1) @ v).transpose(1, 2).reshape(B, N, D)
    return F.linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"])


def block(sd: Dict[str, Tensor], i: int, cfg: Dinov3Config, x: Tensor, rope: Tuple[Tensor, Tensor]) -> Tensor:
    pre = f"blocks.{i}."
    D = cfg.embed_dim
    y = attention(sd, pre + "attn.", cfg, F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], cfg.ln_eps),
                  rope)
    if cfg.layerscale_init:
        y = y * sd[pre + "ls1.gamma"]
    x = x + y
    y = F.layer_norm(x, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], cfg.ln_eps)
    y = F.linear(F.gelu(F.linear(y, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"])), sd[pre + "mlp.fc2.weight"],
                 sd[pre + "mlp.fc2.bias"])
    if cfg.layerscale_init:
        y = y * sd[pre + "ls2.gamma"]
    return x + y


def forward_features(sd: Dict[str, Tensor], cfg: Dinov3Config, x: Tensor, masks: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """x [B, 3, H, W] -> cls [B, D], storage [B, S, D], patch [B, H/p * W/p, D], prenorm [B, N, D]."""
    p, D = cfg.patch_size, cfg.embed_dim
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=p)  # [B, D, H/p, W/p]
    B, _, Hp, Wp = t.shape
    t = t.flatten(2).transpose(1, 2)  # row-major (h, w) token order == patch_embed(...).flatten(1, 2)
    if masks is not None:
        t = torch.where(masks.unsqueeze(-1), sd["mask_token"].to(t.dtype).unsqueeze(0), t)
    toks = [sd["cls_token"].expand(B, -1, -1)]
    if cfg.n_storage_tokens:
        toks.append(sd["storage_tokens"].expand(B, -1, -1))
    xs = torch.cat(toks + [t], dim=1)
    rope = rope_sincos(cfg, Hp, Wp)
    for i in range(cfg.depth):
        xs = block(sd, i, cfg, xs, rope)
    xn = F.layer_norm(xs, (D,), sd["norm.weight"], sd["norm.bias"], cfg.ln_eps)
    S = cfg.n_storage_tokens
    return {"cls": xn[:, 0], "storage": xn[:, 1:1 + S], "patch": xn[:, 1 + S:], "prenorm": xs}
