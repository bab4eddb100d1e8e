"""DINOv3 DinoVisionTransformer (the distillation TEACHER) on B200 kernels -- forward only.

Mirror of LT/_models/dinov3/dinov3_src/models/vision_transformer.py:50-330 for the configuration the hub ViTs use
(`dinov3_vitb16` etc., hub/backbones.py:426-464): axial RoPE on the patch tokens instead of a positional embedding
(layers/rope_position_encoding.py:62-117), `n_storage_tokens` register tokens, LayerScale, `mask_k_bias` (the k third of
the qkv bias is multiplied by a 0/1 buffer, layers/attention.py:37-53), LayerNorm eps 1e-5 ("layernormbf16"), MLP FFN.
Same parameter / buffer names (`cls_token`, `storage_tokens`, `mask_token`, `patch_embed.proj.*`, `rope_embed.periods`,
`blocks.{i}.{norm1,attn.qkv(+bias_mask),attn.proj,ls1,norm2,mlp.fc1,mlp.fc2,ls2}.*`, `norm.*`) so reference checkpoints
load unchanged.  The teacher is frozen (DistillationV3, LT/_methods/distillationv3/distillationv3.py:276-303), so only
the inference schedule exists: LN -> GEMM qkv -> RoPE (in place) -> attention -> GEMM proj (+LayerScale+residual) -> LN ->
GEMM fc1 (+bias+GELU) -> GEMM fc2 (+LayerScale+residual), all through libb200dino.so.

Rounding follows bf16 autocast like the DINOv2 mirror; attention keeps the DINOv2 kernels' bf16-rounded scores (the
reference calls SDPA here, whose flash kernel keeps fp32 scores: a difference inside the bf16 noise of the features).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops
from .._arena import Arena
from .dinov2_vit import attach_params


def dinov3_param_shapes(embed_dim: int, depth: int, patch_size: int, in_chans: int, hidden: int, n_storage_tokens: int,
                        layerscale: bool) -> Dict[str, Tuple[int, ...]]:
    D, p = embed_dim, patch_size
    s: Dict[str, Tuple[int, ...]] = {"cls_token": (1, 1, D)}
    if n_storage_tokens:
        s["storage_tokens"] = (1, n_storage_tokens, D)
    s["mask_token"] = (1, D)
    s["patch_embed.proj.weight"] = (D, in_chans, p, p)
    s["patch_embed.proj.bias"] = (D,)
    for i in range(depth):
        b = f"blocks.{i}."
        s[b + "norm1.weight"] = (D,); s[b + "norm1.bias"] = (D,)
        s[b + "attn.qkv.weight"] = (3 * D, D); s[b + "attn.qkv.bias"] = (3 * D,)
        s[b + "attn.proj.weight"] = (D, D); s[b + "attn.proj.bias"] = (D,)
        if layerscale:
            s[b + "ls1.gamma"] = (D,)
        s[b + "norm2.weight"] = (D,); s[b + "norm2.bias"] = (D,)
        s[b + "mlp.fc1.weight"] = (hidden, D); s[b + "mlp.fc1.bias"] = (hidden,)
        s[b + "mlp.fc2.weight"] = (D, hidden); s[b + "mlp.fc2.bias"] = (D,)
        if layerscale:
            s[b + "ls2.gamma"] = (D,)
    s["norm.weight"] = (D,); s["norm.bias"] = (D,)
    return s


class _RopeBuffers(nn.Module):
    def __init__(self, periods: Tensor) -> None:
        super().__init__()
        self.register_buffer("periods", periods, persistent=True)


class DinoV3VisionTransformer(nn.Module):
    def __init__(self, *, img_size: int = 224, patch_size: int = 16, in_chans: int = 3, pos_embed_rope_base: float = 100.0,
                 pos_embed_rope_normalize_coords: str = "separate", pos_embed_rope_dtype: str = "fp32", embed_dim: int = 768,
                 depth: int = 12, num_heads: int = 12, ffn_ratio: float = 4.0, qkv_bias: bool = True,
                 layerscale_init: Optional[float] = None, norm_layer: str = "layernorm", ffn_layer: str = "mlp",
                 n_storage_tokens: int = 0, mask_k_bias: bool = False, device: str = "cuda", **ignored_kwargs) -> None:
        super().__init__()
        if ffn_layer != "mlp" or not qkv_bias or embed_dim // num_heads != 64 or embed_dim % num_heads:
            raise NotImplementedError("b200 DinoV3VisionTransformer: mlp FFN, qkv bias, head_dim 64")
        if norm_layer not in ("layernorm", "layernormbf16"):
            raise NotImplementedError("b200 DinoV3VisionTransformer: LayerNorm variants only")
        if pos_embed_rope_dtype not in ("fp32", "bf16"):
            raise NotImplementedError(pos_embed_rope_dtype)
        self.embed_dim = self.num_features = embed_dim
        self.n_blocks = depth
        self.num_heads = num_heads
        self.patch_size = patch_size
        self.in_chans = in_chans
        self.n_storage_tokens = n_storage_tokens
        self.mask_k_bias = mask_k_bias
        self.layerscale = bool(layerscale_init)
        self.ln_eps = 1e-5 if norm_layer == "layernormbf16" else 1e-6  # vision_transformer.py:43-47
        self.hidden_dim = int(embed_dim * ffn_ratio)
        self.rope_normalize_coords = pos_embed_rope_normalize_coords
        self.rope_dtype = torch.float32 if pos_embed_rope_dtype == "fp32" else torch.bfloat16
        shapes = dinov3_param_shapes(embed_dim, depth, patch_size, in_chans, self.hidden_dim, n_storage_tokens, self.layerscale)
        self.arena = Arena(shapes, device, with_grad=False, with_optim_state=False)
        attach_params(self, self.arena, "", list(shapes), requires_grad=False)
        # attach_params builds `blocks` as a name-keyed Module tree; make it indexable like the reference's ModuleList
        # (same state_dict names: blocks.{i}.*)
        self._modules["blocks"] = nn.ModuleList([self._modules["blocks"]._modules[str(i)] for i in range(depth)])
        d_head = embed_dim // num_heads
        periods = pos_embed_rope_base ** (2 * torch.arange(d_head // 4, dtype=self.rope_dtype) / (d_head // 2))  # :120-126
        self.rope_embed = _RopeBuffers(periods.to(device))
        for i in range(depth):  # LinearKMaskedBias.bias_mask (attention.py:42-46), filled by init_weights (:206-211)
            m = torch.ones(3 * embed_dim, device=device)
            if mask_k_bias:
                m[embed_dim:2 * embed_dim] = 0
            self.blocks[i].attn.qkv.register_buffer("bias_mask", m)  # type: ignore[index]
        self._rope_cache: Dict[Tuple[int, int], Tuple[Tensor, Tensor]] = {}
        self._qkv_bias_eff: Optional[list] = None
        self.init_weights(layerscale_init)
        self.register_load_state_dict_post_hook(lambda mod, inc: mod._invalidate())

    def _invalidate(self) -> None:
        self.arena.bf16_valid = False
        self._qkv_bias_eff = None
        self._rope_cache.clear()

    @torch.no_grad()
    def init_weights(self, layerscale_init: Optional[float]) -> None:
        """vision_transformer.py:196-213 + init_weights_vit (:33-47, 338+): trunc_normal(0.02) linears, zero biases, unit norms."""
        for name, prm in self.named_parameters():
            if name in ("cls_token", "storage_tokens"):
                nn.init.normal_(prm, std=0.02)
            elif name == "mask_token":
                prm.zero_()
            elif name == "patch_embed.proj.weight":
                nn.init.kaiming_uniform_(prm, a=math.sqrt(5))
            elif name == "patch_embed.proj.bias":
                bound = 1.0 / math.sqrt(self.in_chans * self.patch_size ** 2)
                nn.init.uniform_(prm, -bound, bound)
            elif name.endswith("gamma"):
                prm.fill_(layerscale_init)
            elif "norm" in name and name.endswith("weight"):
                prm.fill_(1.0)
            elif name.endswith("bias"):
                prm.zero_()
            else:
                nn.init.trunc_normal_(prm, std=0.02)
        self._invalidate()

    # ------------------------------------------------------------------ helpers
    def _P(self, name: str) -> Tensor:
        return self.arena.p(name)

    def _W(self, name: str) -> Tensor:
        return self.arena.w(name)

    def _rope(self, Hp: int, Wp: int) -> Tuple[Tensor, Tensor]:
        """sin / cos tables [Hp*Wp, 64] (rope_position_encoding.py:62-117, eval mode: no shift / jitter / rescale), computed in
        the rope dtype like the reference and handed to the kernel as fp32."""
        key = (Hp, Wp)
        if key not in self._rope_cache:
            dev, dt = self.arena.device, self.rope_dtype
            periods = self.rope_embed.periods.to(dt)
            if self.rope_normalize_coords == "separate":
                dh, dw = Hp, Wp
            elif self.rope_normalize_coords == "max":
                dh = dw = max(Hp, Wp)
            elif self.rope_normalize_coords == "min":
                dh = dw = min(Hp, Wp)
            else:
                raise ValueError(f"Unknown normalize_coords: {self.rope_normalize_coords}")
            ch = torch.arange(0.5, Hp, device=dev, dtype=dt) / dh
            cw = torch.arange(0.5, Wp, device=dev, dtype=dt) / dw
            coords = torch.stack(torch.meshgrid(ch, cw, indexing="ij"), dim=-1).flatten(0, 1)
            coords = 2.0 * coords - 1.0
            angles = (2 * math.pi * coords[:, :, None] / periods[None, None, :]).flatten(1, 2)
            angles = torch.cat((angles, angles), dim=-1)
            self._rope_cache[key] = (torch.sin(angles).float().contiguous(), torch.cos(angles).float().contiguous())
        return self._rope_cache[key]

    def _bias_eff(self, i: int) -> Tensor:
        if self._qkv_bias_eff is None:
            self._qkv_bias_eff = [(self._P(f"blocks.{j}.attn.qkv.bias") * self.blocks[j].attn.qkv.bias_mask).contiguous()
                                  for j in range(self.n_blocks)]
        return self._qkv_bias_eff[i]

    # ------------------------------------------------------------------ forward (inference schedule)
    @torch.no_grad()
    def forward_features(self, x: Tensor, masks: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """vision_transformer.py:260-311 (single resolution): returns x_norm_clstoken [B, D], x_storage_tokens [B, S, D],
        x_norm_patchtokens [B, Hp*Wp, D], x_prenorm [B, N, D], masks."""
        if not self.arena.bf16_valid:
            self.arena.refresh_bf16()
        dev = x.device
        Bc, Cin, H, Wimg = x.shape
        p, D, Hd, h = self.patch_size, self.embed_dim, self.hidden_dim, self.num_heads
        Hp, Wp = H // p, Wimg // p
        Np, S = Hp * Wp, self.n_storage_tokens
        N = 1 + S + Np
        T = Bc * N
        bf, f32 = torch.bfloat16, torch.float32
        E = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        cols = E(Bc * Np, Cin * p * p)
        ops.im2col(x.contiguous().float(), p, cols)
        tok = E(Bc * Np, D)
        ops.gemm(cols, self._W("patch_embed.proj.weight").view(D, -1), tok, bias=self._P("patch_embed.proj.bias"))
        masks_u8 = masks.to(torch.uint8).contiguous() if masks is not None else None
        xs = E(Bc, N, D, dt=f32)
        zero_pos = torch.zeros(1 + Np, D, device=dev, dtype=f32)  # no additive positional embedding: RoPE inside attention
        ops.assemble_tokens(tok, masks_u8, self._P("mask_token") if masks is not None else None, self._P("cls_token").view(D),
                            self._P("storage_tokens").view(S, D) if S else None, zero_pos, Bc, Np, S, D, xs)
        xcur = xs.view(T, D)
        sin_t, cos_t = self._rope(Hp, Wp)
        scale = 64 ** -0.5
        for i in range(self.n_blocks):
            b = f"blocks.{i}."
            xn = E(T, D)
            ops.layernorm_fwd(xcur, self._P(b + "norm1.weight"), self._P(b + "norm1.bias"), self.ln_eps, xn)
            qkv = E(T, 3 * D)
            ops.gemm(xn, self._W(b + "attn.qkv.weight"), qkv, bias=self._bias_eff(i))
            ops.rope_apply(qkv, Bc, N, 1 + S, h, sin_t, cos_t)
            att = E(T, D)
            ops.attention_fwd(qkv, Bc, N, h, att, None, scale)
            xmid = E(T, D, dt=f32)
            ops.gemm(att, self._W(b + "attn.proj.weight"), xmid, epi=ops.EPI_RESIDUAL, bias=self._P(b + "attn.proj.bias"),
                     aux=xcur, gamma=self._P(b + "ls1.gamma") if self.layerscale else None)
            xn2 = E(T, D)
            ops.layernorm_fwd(xmid, self._P(b + "norm2.weight"), self._P(b + "norm2.bias"), self.ln_eps, xn2)
            hh = E(T, Hd)
            ops.gemm(xn2, self._W(b + "mlp.fc1.weight"), hh, epi=ops.EPI_BIAS_GELU, bias=self._P(b + "mlp.fc1.bias"))
            xout = E(T, D, dt=f32)
            ops.gemm(hh, self._W(b + "mlp.fc2.weight"), xout, epi=ops.EPI_RESIDUAL, bias=self._P(b + "mlp.fc2.bias"), aux=xmid,
                     gamma=self._P(b + "ls2.gamma") if self.layerscale else None)
            xcur = xout
        xnorm = E(T, D, dt=f32)
        ops.layernorm_fwd(xcur, self._P("norm.weight"), self._P("norm.bias"), self.ln_eps, xnorm)
        xn3 = xnorm.view(Bc, N, D)
        return {"x_norm_clstoken": xn3[:, 0], "x_storage_tokens": xn3[:, 1:1 + S], "x_norm_patchtokens": xn3[:, 1 + S:],
                "x_prenorm": xcur.view(Bc, N, D), "masks": masks}

    def forward(self, *args, is_training: bool = False, **kwargs):
        ret = self.forward_features(*args, **kwargs)
        return ret if is_training else ret["x_norm_clstoken"]


class DINOv3ViTModelWrapper(nn.Module):
    """LT/_models/dinov3/dinov3_vit.py:31-97: forward_features -> {"features" [B, D, H, W], "cls_token"}, forward_pool."""

    def __init__(self, model: DinoV3VisionTransformer) -> None:
        super().__init__()
        self._model = model

    def feature_dim(self) -> int:
        return int(self._model.embed_dim)

    def patch_size(self) -> int:
        return int(self._model.patch_size)

    def get_model(self) -> DinoV3VisionTransformer:
        return self._model

    @torch.no_grad()
    def forward_features(self, x: Tensor, masks: Optional[Tensor] = None) -> Dict[str, Tensor]:
        rt = self._model.forward_features(x, masks)
        pt = rt["x_norm_patchtokens"]
        b, _, d = pt.shape
        hh, ww = x.shape[2] // self._model.patch_size, x.shape[3] // self._model.patch_size
        return {"features": pt.permute(0, 2, 1).reshape(b, d, hh, ww), "cls_token": rt["x_norm_clstoken"]}

    def forward_pool(self, x: Dict[str, Tensor]) -> Dict[str, Tensor]:
        return {"pooled_features": x["cls_token"][..., None, None]}
