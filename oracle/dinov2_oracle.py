"""CPU oracle for the DINOv2 training-step hot path of lightly-train (TEST INFRASTRUCTURE ONLY).

This module is a plain-PyTorch (CPU, fp32) restatement of the arithmetic the reference executes on the path
named by BASELINE.json's north_star.  It is the *checker* for the CUDA kernels in lightly_train_b200/csrc:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.  The
product path (lightly_train_b200.*) never imports it and has no CPU fallback.

Parity pinning: `tests/test_oracle_goldens.py` checks every function here against
  (a) the golden numbers held by the reference's own tests (DINOLoss 1.5565, IBOTPatchLoss 0.4057, center 0.2,
      EMA [[2.5,3.5],[4.5,5.5]] -- tests/_methods/dinov2/test_dinov2_loss.py, tests/test__torch_helpers.py), and
  (b) fixtures under tests/golden/ produced by importing the reference's own modules from /root/reference
      in the build container (tools/make_golden.py, committed).
KoLeoLoss / cosine_schedule live in the third-party `lightly` package (1.5.26, not vendored, not installed):
restated from their published definitions; the reference tests do not pin their values ("parity unpinned"
for those two functions only).

Every function cites the reference file:line it follows (LT = src/lightly_train).

`autocast=True` emulates torch.autocast("cuda", bfloat16) rounding points: matmul-class
ops take bf16-rounded inputs, accumulate in fp32 and round the result to bf16 once; layer_norm / softmax /
normalize / losses stay fp32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor


# --------------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------------
@dataclass
class ViTConfig:
    """Subset of DinoVisionTransformer.__init__ arguments (LT/_models/dinov2_vit/dinov2_vit_src/models/
    vision_transformer.py:84-107) that changes the arithmetic."""

    embed_dim: int = 384
    depth: int = 12
    num_heads: int = 6
    patch_size: int = 16
    img_size: int = 224
    mlp_ratio: float = 4.0
    init_values: Optional[float] = 1e-5
    num_register_tokens: int = 0
    interpolate_offset: float = 0.1
    interpolate_antialias: bool = False
    ln_eps: float = 1e-6
    ffn_layer: str = "mlp"  # "mlp" | "swiglu" / "swiglufused" (vision_transformer.py:179-184)

    @property
    def num_patches(self) -> int:
        return (self.img_size // self.patch_size) ** 2

    @property
    def hidden_dim(self) -> int:
        h = int(self.embed_dim * self.mlp_ratio)
        if self.ffn_layer in ("swiglu", "swiglufused"):
            h = (int(h * 2 / 3) + 7) // 8 * 8  # SwiGLUFFNFused, layers/swiglu_ffn.py:60-63
        return h


@dataclass
class HeadConfig:
    """DINOv2ProjectionHead arguments (LT/_methods/dinov2/dinov2_head.py:33-42), use_bn=False, nlayers=3."""

    in_dim: int = 384
    hidden_dim: int = 2048
    bottleneck_dim: int = 256
    out_dim: int = 65536


# --------------------------------------------------------------------------------------------------
# autocast emulation primitives
# --------------------------------------------------------------------------------------------------
def _r(x: Tensor) -> Tensor:
    """Round to bf16 and return as fp32 (value-level emulation of a bf16 tensor)."""
    return x.to(torch.bfloat16).to(torch.float32)


def linear(x: Tensor, w: Tensor, b: Optional[Tensor], autocast: bool) -> Tensor:
    """F.linear; under autocast: bf16 inputs, fp32 accumulate (+bias in fp32), one rounding to bf16."""
    if not autocast:
        return F.linear(x, w, b)
    y = F.linear(_r(x), _r(w), None if b is None else _r(b))
    return _r(y)


def matmul(a: Tensor, b: Tensor, autocast: bool) -> Tensor:
    if not autocast:
        return a @ b
    return _r(_r(a) @ _r(b))


def gelu(x: Tensor, autocast: bool) -> Tensor:
    """nn.GELU() (erf form). On a bf16 input the op computes in fp32 and rounds the output to bf16."""
    y = F.gelu(x)
    return _r(y) if autocast else y


# --------------------------------------------------------------------------------------------------
# ViT backbone  (vision_transformer.py:251-384, layers/*.py)
# --------------------------------------------------------------------------------------------------
def interpolate_pos_encoding(sd: Dict[str, Tensor], cfg: ViTConfig, npatch: int, w: int, h: int) -> Tensor:
    """vision_transformer.py:251-305. Returns [1, 1+npatch, D] fp32."""
    pos_embed = sd["pos_embed"].float()
    N = pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return pos_embed
    class_pos = pos_embed[:, :1]
    patch_pos = pos_embed[:, 1:]
    dim = pos_embed.shape[-1]
    w0 = w // cfg.patch_size
    h0 = h // cfg.patch_size
    M = int(math.sqrt(N))
    assert M * M == N
    kwargs = {}
    if cfg.interpolate_offset:
        # the reference passes scale factors (w0+offset)/M, not an output size (:283-291)
        kwargs["scale_factor"] = (float(w0 + cfg.interpolate_offset) / M, float(h0 + cfg.interpolate_offset) / M)
    else:
        kwargs["size"] = (w0, h0)
    grid = patch_pos.reshape(1, M, M, dim).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, mode="bicubic", antialias=cfg.interpolate_antialias, **kwargs)
    assert (w0, h0) == tuple(grid.shape[-2:])
    grid = grid.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat([class_pos, grid], dim=1)


def patch_embed(sd: Dict[str, Tensor], cfg: ViTConfig, x: Tensor, autocast: bool) -> Tensor:
    """layers/patch_embed.py:92-113: Conv2d(k=s=patch) == per-patch linear over (c, ky, kx). -> [B, Np, D]"""
    B, Cin, H, W = x.shape
    p = cfg.patch_size
    assert H % p == 0 and W % p == 0, "oracle covers patch-aligned inputs only"
    w = sd["patch_embed.proj.weight"].reshape(cfg.embed_dim, Cin * p * p)
    b = sd["patch_embed.proj.bias"]
    cols = x.reshape(B, Cin, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), Cin * p * p)
    return linear(cols, w, b, autocast)


def prepare_tokens(sd: Dict[str, Tensor], cfg: ViTConfig, x: Tensor, masks: Optional[Tensor], autocast: bool) -> Tensor:
    """vision_transformer.py:307-329. -> [B, 1+R+Np, D] fp32."""
    B, _, H, W = x.shape
    tok = patch_embed(sd, cfg, x, autocast)
    if masks is not None:
        mt = sd["mask_token"]
        mt = _r(mt) if autocast else mt  # mask_token.to(x.dtype), x is bf16 under autocast
        tok = torch.where(masks.unsqueeze(-1), mt.unsqueeze(0), tok)
    cls = sd["cls_token"].expand(B, -1, -1)
    tok = torch.cat([cls, tok], dim=1)  # promotes to fp32
    tok = tok + interpolate_pos_encoding(sd, cfg, tok.shape[1] - 1, H, W)
    if cfg.num_register_tokens:
        reg = sd["register_tokens"].expand(B, -1, -1)
        tok = torch.cat([tok[:, :1], reg, tok[:, 1:]], dim=1)
    return tok


def attention(sd: Dict[str, Tensor], pre: str, cfg: ViTConfig, x: Tensor, autocast: bool) -> Tensor:
    """layers/attention.py:49-66 (the non-xformers path). x: LN output [B,N,D]."""
    B, N, D = x.shape
    h = cfg.num_heads
    dh = D // h
    qkv = linear(x, sd[pre + "qkv.weight"], sd[pre + "qkv.bias"], autocast)
    qkv = qkv.reshape(B, N, 3, h, dh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (dh ** -0.5), qkv[1], qkv[2]
    if autocast:
        q = _r(q)
    s = matmul(q, k.transpose(-2, -1), autocast)
    p = s.softmax(dim=-1)  # fp32
    o = matmul(p, v, autocast).transpose(1, 2).reshape(B, N, D)
    return linear(o, sd[pre + "proj.weight"], sd[pre + "proj.bias"], autocast)


def mlp(sd: Dict[str, Tensor], pre: str, x: Tensor, autocast: bool) -> Tensor:
    """layers/mlp.py:36-42, or SwiGLUFFN.forward (layers/swiglu_ffn.py:31-35) when the block carries w12/w3."""
    if pre + "w12.weight" in sd:
        x12 = linear(x, sd[pre + "w12.weight"], sd[pre + "w12.bias"], autocast)
        x1, x2 = x12.chunk(2, dim=-1)
        a = F.silu(x1)
        hidden = _r(_r(a) * x2) if autocast else a * x2  # bf16 silu output, bf16 product under autocast
        return linear(hidden, sd[pre + "w3.weight"], sd[pre + "w3.bias"], autocast)
    u = linear(x, sd[pre + "fc1.weight"], sd[pre + "fc1.bias"], autocast)
    return linear(gelu(u, autocast), sd[pre + "fc2.weight"], sd[pre + "fc2.bias"], autocast)


def block(sd: Dict[str, Tensor], i: int, cfg: ViTConfig, x: Tensor, autocast: bool,
          keep_scale: Optional[Tensor] = None) -> Tensor:
    """layers/block.py:90-115 (plain residual / per-sample DropPath regimes).

    keep_scale: optional [2, B] per-sample multipliers (bernoulli(keep)/keep) for the two residual branches,
    i.e. the random tensor of drop_path.py:23-27 made an explicit input.
    """
    pre = f"blocks.{i}."
    D = cfg.embed_dim
    y = F.layer_norm(x, (D,), sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], cfg.ln_eps)
    y = attention(sd, pre + "attn.", cfg, y, autocast)
    if cfg.init_values:
        y = y * sd[pre + "ls1.gamma"]
    if keep_scale is not None:
        y = y * keep_scale[0].view(-1, 1, 1)
    x = x + y
    y = F.layer_norm(x, (D,), sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], cfg.ln_eps)
    y = mlp(sd, pre + "mlp.", y, autocast)
    if cfg.init_values:
        y = y * sd[pre + "ls2.gamma"]
    if keep_scale is not None:
        y = y * keep_scale[1].view(-1, 1, 1)
    return x + y


def vit_forward_features(sd: Dict[str, Tensor], cfg: ViTConfig, x: Tensor, masks: Optional[Tensor] = None,
                         autocast: bool = False, keep_scales: Optional[Sequence[Tensor]] = None,
                         taps: Optional[dict] = None) -> Dict[str, Tensor]:
    """DinoVisionTransformer.forward_features (vision_transformer.py:361-384).

    Returns {"cls": [B,D], "patch": [B,Np,D], "prenorm": [B,N,D]}. `taps` (if given) receives intermediates.
    """
    t = prepare_tokens(sd, cfg, x, masks, autocast)
    if taps is not None:
        taps["tokens"] = t.detach().clone()
    for i in range(cfg.depth):
        t = block(sd, i, cfg, t, autocast, None if keep_scales is None else keep_scales[i])
        if taps is not None:
            taps[f"block{i}"] = t.detach().clone()
    xn = F.layer_norm(t, (cfg.embed_dim,), sd["norm.weight"], sd["norm.bias"], cfg.ln_eps)
    R = cfg.num_register_tokens
    return {"cls": xn[:, 0], "patch": xn[:, 1 + R:], "prenorm": t}


# --------------------------------------------------------------------------------------------------
# projection head (LT/_methods/dinov2/dinov2_head.py:44-71)
# --------------------------------------------------------------------------------------------------
def weight_norm_weight(g: Tensor, v: Tensor) -> Tensor:
    """parametrizations.weight_norm(dim=0): W[o,:] = g[o] * v[o,:] / ||v[o,:]||_2."""
    return g * v / v.norm(dim=1, keepdim=True)


def head_forward(sd: Dict[str, Tensor], x: Tensor, autocast: bool = False, taps: Optional[dict] = None) -> Tensor:
    y = gelu(linear(x, sd["mlp.0.weight"], sd["mlp.0.bias"], autocast), autocast)
    y = gelu(linear(y, sd["mlp.2.weight"], sd["mlp.2.bias"], autocast), autocast)
    y = linear(y, sd["mlp.4.weight"], sd["mlp.4.bias"], autocast)
    y = F.normalize(y, dim=-1, p=2, eps=1e-12)  # fp32 under autocast
    if taps is not None:
        taps["bottleneck"] = y.detach().clone()
    w = weight_norm_weight(sd["last_layer.parametrizations.weight.original0"],
                           sd["last_layer.parametrizations.weight.original1"])
    return linear(y, w, None, autocast)


# --------------------------------------------------------------------------------------------------
# losses (LT/_methods/dinov2/dinov2_loss.py)
# --------------------------------------------------------------------------------------------------
def softmax_center_teacher(t: Tensor, center: Tensor, teacher_temp: float) -> Tensor:
    """dinov2_loss.py:76-82 / :178-186."""
    return F.softmax((t.float() - center) / teacher_temp, dim=-1)


def center_batch_sum_dino(t: Tensor) -> Tensor:
    """reduce_center_update, dinov2_loss.py:139-145: per-rank sum over rows (before all-reduce). [1,K]"""
    return t.float().sum(dim=0, keepdim=True)


def center_batch_sum_ibot(t: Tensor) -> Tensor:
    """IBOT reduce_center_update, dinov2_loss.py:274-282 with t [1,M,K]: sum over dim0 of mean over dim1."""
    return t.float().mean(dim=1).sum(dim=0, keepdim=True)


def center_ema(center: Tensor, batch_sum: Tensor, n_rows_total: int, momentum: float) -> Tensor:
    """apply_center_update, dinov2_loss.py:148-160: center*m + (sum/(len*world))*(1-m)."""
    return center * momentum + (batch_sum / n_rows_total) * (1 - momentum)


def sinkhorn_knopp(t: Tensor, teacher_temp: float, n_samples_total: float, n_iterations: int = 3,
                   all_reduce=None) -> Tensor:
    """dinov2_loss.py:84-115 / :188-224. `all_reduce(x)` sums a tensor over ranks (identity if None).

    No max-subtraction before exp, and the in-place division order is kept (fp32 rounding order matters).
    """
    ar = all_reduce if all_reduce is not None else (lambda z: z)
    Q = torch.exp(t.float() / teacher_temp).t().contiguous()
    K = Q.shape[0]
    Bt = n_samples_total
    Q = Q / ar(Q.sum())
    for _ in range(n_iterations):
        Q = Q / ar(Q.sum(dim=1, keepdim=True))
        Q = Q / K
        Q = Q / Q.sum(dim=0, keepdim=True)
        Q = Q / Bt
    Q = Q * Bt
    return Q.t()


def dino_loss(student_list: Sequence[Tensor], teacher_list: Sequence[Tensor], student_temp: float = 0.1) -> Tensor:
    """DINOLoss.forward, dinov2_loss.py:117-133."""
    total = torch.zeros((), dtype=torch.float32)
    for s in student_list:
        lsm = F.log_softmax(s.float() / student_temp, dim=-1)
        for t in teacher_list:
            total = total - (t * lsm).sum(dim=-1).mean()
    return total


def ibot_loss_masked(s: Tensor, t: Tensor, masks_weight: Tensor, n_images: int, student_temp: float = 0.1) -> Tensor:
    """IBOTPatchLoss.forward_masked, dinov2_loss.py:246-268 (+ lossfunc :55-56)."""
    per_tok = (t * F.log_softmax(s.float() / student_temp, dim=-1)).sum(dim=-1)
    return -(per_tok * masks_weight).sum() / n_images


def masks_weight_from_masks(masks: Tensor) -> Tensor:
    """dinov2/utils.py:141-146."""
    return (1 / masks.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(masks)[masks]


def koleo_loss(x: Tensor, eps: float = 1e-8, autocast: bool = False) -> Tensor:
    """lightly.loss.KoLeoLoss (lightly 1.5.26; call site LT/_methods/dinov2/dinov2.py:377-380).

    L2-normalise, nearest neighbour by cosine similarity (diagonal excluded), PairwiseDistance(p=2, eps)
    (= ||a - b + eps||_2), loss = -mean(log(d + eps)). Under autocast the similarity matmul is bf16.
    """
    xn = F.normalize(x.float(), p=2, dim=-1, eps=eps)
    sim = matmul(xn, xn.t(), autocast).clone()
    sim.fill_diagonal_(-2.0)
    idx = sim.argmax(dim=1)
    d = F.pairwise_distance(xn, xn[idx], p=2.0, eps=eps)
    return -(d + eps).log().mean()


# --------------------------------------------------------------------------------------------------
# schedules / EMA / optimizer
# --------------------------------------------------------------------------------------------------
def linear_warmup_schedule(step: int, warmup_steps: int, start_value: float, end_value: float) -> float:
    """LT/_methods/dinov2/scheduler.py:13-34."""
    if step < warmup_steps:
        return start_value + step / warmup_steps * (end_value - start_value)
    return end_value


def cosine_schedule(step: int, max_steps: int, start_value: float, end_value: float) -> float:
    """lightly.utils.scheduler.cosine_schedule (period=None branch), call sites dinov2.py:602-607,648-653."""
    if max_steps == 1:
        return end_value
    if step == max_steps:
        return end_value
    return end_value - (end_value - start_value) * (math.cos(math.pi * step / (max_steps - 1)) + 1) / 2


def cosine_warmup_lr_factor(step: int, warmup_steps: int, max_steps: int, end_value: float) -> float:
    """lightly.utils.scheduler.CosineWarmupScheduler.scale_lr (lightly 1.5.26), used at dinov2.py:576-583:
    linear warmup (step+1)/warmup for step < warmup, then cosine from 1.0 to end_value."""
    if warmup_steps > 0 and step < warmup_steps:
        return (step + 1) / warmup_steps
    return cosine_schedule(step - warmup_steps, max_steps - warmup_steps, 1.0, end_value)


def update_ema(params: Sequence[Tensor], params_ema: Sequence[Tensor], m: float) -> None:
    """LT/_torch_helpers.py:75-96: ema = ema*m + p*(1-m), in place."""
    with torch.no_grad():
        for p, e in zip(params, params_ema):
            e.mul_(m).add_(p.to(e.dtype), alpha=1.0 - m)


def vit_lr_decay_rate(name: str, lr_decay_rate: float, num_layers: int) -> float:
    """get_vit_lr_decay_rate, LT/_methods/dinov2/utils.py:155-186 (non-chunked blocks)."""
    layer_id = num_layers + 1
    if any(k in name for k in ("pos_embed", "patch_embed", "mask_token", "cls_token", "register_tokens")):
        layer_id = 0
    elif "blocks." in name and "residual." not in name:
        layer_id = int(name[name.find("blocks."):].split(".")[1]) + 1
    return lr_decay_rate ** (num_layers + 1 - layer_id)


def param_hparams(name: str, is_backbone: bool, base_lr: float, weight_decay: float, num_layers: int,
                  layerwise_decay: float = 0.9, patch_embed_lr_multiplier: float = 0.2) -> Dict[str, float]:
    """Per-parameter lr / weight-decay of get_optimizer_with_decay, utils.py:191-250."""
    decay = vit_lr_decay_rate(name, layerwise_decay, num_layers) if is_backbone else 1.0
    lr = base_lr * decay
    wd = weight_decay
    if name.endswith(".bias") or "norm" in name or "gamma" in name:
        wd = 0.0
    if "patch_embed" in name:
        lr = lr * patch_embed_lr_multiplier
    return {"lr": lr, "weight_decay": wd}


def clip_grad_norm(grads: Sequence[Tensor], max_norm: float) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ (Lightning clip_gradients 'norm', dinov2.py:588-598), in place."""
    total = torch.sqrt(sum((g.float() ** 2).sum() for g in grads))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, wd: float,
               beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.AdamW single-tensor update (LT/_optim/adamw_args.py:33-36), in place; step is 1-based."""
    p.mul_(1 - lr * wd)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------------------
# the training step (LT/_methods/dinov2/dinov2.py:259-519)
# --------------------------------------------------------------------------------------------------
@dataclass
class StepConfig:
    vit: ViTConfig = field(default_factory=ViTConfig)
    head: HeadConfig = field(default_factory=HeadConfig)
    ibot_separate_head: bool = False
    center_method: str = "softmax"
    student_temp: float = 0.1
    center_momentum: float = 0.9
    dino_loss_weight: float = 1.0
    ibot_loss_weight: float = 1.0
    koleo_loss_weight: float = 0.1


def _sub(sd: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def training_step(cfg: StepConfig, student: Dict[str, Tensor], teacher: Dict[str, Tensor],
                  centers: Dict[str, Tensor], views: List[Tensor], masks: Tensor, mask_indices: Tensor,
                  masks_weight: Tensor, teacher_temp: float, autocast: bool = False,
                  keep_scales_global: Optional[Sequence[Tensor]] = None,
                  keep_scales_local: Optional[Sequence[Tensor]] = None,
                  taps: Optional[dict] = None) -> Dict[str, Tensor]:
    """One DINOv2 loss evaluation (single rank, world_size 1).

    student / teacher: flat dicts with keys "backbone.<vit key>", "dino_head.<head key>" and (if
    ibot_separate_head) "ibot_head.<head key>".  centers: {"dino": [1,K], "ibot": [1,1,K]} -- the centers as
    they are when softmax_center_teacher runs (pending update already applied).
    Returns loss terms plus "dino_center_sum"/"ibot_center_sum" (the per-rank sums feeding the next center
    update) so callers can apply center_ema().
    """
    n_global = 2
    n_local = len(views) - n_global
    g_terms = (n_global - 1) * n_global
    l_terms = max(n_local * n_global, 1)
    gv = torch.cat(views[:n_global])
    B = gv.shape[0] // n_global
    M = mask_indices.shape[0]
    D = cfg.vit.embed_dim

    s_bb, t_bb = _sub(student, "backbone."), _sub(teacher, "backbone.")
    s_dino, t_dino = _sub(student, "dino_head."), _sub(teacher, "dino_head.")
    s_ibot = _sub(student, "ibot_head.") if cfg.ibot_separate_head else s_dino
    t_ibot = _sub(teacher, "ibot_head.") if cfg.ibot_separate_head else t_dino

    out: Dict[str, Tensor] = {}
    # ---- teacher (dinov2.py:399-472), no grad
    with torch.no_grad():
        tt = vit_forward_features(t_bb, cfg.vit, gv, None, autocast)
        t_cls = torch.cat([tt["cls"][B:], tt["cls"][:B]])  # swap global crops A<->B (:415-417)
        t_cls_logits = head_forward(t_dino, t_cls, autocast)
        t_patch = tt["patch"].reshape(-1, D).index_select(0, mask_indices)
        t_patch_logits = head_forward(t_ibot, t_patch, autocast)
        if cfg.center_method == "softmax":
            t_cls_probs = softmax_center_teacher(t_cls_logits, centers["dino"], teacher_temp)
            t_patch_probs = softmax_center_teacher(t_patch_logits.unsqueeze(0), centers["ibot"], teacher_temp).squeeze(0)
            out["dino_center_sum"] = center_batch_sum_dino(t_cls_logits)
            out["ibot_center_sum"] = center_batch_sum_ibot(t_patch_logits.unsqueeze(0))
        elif cfg.center_method == "sinkhorn_knopp":
            t_cls_probs = sinkhorn_knopp(t_cls_logits, teacher_temp, float(t_cls_logits.shape[0]))
            t_patch_probs = sinkhorn_knopp(t_patch_logits, teacher_temp, float(M))
        else:
            raise ValueError(cfg.center_method)
        t_cls_probs_g = t_cls_probs.view(2, B, -1)
    if taps is not None:
        taps.update(t_cls_logits=t_cls_logits, t_patch_logits=t_patch_logits, t_cls_probs=t_cls_probs,
                    t_patch_probs=t_patch_probs, t_cls=tt["cls"], t_patch=tt["patch"])

    # ---- student global (dinov2.py:474-505)
    sg = vit_forward_features(s_bb, cfg.vit, gv, masks, autocast, keep_scales_global)
    s_cls_g = sg["cls"]
    s_cls_logits_g = head_forward(s_dino, s_cls_g, autocast)
    s_patch = sg["patch"].reshape(-1, D).index_select(0, mask_indices)
    s_patch_logits = head_forward(s_ibot, s_patch, autocast)

    dino_global = dino_loss([s_cls_logits_g], [t_cls_probs_g.flatten(0, 1)], cfg.student_temp) * 2 / (g_terms + l_terms)

    # ---- student local (dinov2.py:507-519)
    dino_local = torch.zeros_like(dino_global)
    s_cls_logits_l = None
    if n_local > 0:
        lv = torch.cat(views[n_global:])
        sl = vit_forward_features(s_bb, cfg.vit, lv, None, autocast, keep_scales_local)
        s_cls_logits_l = head_forward(s_dino, sl["cls"], autocast)
        dino_local = dino_loss(s_cls_logits_l.chunk(n_local), list(t_cls_probs_g), cfg.student_temp) / (g_terms + l_terms)

    ibot = ibot_loss_masked(s_patch_logits, t_patch_probs, masks_weight, masks.shape[0], cfg.student_temp)
    koleo = sum(koleo_loss(c, autocast=autocast) for c in s_cls_g.chunk(2))

    loss = (cfg.dino_loss_weight * dino_global + cfg.dino_loss_weight * dino_local
            + cfg.ibot_loss_weight * ibot + cfg.koleo_loss_weight * koleo)
    out.update(loss=loss, dino_global_loss=dino_global, dino_local_loss=dino_local, ibot_loss=ibot, koleo_loss=koleo)
    if taps is not None:
        taps.update(s_cls_logits_g=s_cls_logits_g, s_patch_logits=s_patch_logits, s_cls_logits_l=s_cls_logits_l,
                    s_cls_g=s_cls_g, s_patch=sg["patch"])
    return out
