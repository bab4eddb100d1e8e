"""DinoVisionTransformer on B200 kernels.

Mirror of LT/_models/dinov2_vit/dinov2_vit_src/models/vision_transformer.py:83-487 (class DinoVisionTransformer,
non-chunked blocks, ffn_layer "mlp" or "swiglu"/"swiglufused"): same constructor arguments, same parameter names / shapes
(`cls_token`, `pos_embed`, `mask_token`, `register_tokens`, `patch_embed.proj.*`, `blocks.{i}.{norm1,attn.qkv,
attn.proj,ls1,norm2,mlp.fc1,mlp.fc2,ls2}.*`, `norm.*`) so reference checkpoints load unchanged, and the same
`forward_features` contract.  The arithmetic runs on the sm_100a kernels of libb200dino.so; forward and
backward are explicit schedules of kernel launches (no autograd graph): `_fwd` returns a context of saved
activations, `_bwd` consumes it and accumulates into the gradient arena.

Rounding points follow torch.autocast("cuda", bfloat16): bf16 GEMM operands/outputs
with fp32 accumulation, fp32 residual stream / LayerNorm / softmax.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops
from .._arena import Arena
from .pos_embed import pos_embed_operator

# LayerNorm as the A-operand prologue of the following GEMM (b200_ln_gemm; embed_dim <= 384):
#   "0" never | "1" whenever supported | "auto" only when every CTA gets at most one 128-row panel (a second panel's
#   normalisation cannot overlap the first one's MMAs: the panel is single-buffered)
LN_GEMM = os.environ.get("B200_LN_GEMM", "0")
_NUM_SMS: Dict[str, int] = {}


def _ln_gemm_wanted(T: int, D: int, dev: torch.device) -> bool:
    if LN_GEMM == "0" or D % 64 != 0 or D > ops.LN_GEMM_MAX_K or dev.type != "cuda":
        return False
    if LN_GEMM == "1":
        return True
    key = str(dev)
    if key not in _NUM_SMS:
        _NUM_SMS[key] = torch.cuda.get_device_properties(dev).multi_processor_count
    return (T + 127) // 128 <= _NUM_SMS[key]


def vit_param_shapes(embed_dim: int, depth: int, patch_size: int, in_chans: int, num_patches: int, hidden: int,
                     num_register_tokens: int, layerscale: bool, swiglu: bool = False) -> Dict[str, Tuple[int, ...]]:
    D, p = embed_dim, patch_size
    s: Dict[str, Tuple[int, ...]] = {"cls_token": (1, 1, D), "pos_embed": (1, 1 + num_patches, D)}
    if num_register_tokens:
        s["register_tokens"] = (1, num_register_tokens, D)
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
        if swiglu:  # SwiGLUFFN: w12 = Linear(D, 2H), w3 = Linear(H, D)  (layers/swiglu_ffn.py:28-29)
            s[b + "mlp.w12.weight"] = (2 * hidden, D); s[b + "mlp.w12.bias"] = (2 * hidden,)
            s[b + "mlp.w3.weight"] = (D, hidden); s[b + "mlp.w3.bias"] = (D,)
        else:
            s[b + "mlp.fc1.weight"] = (hidden, D); s[b + "mlp.fc1.bias"] = (hidden,)
            s[b + "mlp.fc2.weight"] = (D, hidden); s[b + "mlp.fc2.bias"] = (D,)
        if layerscale:
            s[b + "ls2.gamma"] = (D,)
    s["norm.weight"] = (D,); s["norm.bias"] = (D,)
    s["mask_token"] = (1, D)
    return s


def attach_params(root: nn.Module, arena: Arena, prefix: str, names, requires_grad: bool) -> None:
    """Register arena views as nn.Parameters under `root` following the dotted reference names."""
    for full in names:
        if not full.startswith(prefix):
            continue
        name = full[len(prefix):]
        mod = root
        parts = name.split(".")
        for part in parts[:-1]:
            if not hasattr(mod, part) or not isinstance(getattr(mod, part), nn.Module):
                mod.add_module(part, nn.Module())
            mod = getattr(mod, part)
        param = nn.Parameter(arena.p(full), requires_grad=requires_grad)
        if requires_grad and arena.grad is not None:
            param.grad = arena.g(full)
        mod.register_parameter(parts[-1], param)


def _unchunk_block_names(state_dict, prefix, *args) -> None:
    """load_state_dict pre-hook: `blocks.{chunk}.{i}.x` -> `blocks.{i}.x` (vision_transformer.py:215-226 keeps the global
    block index inside each chunk, so dropping the chunk index is the whole mapping)."""
    for k in list(state_dict.keys()):
        if not k.startswith(prefix + "blocks."):
            continue
        parts = k[len(prefix):].split(".")
        if len(parts) > 3 and parts[1].isdigit() and parts[2].isdigit():
            state_dict[prefix + ".".join(parts[:1] + parts[2:])] = state_dict.pop(k)


class VitCtx:
    """Saved activations of one forward pass (everything the explicit backward needs)."""

    def __init__(self) -> None:
        self.blocks: List[dict] = []


class DinoVisionTransformer(nn.Module):
    def __init__(self, img_size: int = 224, patch_size: int = 16, in_chans: int = 3, embed_dim: int = 768,
                 depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0, qkv_bias: bool = True,
                 ffn_bias: bool = True, proj_bias: bool = True, drop_path_rate: float = 0.0,
                 drop_path_uniform: bool = False, init_values: Optional[float] = None, ffn_layer: str = "mlp",
                 block_chunks: int = 0, num_register_tokens: int = 0, interpolate_antialias: bool = False,
                 interpolate_offset: float = 0.1, *, arena: Optional[Arena] = None, prefix: str = "",
                 device: str = "cuda", requires_grad: bool = True) -> None:
        super().__init__()
        if ffn_layer not in ("mlp", "swiglu", "swiglufused") or block_chunks not in (0,) or not (qkv_bias and ffn_bias and proj_bias):
            raise NotImplementedError("b200 DinoVisionTransformer: ffn_layer in {mlp, swiglu, swiglufused}, block_chunks=0, biases on")
        self.swiglu = ffn_layer != "mlp"
        # parameter names of the FFN's input / output projections
        self._ffn_in, self._ffn_out = ("mlp.w12.", "mlp.w3.") if self.swiglu else ("mlp.fc1.", "mlp.fc2.")
        if embed_dim % num_heads or embed_dim // num_heads != 64:
            raise NotImplementedError("b200 attention kernels are specialised for head_dim 64")
        self.num_features = self.embed_dim = embed_dim
        self.n_blocks = depth
        self.num_heads = num_heads
        self.patch_size = patch_size
        self.in_chans = in_chans
        self.img_size = img_size
        self.num_register_tokens = num_register_tokens
        self.interpolate_antialias = interpolate_antialias
        self.interpolate_offset = interpolate_offset
        self.hidden_dim = int(embed_dim * mlp_ratio)
        if self.swiglu:
            self.hidden_dim = (int(self.hidden_dim * 2 / 3) + 7) // 8 * 8  # SwiGLUFFNFused (layers/swiglu_ffn.py:60-63)
        self.chunked_blocks = False
        self.layerscale = bool(init_values)
        self.ln_eps = 1e-6
        self.num_patches = (img_size // patch_size) ** 2
        if drop_path_uniform:
            self.dpr = [drop_path_rate] * depth
        else:
            self.dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]  # vision_transformer.py:161-166
        self._dp_keep: dict = {}
        self._subset_rng: dict = {}  # device -> (seed, launch counter int64[1] on the device): batch-subset stochastic depth
        self.prefix = prefix
        shapes = vit_param_shapes(embed_dim, depth, patch_size, in_chans, self.num_patches, self.hidden_dim,
                                  num_register_tokens, self.layerscale, self.swiglu)
        if arena is None:
            arena = Arena({prefix + k: v for k, v in shapes.items()}, device, with_grad=requires_grad,
                          with_optim_state=False)
        self.arena = arena
        attach_params(self, arena, prefix, [prefix + k for k in shapes], requires_grad)
        self._pos_ops: Dict[Tuple[int, int], Tensor] = {}
        # configured post-instantiation, like DINOv2ViTModelWrapper.set_activation_checkpointing (dinov2_vit.py:55-59)
        self._activation_checkpointing = False
        self._activation_checkpointing_every_n_blocks = 1
        # batch-subset stochastic depth: run the residual branches on the kept samples only (True) or on all samples with the
        # dropped ones scaled by zero (False: the dense statement of the same arithmetic)
        self.subset_skips_compute = True
        # below this rate the gather / write-back copies and the un-fused LayerNorm backward cost more than the skipped rows
        # save (cfg2: the last block's rate is float32(0.1) > 0.1, i.e. a subset block that would skip 10 % of one block)
        self.subset_compact_min_rate = float(os.environ.get("B200_SUBSET_MIN_RATE", "0.15"))
        self.init_weights(init_values)
        # checkpoints written with block_chunks > 0 (zoo ViT-L/g configs) name blocks `blocks.{chunk}.{i}.*`
        self._register_load_state_dict_pre_hook(_unchunk_block_names)
        # weights land in the fp32 arena: the bf16 GEMM shadow is stale after any load
        self.register_load_state_dict_post_hook(lambda mod, incompatible: setattr(mod.arena, "bf16_valid", False))

    # ------------------------------------------------------------------ init (vision_transformer.py:244-249,574+)
    @torch.no_grad()
    def init_weights(self, init_values: Optional[float]) -> None:
        for name, prm in self.named_parameters():
            if name == "pos_embed":
                nn.init.trunc_normal_(prm, std=0.02)
            elif name in ("cls_token", "register_tokens"):
                nn.init.normal_(prm, std=1e-6)
            elif name == "mask_token":
                prm.zero_()
            elif name == "patch_embed.proj.weight":
                nn.init.kaiming_uniform_(prm, a=math.sqrt(5))  # nn.Conv2d default
            elif name == "patch_embed.proj.bias":
                bound = 1.0 / math.sqrt(self.in_chans * self.patch_size ** 2)
                nn.init.uniform_(prm, -bound, bound)
            elif name.endswith("gamma"):
                prm.fill_(init_values)
            elif "norm" in name and name.endswith("weight"):
                prm.fill_(1.0)
            elif name.endswith("bias"):
                prm.zero_()
            else:
                nn.init.trunc_normal_(prm, std=0.02)
        self.arena.bf16_valid = False

    # ------------------------------------------------------------------ helpers
    def _P(self, name: str) -> Tensor:
        return self.arena.p(self.prefix + name)

    def _W(self, name: str) -> Tensor:
        return self.arena.w(self.prefix + name)

    def _G(self, name: str) -> Tensor:
        return self.arena.g(self.prefix + name)

    def _pos_operator(self, w0: int, h0: int) -> Tensor:
        key = (w0, h0)
        if key not in self._pos_ops:
            M = int(math.sqrt(self.num_patches))
            W = pos_embed_operator(M, w0, h0, self.interpolate_offset, self.interpolate_antialias)
            self._pos_ops[key] = torch.from_numpy(W).to(self.arena.device)
        return self._pos_ops[key]

    def _pos_for(self, Np: int, H: int, W: int) -> Tuple[Tensor, bool]:
        """[1+Np, D] positional embedding for this crop size (interpolate_pos_encoding, :251-305)."""
        pos = self._P("pos_embed")[0]
        if Np == self.num_patches and H == W:
            return pos, False
        w0, h0 = H // self.patch_size, W // self.patch_size  # reference names (new_H, new_W) as (w, h)
        op = self._pos_operator(w0, h0)
        out = torch.empty(1 + Np, self.embed_dim, device=pos.device, dtype=torch.float32)
        out[0].copy_(pos[0])
        ops.small_matmul(op, pos[1:], out[1:])
        return out, True

    # ------------------------------------------------------------------ one block forward (Block.forward, block.py:90-115)
    def _ln_linear(self, x: Tensor, norm: str, wname: str, bname: str, out: Tensor, save: bool, epi: int = ops.EPI_BF16,
                   out2: Optional[Tensor] = None):
        """out = epi(Linear(LayerNorm(x))) -- norm1 -> attn.qkv and norm2 -> mlp.fc1 / w12 of Block.forward (block.py:90-115).
        With B200_LN_GEMM=1 and embed_dim <= 384 the LayerNorm runs as the GEMM's A-operand prologue (b200_ln_gemm): the
        normalised rows go straight into the MMA's shared-memory tiles; the bf16 copy + mean / rstd the backward needs are
        side outputs (only when `save`).  Returns (xn | None, mean | None, rstd | None)."""
        T, D = x.shape
        dev = x.device
        mean, rstd = (torch.empty(T, device=dev, dtype=torch.float32), torch.empty(T, device=dev, dtype=torch.float32)) if save \
            else (None, None)
        w, b = self._P(norm + "weight"), self._P(norm + "bias")
        if _ln_gemm_wanted(T, D, dev):
            xn = torch.empty(T, D, device=dev, dtype=torch.bfloat16) if save else None
            ops.ln_gemm(x, w, b, self.ln_eps, self._W(wname), out, epi=epi, bias=self._P(bname), out2=out2, xn_out=xn, mean=mean,
                        rstd=rstd)
            return xn, mean, rstd
        xn = torch.empty(T, D, device=dev, dtype=torch.bfloat16)
        ops.layernorm_fwd(x, w, b, self.ln_eps, xn, mean, rstd)
        ops.gemm(xn, self._W(wname), out, epi=epi, bias=self._P(bname), out2=out2)
        return xn, mean, rstd

    def _block_fwd(self, i: int, xcur: Tensor, Bc: int, N: int, rs1: Optional[Tensor], rs2: Optional[Tensor], save: bool) -> dict:
        dev = xcur.device
        D, Hd, h = self.embed_dim, self.hidden_dim, self.num_heads
        T = xcur.shape[0]
        bf, f32 = torch.bfloat16, torch.float32
        E = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        b = f"blocks.{i}."
        scale = 64 ** -0.5
        qkv = E(T, 3 * D)
        xn, mean1, rstd1 = self._ln_linear(xcur, b + "norm1.", b + "attn.qkv.weight", b + "attn.qkv.bias", qkv, save)
        att = E(T, D)
        lse = E(Bc * h, N, dt=f32) if save else None
        ops.attention_fwd(qkv, Bc, N, h, att, lse, scale)
        o1 = E(T, D) if save else None
        xmid = E(T, D, dt=f32)
        ops.gemm(att, self._W(b + "attn.proj.weight"), xmid, epi=ops.EPI_RESIDUAL, bias=self._P(b + "attn.proj.bias"),
                 out2=o1, aux=xcur, gamma=self._P(b + "ls1.gamma") if self.layerscale else None,
                 rowscale=rs1, rows_per_scale=N)
        hh = E(T, Hd)
        if self.swiglu:
            u = E(T, 2 * Hd)  # x12 = w12(x): kept for the backward when saving
            xn2, mean2, rstd2 = self._ln_linear(xmid, b + "norm2.", b + "mlp.w12.weight", b + "mlp.w12.bias", u, save)
            ops.swiglu_fwd(u, hh)
            if not save:
                u = None
        else:
            u = E(T, Hd) if save else None  # u holds gelu'(fc1 out) for the backward
            xn2, mean2, rstd2 = self._ln_linear(xmid, b + "norm2.", b + "mlp.fc1.weight", b + "mlp.fc1.bias", hh, save,
                                                epi=ops.EPI_BIAS_GELU_DG if save else ops.EPI_BIAS_GELU, out2=u)
        o2 = E(T, D) if save else None
        xout = E(T, D, dt=f32)
        ops.gemm(hh, self._W(b + self._ffn_out + "weight"), xout, epi=ops.EPI_RESIDUAL, bias=self._P(b + self._ffn_out + "bias"),
                 out2=o2, aux=xmid, gamma=self._P(b + "ls2.gamma") if self.layerscale else None,
                 rowscale=rs2, rows_per_scale=N)
        sv = {"x_out": xout}
        if save:
            sv.update(x_in=xcur, mean1=mean1, rstd1=rstd1, xn=xn, qkv=qkv, att=att, lse=lse, o1=o1, x_mid=xmid,
                      mean2=mean2, rstd2=rstd2, xn2=xn2, u=u, h=hh, o2=o2, rs1=rs1, rs2=rs2)
        return sv

    # ------------------------------------------------------------------ batch-subset stochastic depth (block.py:118-141)
    def _branch_fwd(self, i: int, which: int, xs: Tensor, bsub: int, N: int, scale_vec: Tensor, save: bool) -> dict:
        """One residual branch of block i on a COMPACT subset [bsub*N, D] of the residual stream:
        which=0: LN1 -> qkv -> attention -> proj ; which=1: LN2 -> FFN.  Returns xo = xs + (b/b') * LayerScale(branch)."""
        dev = xs.device
        D, Hd, h = self.embed_dim, self.hidden_dim, self.num_heads
        Ts = xs.shape[0]
        bf, f32 = torch.bfloat16, torch.float32
        E = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        b = f"blocks.{i}."
        mean, rstd = (E(Ts, dt=f32), E(Ts, dt=f32)) if save else (None, None)
        xn = E(Ts, D)
        o = E(Ts, D) if save else None
        xo = E(Ts, D, dt=f32)
        if which == 0:
            ops.layernorm_fwd(xs, self._P(b + "norm1.weight"), self._P(b + "norm1.bias"), self.ln_eps, xn, mean, rstd)
            qkv = E(Ts, 3 * D)
            ops.gemm(xn, self._W(b + "attn.qkv.weight"), qkv, bias=self._P(b + "attn.qkv.bias"))
            att = E(Ts, D)
            lse = E(bsub * h, N, dt=f32) if save else None
            ops.attention_fwd(qkv, bsub, N, h, att, lse, 64 ** -0.5)
            ops.gemm(att, self._W(b + "attn.proj.weight"), xo, epi=ops.EPI_RESIDUAL, bias=self._P(b + "attn.proj.bias"), out2=o,
                     aux=xs, gamma=self._P(b + "ls1.gamma") if self.layerscale else None, rowscale=scale_vec, rows_per_scale=N)
            sv = dict(xs=xs, mean=mean, rstd=rstd, xn=xn, qkv=qkv, att=att, lse=lse, o=o) if save else {}
        else:
            ops.layernorm_fwd(xs, self._P(b + "norm2.weight"), self._P(b + "norm2.bias"), self.ln_eps, xn, mean, rstd)
            hh = E(Ts, Hd)
            if self.swiglu:
                u = E(Ts, 2 * Hd)
                ops.gemm(xn, self._W(b + "mlp.w12.weight"), u, bias=self._P(b + "mlp.w12.bias"))
                ops.swiglu_fwd(u, hh)
            else:
                u = E(Ts, Hd) if save else None
                ops.gemm(xn, self._W(b + "mlp.fc1.weight"), hh, epi=ops.EPI_BIAS_GELU_DG if save else ops.EPI_BIAS_GELU,
                         bias=self._P(b + "mlp.fc1.bias"), out2=u)
            ops.gemm(hh, self._W(b + self._ffn_out + "weight"), xo, epi=ops.EPI_RESIDUAL, bias=self._P(b + self._ffn_out + "bias"),
                     out2=o, aux=xs, gamma=self._P(b + "ls2.gamma") if self.layerscale else None, rowscale=scale_vec, rows_per_scale=N)
            sv = dict(xs=xs, mean=mean, rstd=rstd, xn=xn, u=u, h=hh, o=o) if save else {}
        sv["xo"] = xo
        return sv

    def _block_fwd_subset(self, i: int, xcur: Tensor, Bc: int, N: int, idx1: Tensor, idx2: Tensor, save: bool, ckpt: bool) -> dict:
        """Block i with both residual branches evaluated on random subsets of b' = max(int(b (1 - r)), 1) samples and added
        back with alpha = b / b' (drop_add_residual_stochastic_depth): the dropped samples' branch is NOT computed.
        The residual stream `xcur` [Bc*N, D] is updated IN PLACE (only the subset rows change); what the backward needs of
        the old values are the compact copies made here."""
        D = self.embed_dim
        dev = xcur.device
        f32 = torch.float32
        x3 = xcur.view(Bc, N, D)
        out = {"subset": True, "idx": (idx1, idx2), "branches": [None, None]}
        for which, idx in ((0, idx1), (1, idx2)):
            bsub = idx.numel()
            xs = torch.empty(bsub * N, D, device=dev, dtype=f32)
            ops.gather_samples(x3, idx, xs.view(bsub, N, D))
            scale_vec = torch.full((bsub,), Bc / bsub, device=dev, dtype=f32)
            if save and ckpt:  # keep only the compact branch input; the branch is recomputed in the backward
                br = self._branch_fwd(i, which, xs, bsub, N, scale_vec, False)
                out["branches"][which] = {"ckpt": True, "xs": xs, "scale_vec": scale_vec}
            else:
                br = self._branch_fwd(i, which, xs, bsub, N, scale_vec, save)
                if save:
                    br["scale_vec"] = scale_vec
                    out["branches"][which] = br
            ops.scatter_samples(br.pop("xo").view(bsub, N, D), idx, x3)
        # dense per-sample scales (introspection / tests): b/b' on the subset, 0 elsewhere
        if save:
            out["rs1"] = torch.zeros(Bc, device=dev, dtype=f32).index_fill_(0, idx1, Bc / idx1.numel())
            out["rs2"] = torch.zeros(Bc, device=dev, dtype=f32).index_fill_(0, idx2, Bc / idx2.numel())
        return out

    def _block_bwd_subset(self, i: int, sv: dict, dx: Tensor, Bc: int, N: int, wgrad) -> None:
        """Backward of a subset block: dx [Bc*N, D] (gradient wrt the block output) becomes the gradient wrt the block input,
        in place: for each branch (MLP first) the subset rows of dx are copied out, the compact branch backward adds the
        branch's input gradient to the copy (identity path + branch path), and the rows are written back."""
        D, Hd, h = self.embed_dim, self.hidden_dim, self.num_heads
        dev = dx.device
        bf, f32 = torch.bfloat16, torch.float32
        E = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        ls = self.layerscale
        b = f"blocks.{i}."
        dx3 = dx.view(Bc, N, D)
        for which in (1, 0):
            idx = sv["idx"][which]
            bsub = idx.numel()
            br = sv["branches"][which]
            if br.get("ckpt"):
                sc = br["scale_vec"]
                br = self._branch_fwd(i, which, br["xs"], bsub, N, sc, True)
                br["scale_vec"] = sc
                br.pop("xo")
            Ts = bsub * N
            ddx = torch.empty(Ts, D, device=dev, dtype=f32)
            ops.gather_samples(dx3, idx, ddx.view(bsub, N, D))
            do = E(Ts, D)
            if which == 1:
                ops.layerscale_bwd(ddx, br["o"], self._P(b + "ls2.gamma") if ls else None, br["scale_vec"], N, do,
                                   self._G(b + "ls2.gamma") if ls else None, self._G(b + self._ffn_out + "bias"))
                if self.swiglu:
                    dH = E(Ts, Hd)
                    ops.gemm(do, self._W(b + "mlp.w3.weight"), dH, b_mn=True)
                    dU = E(Ts, 2 * Hd)
                    ops.swiglu_bwd(br["u"], dH, dU)
                else:
                    dU = E(Ts, Hd)
                    ops.gemm(do, self._W(b + "mlp.fc2.weight"), dU, b_mn=True, epi=ops.EPI_MUL_AUX, aux=br["u"])
                wgrad(do, br["h"], b + self._ffn_out + "weight")
                ops.col_reduce(dU, self._G(b + self._ffn_in + "bias"))
                wgrad(dU, br["xn"], b + self._ffn_in + "weight")
                dxn = E(Ts, D)
                ops.gemm(dU, self._W(b + self._ffn_in + "weight"), dxn, b_mn=True)
                ops.layernorm_bwd(dxn, br["xs"], self._P(b + "norm2.weight"), br["mean"], br["rstd"], ddx, True,
                                  self._G(b + "norm2.weight"), self._G(b + "norm2.bias"))
            else:
                ops.layerscale_bwd(ddx, br["o"], self._P(b + "ls1.gamma") if ls else None, br["scale_vec"], N, do,
                                   self._G(b + "ls1.gamma") if ls else None, self._G(b + "attn.proj.bias"))
                datt = E(Ts, D)
                ops.gemm(do, self._W(b + "attn.proj.weight"), datt, b_mn=True)
                wgrad(do, br["att"], b + "attn.proj.weight")
                dqkv = E(Ts, 3 * D)
                ops.attention_bwd(br["qkv"], br["att"], datt, br["lse"], bsub, N, h, dqkv, 64 ** -0.5,
                                  colsum=self._G(b + "attn.qkv.bias").view(-1))
                wgrad(dqkv, br["xn"], b + "attn.qkv.weight")
                dxn = E(Ts, D)
                ops.gemm(dqkv, self._W(b + "attn.qkv.weight"), dxn, b_mn=True)
                ops.layernorm_bwd(dxn, br["xs"], self._P(b + "norm1.weight"), br["mean"], br["rstd"], ddx, True,
                                  self._G(b + "norm1.weight"), self._G(b + "norm1.bias"))
            ops.scatter_samples(ddx.view(bsub, N, D), idx, dx3)

    # ------------------------------------------------------------------ forward
    def _fwd(self, x: Tensor, masks: Optional[Tensor], save: bool, drop_path: bool = False,
             keep_scales: Optional[List[Tensor]] = None) -> VitCtx:
        if not self.arena.bf16_valid:
            self.arena.refresh_bf16()
        dev = x.device
        Bc, Cin, H, Wimg = x.shape
        p, D, Hd, h = self.patch_size, self.embed_dim, self.hidden_dim, self.num_heads
        Np = (H // p) * (Wimg // p)
        R = self.num_register_tokens
        N = 1 + R + Np
        T = Bc * N
        bf, f32 = torch.bfloat16, torch.float32
        ctx = VitCtx()
        ctx.dims = (Bc, Np, R, N, T, H, Wimg)
        E = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731

        # conv-as-GEMM K = Cin*p*p; TMA needs 16-byte row pitches, so K is padded to a multiple of 8 (patch 14: 588 -> 592)
        # with zero columns in `cols` and zero columns in a padded bf16 copy of the weight shadow
        Kc = Cin * p * p
        Kp = (Kc + 7) // 8 * 8
        if Kp != Kc:
            cols = torch.zeros(Bc * Np, Kp, device=dev, dtype=bf)
            wpe = torch.zeros(D, Kp, device=dev, dtype=bf)
            wpe[:, :Kc].copy_(self._W("patch_embed.proj.weight").view(D, Kc))
        else:
            cols = E(Bc * Np, Kc)
            wpe = self._W("patch_embed.proj.weight").view(D, Kc)
        ops.im2col(x.contiguous(), p, cols)
        tok = E(Bc * Np, D)
        ops.gemm(cols, wpe, tok, bias=self._P("patch_embed.proj.bias"))
        pos, interp = self._pos_for(Np, H, Wimg)
        masks_u8 = masks.to(torch.uint8).contiguous() if masks is not None else None
        xs = E(Bc, N, D, dt=f32)
        ops.assemble_tokens(tok, masks_u8, self._P("mask_token") if masks is not None else None,
                            self._P("cls_token").view(D), self._P("register_tokens").view(R, D) if R else None, pos,
                            Bc, Np, R, D, xs)
        xcur = xs.view(T, D)
        if save:
            ctx.cols, ctx.masks_u8, ctx.interp = cols, masks_u8, interp
        # per-sample DropPath scales of every bernoulli block in four launches (rand, +keep, floor, /keep) instead of
        # four per block: bernoulli(keep)/keep == floor(u + keep)/keep for u ~ U[0,1)  (drop_path.py:23-27)
        bern = [i for i in range(self.n_blocks) if 0.0 < self.dpr[i] <= 0.1] if (drop_path and keep_scales is None) else []
        bern_scales = None
        if bern:
            key = str(dev)
            if key not in self._dp_keep:
                self._dp_keep[key] = torch.tensor([1.0 - self.dpr[i] for i in bern for _ in range(2)], device=dev,
                                                  dtype=f32).view(-1, 1)
            keep2 = self._dp_keep[key]
            bern_scales = torch.rand(2 * len(bern), Bc, device=dev, dtype=f32).add_(keep2).floor_().div_(keep2)
        for i in range(self.n_blocks):
            rs1 = rs2 = None
            if keep_scales is not None:
                rs1, rs2 = keep_scales[i][0].contiguous(), keep_scales[i][1].contiguous()
            elif bern_scales is not None and 0.0 < self.dpr[i] <= 0.1:
                j = bern.index(i)
                rs1, rs2 = bern_scales[2 * j], bern_scales[2 * j + 1]
            elif drop_path and self.dpr[i] > 0.1:
                # drop_add_residual_stochastic_depth (layers/block.py:118-141): the branch runs on a random subset of
                # b' = max(int(b*(1-r)), 1) samples and is added back with alpha = b/b'.
                bsub = max(int(Bc * (1.0 - self.dpr[i])), 1)
                # (random subset = the bsub smallest of Bc random keys: one small launch, graph-capturable -- the device-side
                # counter advances on every replay -- unlike randperm / index_put)
                idx1, idx2 = self._random_subset(Bc, bsub, dev), self._random_subset(Bc, bsub, dev)
                if self.subset_skips_compute and self.dpr[i] >= self.subset_compact_min_rate:
                    # compact schedule: only the subset's rows go through the branch (20-30 % of the block FLOPs saved)
                    ckpt = save and self._activation_checkpointing and (i % self._activation_checkpointing_every_n_blocks == 0)
                    sv = self._block_fwd_subset(i, xcur, Bc, N, idx1, idx2, save, ckpt)
                    if save:
                        ctx.blocks.append(sv)
                    continue
                # dense statement of the same arithmetic: per-sample scale b/b' on the subset, 0 elsewhere (every sample's
                # branch output is computed, the dropped ones are then multiplied by zero)
                rs1 = torch.zeros(Bc, device=dev, dtype=f32).index_fill_(0, idx1, Bc / bsub)
                rs2 = torch.zeros(Bc, device=dev, dtype=f32).index_fill_(0, idx2, Bc / bsub)
            elif drop_path and self.dpr[i] > 0.0:
                keep = 1.0 - self.dpr[i]
                rs1 = torch.empty(Bc, device=dev, dtype=f32).bernoulli_(keep).div_(keep)  # drop_path.py:23-27
                rs2 = torch.empty(Bc, device=dev, dtype=f32).bernoulli_(keep).div_(keep)
            ckpt = save and self._activation_checkpointing and (i % self._activation_checkpointing_every_n_blocks == 0)
            sv = self._block_fwd(i, xcur, Bc, N, rs1, rs2, save and not ckpt)
            xout = sv.pop("x_out")
            if save:
                if ckpt:  # keep only what the recomputation needs (LT/_activation_checkpointing.py:43-73)
                    sv = {"ckpt": True, "x_in": xcur, "rs1": rs1, "rs2": rs2}
                ctx.blocks.append(sv)
            xcur = xout
        meanf, rstdf = (E(T, dt=f32), E(T, dt=f32)) if save else (None, None)
        xnorm = E(T, D, dt=f32)
        ops.layernorm_fwd(xcur, self._P("norm.weight"), self._P("norm.bias"), self.ln_eps, xnorm, meanf, rstdf)
        ctx.x_prenorm, ctx.xnorm = xcur, xnorm
        if save:
            ctx.meanf, ctx.rstdf = meanf, rstdf
        return ctx

    def _random_subset(self, n: int, k: int, dev: torch.device) -> Tensor:
        """int64 [k]: a uniformly random k-subset of range(n) (torch.randperm(n)[:k] of layers/block.py:125-127).  The RNG
        state (seed drawn once from torch's CPU generator, so torch.manual_seed reproduces it; a launch counter in device
        memory) is created on first use, which must not be inside a CUDA-graph capture (the step's eager warm-up comes first)."""
        if n > 2048:
            return torch.rand(n, device=dev).argsort()[:k]
        key = str(dev)
        st = self._subset_rng.get(key)
        if st is None:
            if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
                raise RuntimeError("stochastic-depth RNG state must be created before CUDA-graph capture (run one eager step first)")
            rank = torch.distributed.get_rank() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 0
            seed = (int(torch.randint(0, 2 ** 62, (1,)).item()) ^ (rank * 0x9E3779B97F4A7C15)) & ((1 << 62) - 1)
            st = self._subset_rng[key] = (seed, torch.zeros(1, device=dev, dtype=torch.int64))
        idx = torch.empty(k, device=dev, dtype=torch.int64)
        ops.random_subset(n, k, st[0], st[1], idx)
        return idx

    # ------------------------------------------------------------------ backward
    def _bwd(self, ctx: VitCtx, d_xnorm: Optional[Tensor], wgrad_splits: int = 0, stop_before: int = 0,
             state: Optional[dict] = None) -> Optional[dict]:
        """d_xnorm: f32 [T, D] gradient wrt the final-LayerNorm output. Accumulates into the gradient arena.

        The schedule can be cut between blocks: `stop_before=k` (k > 0) runs the final LayerNorm and blocks nb-1 .. k and
        returns a state; a second call with `state=` runs blocks k-1 .. 0 and the embeddings.  After the first segment of
        the LAST backward pass of a step the gradients of blocks >= k and of `norm` are final, so their all-reduce can
        overlap the second segment (DINOv2._core_b1 / _core_b2)."""
        Bc, Np, R, N, T, H, Wimg = ctx.dims
        D, Hd, h = self.embed_dim, self.hidden_dim, self.num_heads
        dev = (d_xnorm if d_xnorm is not None else state["dx"]).device
        bf, f32 = torch.bfloat16, torch.float32
        E = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        if wgrad_splits < 0:
            wgrad_splits = 0  # 0 = let b200_gemm pick (tile width, split-K) for whole waves over the SMs
        scale = 64 ** -0.5

        def wgrad(dy: Tensor, xin: Tensor, name: str) -> None:
            # dW[out,in] += dy^T x : contraction over tokens, both operands MN-major, split-K + fp32 atomics
            ops.gemm(dy, xin, self._G(name).view(dy.shape[1], xin.shape[1]), a_mn=True, b_mn=True,
                     epi=ops.EPI_F32_ATOMIC, splits=wgrad_splits)

        ls = self.layerscale
        nb = self.n_blocks
        resume = state is not None
        dx = state["dx"] if resume else E(T, D, dt=f32)
        # final LayerNorm backward, fused with the LayerScale backward of the last block's MLP branch
        def materialise(j: int) -> dict:
            """Recompute the activations of a checkpointed block from its saved input (same DropPath scales)."""
            blk = ctx.blocks[j]
            if blk.get("ckpt"):
                blk = self._block_fwd(j, blk["x_in"], Bc, N, blk["rs1"], blk["rs2"], True)
                blk.pop("x_out")
                ctx.blocks[j] = blk
            return blk

        def is_subset(j: int) -> bool:
            return bool(ctx.blocks[j].get("subset"))

        if resume:
            do2 = state["do2"]
            hi = state["next"]
        elif is_subset(nb - 1):
            # batch-subset block: its branches run on compact row subsets, so nothing fuses across the block boundary
            ops.layernorm_bwd(d_xnorm, ctx.x_prenorm, self._P("norm.weight"), ctx.meanf, ctx.rstdf, dx, False,
                              self._G("norm.weight"), self._G("norm.bias"))
            do2 = None
            hi = nb - 1
        else:
            last = materialise(nb - 1)
            bl = f"blocks.{nb - 1}."
            do2 = E(T, D)
            ops.layernorm_bwd_ls(d_xnorm, ctx.x_prenorm, self._P("norm.weight"), ctx.meanf, ctx.rstdf, dx, False,
                                 self._G("norm.weight"), self._G("norm.bias"), last["o2"],
                                 self._P(bl + "ls2.gamma") if ls else None, last["rs2"], N, do2,
                                 self._G(bl + "ls2.gamma") if ls else None, self._G(bl + self._ffn_out + "bias"))
            hi = nb - 1
        for i in range(hi, stop_before - 1, -1):
            b = f"blocks.{i}."
            sv = ctx.blocks[i]
            if sv.get("subset"):
                self._block_bwd_subset(i, sv, dx, Bc, N, wgrad)
                do2 = None
                ctx.blocks[i] = None
                continue
            if do2 is None:
                # the block after this one (or the final LayerNorm) could not fuse this block's LayerScale backward
                sv = materialise(i)
                do2 = E(T, D)
                ops.layerscale_bwd(dx, sv["o2"], self._P(b + "ls2.gamma") if ls else None, sv["rs2"], N, do2,
                                   self._G(b + "ls2.gamma") if ls else None, self._G(b + self._ffn_out + "bias"))
            # ---- MLP branch (do2 = gradient of the fc2 output, produced by the fused kernel above / below)
            if self.swiglu:
                dH = E(T, Hd)
                ops.gemm(do2, self._W(b + "mlp.w3.weight"), dH, b_mn=True)
                dU = E(T, 2 * Hd)
                ops.swiglu_bwd(sv["u"], dH, dU)
            else:
                dU = E(T, Hd)
                ops.gemm(do2, self._W(b + "mlp.fc2.weight"), dU, b_mn=True, epi=ops.EPI_MUL_AUX, aux=sv["u"])
            wgrad(do2, sv["h"], b + self._ffn_out + "weight")
            ops.col_reduce(dU, self._G(b + self._ffn_in + "bias"))
            wgrad(dU, sv["xn2"], b + self._ffn_in + "weight")
            dxn2 = E(T, D)
            ops.gemm(dU, self._W(b + self._ffn_in + "weight"), dxn2, b_mn=True)
            # LN2 backward (dx += ...) fused with the LayerScale backward of this block's attention branch
            do1 = E(T, D)
            ops.layernorm_bwd_ls(dxn2, sv["x_mid"], self._P(b + "norm2.weight"), sv["mean2"], sv["rstd2"], dx, True,
                                 self._G(b + "norm2.weight"), self._G(b + "norm2.bias"), sv["o1"],
                                 self._P(b + "ls1.gamma") if ls else None, sv["rs1"], N, do1,
                                 self._G(b + "ls1.gamma") if ls else None, self._G(b + "attn.proj.bias"))
            # ---- attention branch
            datt = E(T, D)
            ops.gemm(do1, self._W(b + "attn.proj.weight"), datt, b_mn=True)
            wgrad(do1, sv["att"], b + "attn.proj.weight")
            dqkv = E(T, 3 * D)
            # the qkv bias gradient (column sums of dqkv) is accumulated by the attention backward from its registers
            ops.attention_bwd(sv["qkv"], sv["att"], datt, sv["lse"], Bc, N, h, dqkv, scale,
                              colsum=self._G(b + "attn.qkv.bias").view(-1))
            wgrad(dqkv, sv["xn"], b + "attn.qkv.weight")
            dxn = E(T, D)
            ops.gemm(dqkv, self._W(b + "attn.qkv.weight"), dxn, b_mn=True)
            if i > 0 and not is_subset(i - 1):
                # LN1 backward fused with the LayerScale backward of the PREVIOUS block's MLP branch
                pv = materialise(i - 1)
                bp = f"blocks.{i - 1}."
                do2 = E(T, D)
                ops.layernorm_bwd_ls(dxn, sv["x_in"], self._P(b + "norm1.weight"), sv["mean1"], sv["rstd1"], dx, True,
                                     self._G(b + "norm1.weight"), self._G(b + "norm1.bias"), pv["o2"],
                                     self._P(bp + "ls2.gamma") if ls else None, pv["rs2"], N, do2,
                                     self._G(bp + "ls2.gamma") if ls else None, self._G(bp + self._ffn_out + "bias"))
            else:
                ops.layernorm_bwd(dxn, sv["x_in"], self._P(b + "norm1.weight"), sv["mean1"], sv["rstd1"], dx, True,
                                  self._G(b + "norm1.weight"), self._G(b + "norm1.bias"))
                do2 = None
            ctx.blocks[i] = None  # free activations early
        if stop_before > 0:
            return {"dx": dx, "do2": do2, "next": stop_before - 1}
        # ---- embeddings
        dtok = E(Bc * Np, D)
        gpos = self._G("pos_embed")[0]
        if ctx.interp:
            dpos = torch.zeros(1 + Np, D, device=dev, dtype=f32)
        else:
            dpos = gpos
        ops.assemble_tokens_bwd(dx.view(Bc, N, D), ctx.masks_u8, Bc, Np, R, D, dtok, dpos, self._G("cls_token").view(D),
                                self._G("register_tokens").view(R, D) if R else None,
                                self._G("mask_token").view(D) if ctx.masks_u8 is not None else None)
        if ctx.interp:
            w0, h0 = H // self.patch_size, Wimg // self.patch_size
            ops.small_matmul(self._pos_operator(w0, h0), dpos[1:], gpos[1:], a_trans=True, accumulate=True)
            gpos[0].add_(dpos[0])
        ops.col_reduce(dtok, self._G("patch_embed.proj.bias"))
        gw = self._G("patch_embed.proj.weight").view(D, -1)
        if ctx.cols.shape[1] != gw.shape[1]:  # padded K (patch 14): wgrad into a padded scratch, fold the real columns back
            gpad = torch.zeros(D, ctx.cols.shape[1], device=dev, dtype=f32)
            ops.gemm(dtok, ctx.cols, gpad, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=0)
            gw.add_(gpad[:, :gw.shape[1]])
        else:
            ops.gemm(dtok, ctx.cols, gw, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=0)

    # ------------------------------------------------------------------ reference-facing API
    @torch.no_grad()
    def forward_features(self, x: Tensor, masks: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """DinoVisionTransformer.forward_features (vision_transformer.py:361-384); inference schedule (nothing saved)."""
        ctx = self._fwd(x, masks, save=False)
        Bc, Np, R, N, T, _, _ = ctx.dims
        xn = ctx.xnorm.view(Bc, N, self.embed_dim)
        return {"x_norm_clstoken": xn[:, 0], "x_norm_regtokens": xn[:, 1:R + 1], "x_norm_patchtokens": xn[:, R + 1:],
                "x_prenorm": ctx.x_prenorm.view(Bc, N, self.embed_dim), "masks": masks}

    def forward(self, *args, is_training: bool = False, **kwargs):
        ret = self.forward_features(*args, **kwargs)
        return ret if is_training else ret["x_norm_clstoken"]


def vit_small(patch_size=16, num_register_tokens=0, **kw) -> DinoVisionTransformer:
    return DinoVisionTransformer(patch_size=patch_size, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4,
                                 num_register_tokens=num_register_tokens, **kw)


def vit_base(patch_size=16, num_register_tokens=0, **kw) -> DinoVisionTransformer:
    return DinoVisionTransformer(patch_size=patch_size, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4,
                                 num_register_tokens=num_register_tokens, **kw)


def vit_large(patch_size=16, num_register_tokens=0, **kw) -> DinoVisionTransformer:
    return DinoVisionTransformer(patch_size=patch_size, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4,
                                 num_register_tokens=num_register_tokens, **kw)


def vit_tiny(patch_size=16, num_register_tokens=0, **kw) -> DinoVisionTransformer:
    """ViT-T/16 (BASELINE.json cfg1): embed 192, 3 heads of 64."""
    return DinoVisionTransformer(patch_size=patch_size, embed_dim=192, depth=12, num_heads=3, mlp_ratio=4,
                                 num_register_tokens=num_register_tokens, **kw)
