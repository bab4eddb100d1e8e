"""DINOv2ProjectionHead on B200 kernels.

Mirror of LT/_methods/dinov2/dinov2_head.py:32-95: Linear-GELU-Linear-GELU-Linear -> F.normalize ->
weight-normed Linear(bottleneck, out_dim, bias=False).  Same constructor arguments and parameter names
(`mlp.{0,2,4}.{weight,bias}`, `last_layer.parametrizations.weight.original{0,1}`).  `use_bn=True` and
`nlayers != 3` are not implemented on the B200 path (the reference defaults are False / 3).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor, nn

from ... import ops
from ..._arena import Arena
from ..._models.dinov2_vit import attach_params

G_NAME = "last_layer.parametrizations.weight.original0"
V_NAME = "last_layer.parametrizations.weight.original1"


def head_param_shapes(in_dim: int, hidden_dim: int, bottleneck_dim: int, out_dim: int) -> Dict[str, Tuple[int, ...]]:
    return {
        "mlp.0.weight": (hidden_dim, in_dim), "mlp.0.bias": (hidden_dim,),
        "mlp.2.weight": (hidden_dim, hidden_dim), "mlp.2.bias": (hidden_dim,),
        "mlp.4.weight": (bottleneck_dim, hidden_dim), "mlp.4.bias": (bottleneck_dim,),
        G_NAME: (out_dim, 1), V_NAME: (out_dim, bottleneck_dim),
    }


class HeadCtx:
    pass


class DINOv2ProjectionHead(nn.Module):
    def __init__(self, in_dim: int, out_dim: int, use_bn: bool = False, nlayers: int = 3, hidden_dim: int = 2048,
                 bottleneck_dim: int = 256, mlp_bias: bool = True, *, arena: Optional[Arena] = None, prefix: str = "",
                 device: str = "cuda", requires_grad: bool = True) -> None:
        super().__init__()
        if use_bn or nlayers != 3 or not mlp_bias:
            raise NotImplementedError("b200 DINOv2ProjectionHead: use_bn=False, nlayers=3, mlp_bias=True")
        self.in_dim, self.out_dim, self.hidden_dim, self.bottleneck_dim = in_dim, out_dim, hidden_dim, bottleneck_dim
        self.prefix = prefix
        shapes = head_param_shapes(in_dim, hidden_dim, bottleneck_dim, out_dim)
        if arena is None:
            arena = Arena({prefix + k: v for k, v in shapes.items()}, device, with_grad=requires_grad, with_optim_state=False)
        self.arena = arena
        attach_params(self, arena, prefix, [prefix + k for k in shapes], requires_grad)
        self.w_eff = torch.empty(out_dim, bottleneck_dim, device=arena.device, dtype=torch.bfloat16)
        self.init_weights()
        self.register_load_state_dict_post_hook(lambda mod, incompatible: setattr(mod.arena, "bf16_valid", False))

    @torch.no_grad()
    def init_weights(self) -> None:
        for name, prm in self.named_parameters():
            if name == G_NAME:
                prm.fill_(1.0)  # dinov2_head.py:58
            elif name == V_NAME:
                bound = 1.0 / self.bottleneck_dim ** 0.5  # nn.Linear default (kaiming_uniform a=sqrt(5))
                nn.init.uniform_(prm, -bound, bound)
            elif name.endswith("bias"):
                prm.zero_()
            else:
                nn.init.trunc_normal_(prm, std=0.02)
        self.arena.bf16_valid = False

    def _P(self, n: str) -> Tensor:
        return self.arena.p(self.prefix + n)

    def _W(self, n: str) -> Tensor:
        return self.arena.w(self.prefix + n)

    def _G(self, n: str) -> Tensor:
        return self.arena.g(self.prefix + n)

    def refresh_last_layer(self) -> None:
        """W_eff = g * v / ||v|| in bf16 (torch._weight_norm + autocast cast), once per step per head."""
        ops.weightnorm_fwd(self._P(G_NAME), self._P(V_NAME), self.w_eff)

    def _fwd(self, x_bf16: Tensor, save: bool, logits: Optional[Tensor] = None) -> HeadCtx:
        """x_bf16 [R, in_dim] -> ctx.logits bf16 [R, out_dim]."""
        if not self.arena.bf16_valid:
            self.arena.refresh_bf16()
        R = x_bf16.shape[0]
        dev = x_bf16.device
        E = lambda *s, dt=torch.bfloat16: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        ctx = HeadCtx()
        Hd, Bn, K = self.hidden_dim, self.bottleneck_dim, self.out_dim
        h0, u0 = E(R, Hd), (E(R, Hd) if save else None)
        ops.gemm(x_bf16, self._W("mlp.0.weight"), h0, epi=ops.EPI_BIAS_GELU_DG if save else ops.EPI_BIAS_GELU,
                 bias=self._P("mlp.0.bias"), out2=u0)  # u0/u1 hold gelu'(pre-activation) for the backward
        h1, u1 = E(R, Hd), (E(R, Hd) if save else None)
        ops.gemm(h0, self._W("mlp.2.weight"), h1, epi=ops.EPI_BIAS_GELU_DG if save else ops.EPI_BIAS_GELU,
                 bias=self._P("mlp.2.bias"), out2=u1)
        z = E(R, Bn)
        ops.gemm(h1, self._W("mlp.4.weight"), z, bias=self._P("mlp.4.bias"))
        zn, nrm = E(R, Bn), E(R, dt=torch.float32)
        ops.l2norm_fwd(z, zn, nrm, eps=1e-12)
        if logits is None:
            logits = E(R, K)
        ops.gemm(zn, self.w_eff, logits)
        ctx.logits = logits
        if save:
            ctx.x, ctx.u0, ctx.h0, ctx.u1, ctx.h1, ctx.z, ctx.zn, ctx.nrm = x_bf16, u0, h0, u1, h1, z, zn, nrm
        return ctx

    def _bwd(self, ctx: HeadCtx, dlogits: Tensor) -> Tensor:
        """dlogits bf16 [R, K] -> returns dx bf16 [R, in_dim]; accumulates parameter grads into the arena."""
        R, K = dlogits.shape
        dev = dlogits.device
        Hd, Bn = self.hidden_dim, self.bottleneck_dim
        E = lambda *s, dt=torch.bfloat16: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        splits = 0  # auto split-K
        # last layer: dW_eff = dlogits^T zn (fp32), then through the weight-norm parametrisation
        dW = E(K, Bn, dt=torch.float32)
        ops.gemm(dlogits, ctx.zn, dW, a_mn=True, b_mn=True, epi=ops.EPI_F32)
        ops.weightnorm_bwd(dW, self._P(G_NAME), self._P(V_NAME), self._G(G_NAME), self._G(V_NAME))
        # dzn = dlogits @ W_eff (contraction over K = out_dim): split-K into fp32, then round to bf16
        dzn32 = torch.zeros(R, Bn, device=dev, dtype=torch.float32)
        ops.gemm(dlogits, self.w_eff, dzn32, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=8)
        dzn = E(R, Bn)
        ops.cast_bf16(dzn32, dzn)
        dz = E(R, Bn)
        ops.l2norm_bwd(dzn, ctx.z, ctx.nrm, dz)

        def lin_bwd(dy: Tensor, xin: Tensor, wname: str, bname: str) -> None:
            ops.col_reduce(dy, self._G(bname))
            ops.gemm(dy, xin, self._G(wname), a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=splits)

        lin_bwd(dz, ctx.h1, "mlp.4.weight", "mlp.4.bias")
        dU1 = E(R, Hd)
        ops.gemm(dz, self._W("mlp.4.weight"), dU1, b_mn=True, epi=ops.EPI_MUL_AUX, aux=ctx.u1)
        lin_bwd(dU1, ctx.h0, "mlp.2.weight", "mlp.2.bias")
        dU0 = E(R, Hd)
        ops.gemm(dU1, self._W("mlp.2.weight"), dU0, b_mn=True, epi=ops.EPI_MUL_AUX, aux=ctx.u0)
        lin_bwd(dU0, ctx.x, "mlp.0.weight", "mlp.0.bias")
        dx = E(R, self.in_dim)
        ops.gemm(dU0, self._W("mlp.0.weight"), dx, b_mn=True)
        return dx

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        """Reference-facing call (dinov2_head.py:66-71): x [rows, in_dim] (any float dtype) -> bf16 logits."""
        self.refresh_last_layer()
        return self._fwd(x.to(torch.bfloat16).contiguous(), save=False).logits
