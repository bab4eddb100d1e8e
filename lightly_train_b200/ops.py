"""Thin tensor-level wrappers over the C-ABI (include/b200dino.h).

torch is used for device memory and streams only; every op here enqueues hand-written sm_100a kernels from
libb200dino.so on the current CUDA stream.  No op has a CPU / PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_BIAS_GELU, EPI_BIAS_GELU_DG, EPI_DGELU, EPI_F32, EPI_F32_ATOMIC,  # noqa: F401
                   EPI_MUL_AUX, EPI_RESIDUAL, GemmArgs, check)


GEMM_PROFILE = None  # bench.py sets this to a list: (flops, start_event, end_event) per b200_gemm launch
GEMM_PROFILE_KEYS = []
KERNEL_PROFILE = None  # bench.py sets this to a list: (kernel name, algorithmic bytes, start_event, end_event) per launch


def _timed(name: str, nbytes: float, launch) -> None:
    """Run `launch()`; when bench.py profiles, bracket it with a CUDA-event pair on the launch stream."""
    if KERNEL_PROFILE is None:
        launch()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    KERNEL_PROFILE.append((name, float(nbytes), e0, e1))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _req_cuda(*ts: torch.Tensor | None) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.B200Error("b200 ops require CUDA tensors (no CPU fallback)")


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         epi: int = EPI_BF16, bias: torch.Tensor | None = None, out2: torch.Tensor | None = None,
         aux: torch.Tensor | None = None, gamma: torch.Tensor | None = None,
         rowscale: torch.Tensor | None = None, rows_per_scale: int = 1, alpha: float = 1.0,
         splits: int = 1, block_n: int = 0, ws_mode: int = 0) -> torch.Tensor:
    """out[M,N] = epi(alpha * A[M,K] @ B[N,K]^T). a/b are bf16 2-D (last dim contiguous).

    a_mn / b_mn: the tensor passed is the TRANSPOSED operand, i.e. stored [K, M] / [K, N].
    """
    _req_cuda(a, b, out, bias, out2, aux, gamma, rowscale)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1 and out.stride(-1) == 1
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    assert out.shape[0] == M and out.shape[1] == N, (out.shape, M, N)
    args = GemmArgs()
    args.A, args.lda, args.a_mn = a.data_ptr(), a.stride(0), int(a_mn)
    args.B, args.ldb, args.b_mn = b.data_ptr(), b.stride(0), int(b_mn)
    args.M, args.N, args.K = M, N, K
    args.splits, args.epi, args.block_n, args.alpha, args.ws_mode = splits, epi, block_n, alpha, ws_mode
    args.C, args.ldc = out.data_ptr(), out.stride(0)
    if out2 is not None:
        args.C2, args.ldc2 = out2.data_ptr(), out2.stride(0)
    if aux is not None:
        args.aux, args.ldaux = aux.data_ptr(), aux.stride(0)
    args.bias = _ptr(bias)
    args.gamma = _ptr(gamma)
    args.rowscale = _ptr(rowscale)
    args.rows_per_scale = rows_per_scale
    if GEMM_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_lib.lib().b200_gemm(C.byref(args), _stream()), "b200_gemm")
        e1.record()
        GEMM_PROFILE.append((2.0 * M * N * K, e0, e1))
        GEMM_PROFILE_KEYS.append((M, N, K, int(a_mn), int(b_mn), epi, splits))
        return out
    check(_lib.lib().b200_gemm(C.byref(args), _stream()), "b200_gemm")
    return out


LN_GEMM_MAX_K = 384  # b200_ln_gemm: K % 64 == 0 and K <= 384 (the 128 x K bf16 panel stays resident in shared memory)


def ln_gemm(x: torch.Tensor, ln_w: torch.Tensor, ln_b: torch.Tensor, eps: float, b: torch.Tensor, out: torch.Tensor, *,
            epi: int = EPI_BF16, bias: torch.Tensor | None = None, out2: torch.Tensor | None = None,
            xn_out: torch.Tensor | None = None, mean: torch.Tensor | None = None, rstd: torch.Tensor | None = None,
            block_n: int = 0) -> torch.Tensor:
    """out[M,N] = epi(LayerNorm(x)[M,K] @ B[N,K]^T): the LayerNorm is the GEMM's A-operand prologue (x f32, B bf16).
    xn_out (bf16 [M,K]) / mean / rstd (f32 [M]): optional side outputs for the backward."""
    _req_cuda(x, ln_w, ln_b, b, out, bias, out2, xn_out, mean, rstd)
    assert x.dtype == torch.float32 and b.dtype == torch.bfloat16 and x.dim() == 2 and b.dim() == 2
    assert x.stride(1) == 1 and b.stride(1) == 1 and out.stride(-1) == 1
    M, K = x.shape
    N, Kb = b.shape
    assert K == Kb and out.shape[0] == M and out.shape[1] == N
    g = GemmArgs()
    g.A, g.lda, g.a_mn = None, 0, 0
    g.B, g.ldb, g.b_mn = b.data_ptr(), b.stride(0), 0
    g.M, g.N, g.K = M, N, K
    g.splits, g.epi, g.block_n, g.alpha, g.ws_mode = 1, epi, block_n, 1.0, 0
    g.C, g.ldc = out.data_ptr(), out.stride(0)
    if out2 is not None:
        g.C2, g.ldc2 = out2.data_ptr(), out2.stride(0)
    g.bias = _ptr(bias)
    ln = _lib.LnArgs()
    ln.x, ln.ldx = x.data_ptr(), x.stride(0)
    ln.weight, ln.bias, ln.eps = ln_w.data_ptr(), ln_b.data_ptr(), eps
    if xn_out is not None:
        ln.xn_out, ln.ld_xn = xn_out.data_ptr(), xn_out.stride(0)
    ln.mean, ln.rstd = _ptr(mean), _ptr(rstd)
    if GEMM_PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_lib.lib().b200_ln_gemm(C.byref(g), C.byref(ln), _stream()), "b200_ln_gemm")
        e1.record()
        GEMM_PROFILE.append((2.0 * M * N * K, e0, e1))
        GEMM_PROFILE_KEYS.append((M, N, K, 0, 0, epi, -1))
        return out
    check(_lib.lib().b200_ln_gemm(C.byref(g), C.byref(ln), _stream()), "b200_ln_gemm")
    return out


class AdamWArgs(C.Structure):
    _fields_ = [
        ("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("t", C.c_void_p),
        ("p_bf16", C.c_void_p), ("t_bf16", C.c_void_p),
        ("n", C.c_longlong), ("chunk", C.c_int),
        ("lr_scale", C.c_void_p), ("wd_scale", C.c_void_p), ("flags", C.c_void_p),
        ("lr", C.c_float), ("wd", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("step", C.c_int), ("ema_m", C.c_float),
        ("gradnorm_sq", C.c_void_p), ("dyn", C.c_void_p), ("max_norm", C.c_float), ("grad_scale", C.c_float),
        ("freeze_last_layer", C.c_int), ("freeze_backbone", C.c_int),
    ]


def _L():
    return _lib.lib()


# Sequence lengths the tcgen05 attention kernels take (global crops: N = 197 / 201 / ...); the rest stay on the
# warp-level kernels.  Module switches so that tests / A-B timing can pin either implementation.
import os as _os

# validated on hardware (tools/attn_check.py, tests/test_kernels_gpu.py); the defaults follow the measured A/B timings
TC_ATTENTION_FWD = _os.environ.get("B200_TC_ATTN_FWD", "1") == "1"
TC_ATTENTION_BWD = _os.environ.get("B200_TC_ATTN_BWD", "1") == "1"
TC_ATTENTION_PACKED = _os.environ.get("B200_TC_ATTN_PACKED", "1") == "1"  # N <= 128: several images of one head per CTA


def attention_fwd(qkv: torch.Tensor, B: int, N: int, h: int, out: torch.Tensor, lse: torch.Tensor | None,
                  scale: float) -> None:
    """qkv bf16 [B*N, 3*h*64]; out bf16 [B*N, h*64]; lse f32 [B*h, N]."""
    _req_cuda(qkv, out, lse)
    if (TC_ATTENTION_FWD and 128 < N <= 256) or (TC_ATTENTION_PACKED and N <= 128):
        return attention_fwd_tc(qkv, B, N, h, out, lse, scale)
    check(_L().b200_attention_fwd(qkv.data_ptr(), qkv.stride(0), B, N, h, 64, scale, out.data_ptr(), out.stride(0),
                                  _ptr(lse), _stream()), "b200_attention_fwd")


def attention_fwd_tc(qkv, B: int, N: int, h: int, out, lse, scale: float) -> None:
    """tcgen05 / TMEM / TMA forward (csrc/attention_tc.cu); same contract as attention_fwd, N <= 256."""
    _req_cuda(qkv, out)
    check(_L().b200_attention_fwd_tc(qkv.data_ptr(), qkv.stride(0), B, N, h, 64, scale, out.data_ptr(), out.stride(0),
                                     _ptr(lse), _stream()), "b200_attention_fwd_tc")


def attention_bwd(qkv, out, dout, lse, B: int, N: int, h: int, dqkv, scale: float, colsum=None) -> None:
    """colsum (optional f32 [3*h*64]): += column sums of dqkv, i.e. the gradient of the qkv projection's bias."""
    _req_cuda(qkv, out, dout, lse, dqkv)
    assert out.stride(0) == dout.stride(0)
    if (TC_ATTENTION_BWD and 128 < N <= 208) or (TC_ATTENTION_PACKED and N <= 128):
        return attention_bwd_tc(qkv, out, dout, lse, B, N, h, dqkv, scale, colsum)
    check(_L().b200_attention_bwd(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), dout.data_ptr(), dout.stride(0),
                                  lse.data_ptr(), B, N, h, 64, scale, dqkv.data_ptr(), dqkv.stride(0), _ptr(colsum),
                                  _stream()), "b200_attention_bwd")


def attention_bwd_tc(qkv, out, dout, lse, B: int, N: int, h: int, dqkv, scale: float, colsum=None) -> None:
    """tcgen05 / TMEM / TMA backward (csrc/attention_tc.cu); same contract as attention_bwd, N <= 208."""
    _req_cuda(qkv, out, dout, lse, dqkv)
    assert out.stride(0) == dout.stride(0)
    check(_L().b200_attention_bwd_tc(qkv.data_ptr(), qkv.stride(0), out.data_ptr(), dout.data_ptr(), dout.stride(0),
                                     lse.data_ptr(), B, N, h, 64, scale, dqkv.data_ptr(), dqkv.stride(0), _ptr(colsum),
                                     _stream()), "b200_attention_bwd_tc")


def layernorm_fwd(x, w, b, eps: float, y, mean=None, rstd=None) -> None:
    """x f32 [T,D] -> y (bf16 or f32) [T,D]."""
    _req_cuda(x, w, b, y)
    T, D = x.shape
    check(_L().b200_layernorm_fwd(x.data_ptr(), x.stride(0), T, D, w.data_ptr(), b.data_ptr(), eps, y.data_ptr(),
                                  y.stride(0), int(y.dtype == torch.bfloat16), _ptr(mean), _ptr(rstd), _stream()),
          "b200_layernorm_fwd")


def layernorm_bwd(dy, x, w, mean, rstd, dx, accumulate: bool, dw=None, db=None) -> None:
    _req_cuda(dy, x, w, mean, rstd, dx)
    T, D = x.shape
    check(_L().b200_layernorm_bwd(dy.data_ptr(), dy.stride(0), int(dy.dtype == torch.bfloat16), x.data_ptr(),
                                  x.stride(0), T, D, w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                                  dx.stride(0), int(accumulate), _ptr(dw), _ptr(db), _stream()), "b200_layernorm_bwd")


def layernorm_bwd_ls(dy, x, w, mean, rstd, dx, accumulate: bool, dw, db, o, gamma, rowscale, rows_per_scale: int, dout,
                     dgamma, dbias) -> None:
    """ln_bwd fused with the layerscale_bwd that follows it (same dx)."""
    T, D = x.shape
    check(_L().b200_layernorm_bwd_ls(dy.data_ptr(), dy.stride(0), int(dy.dtype == torch.bfloat16), x.data_ptr(), x.stride(0),
                                     T, D, w.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), dx.stride(0),
                                     int(accumulate), _ptr(dw), _ptr(db), o.data_ptr(), o.stride(0), _ptr(gamma),
                                     _ptr(rowscale), rows_per_scale, dout.data_ptr(), dout.stride(0), _ptr(dgamma),
                                     _ptr(dbias), _stream()), "b200_layernorm_bwd_ls")


def im2col(x, p: int, cols) -> None:
    _req_cuda(x, cols)
    B, Cc, H, W = x.shape
    assert x.is_contiguous() and x.dtype == torch.float32
    check(_L().b200_im2col(x.data_ptr(), B, Cc, H, W, p, cols.data_ptr(), cols.stride(0), _stream()), "b200_im2col")


def assemble_tokens(tok, masks_u8, mask_token, cls, reg, pos, B: int, Np: int, R: int, D: int, x) -> None:
    check(_L().b200_assemble_tokens(tok.data_ptr(), tok.stride(0), _ptr(masks_u8), _ptr(mask_token), cls.data_ptr(),
                                    _ptr(reg), pos.data_ptr(), B, Np, R, D, x.data_ptr(), _stream()),
          "b200_assemble_tokens")


def assemble_tokens_bwd(dx, masks_u8, B: int, Np: int, R: int, D: int, dtok, dpos, dcls, dreg, dmask_token) -> None:
    check(_L().b200_assemble_tokens_bwd(dx.data_ptr(), _ptr(masks_u8), B, Np, R, D, dtok.data_ptr(), dtok.stride(0),
                                        dpos.data_ptr(), dcls.data_ptr(), _ptr(dreg), _ptr(dmask_token), _stream()),
          "b200_assemble_tokens_bwd")


def layerscale_bwd(dx, o, gamma, rowscale, rows_per_scale: int, dout, dgamma, dbias) -> None:
    T, D = dx.shape
    check(_L().b200_layerscale_bwd(dx.data_ptr(), dx.stride(0), o.data_ptr(), o.stride(0), _ptr(gamma), _ptr(rowscale),
                                   rows_per_scale, T, D, dout.data_ptr(), dout.stride(0), _ptr(dgamma), _ptr(dbias),
                                   _stream()), "b200_layerscale_bwd")


def swiglu_fwd(x12, hidden) -> None:
    """hidden = silu(x12[:, :H]) * x12[:, H:]  (bf16, autocast rounding points)."""
    _req_cuda(x12, hidden)
    T, H = hidden.shape
    assert x12.shape == (T, 2 * H) and x12.dtype == hidden.dtype == torch.bfloat16
    check(_L().b200_swiglu_fwd(x12.data_ptr(), x12.stride(0), T, H, hidden.data_ptr(), hidden.stride(0), _stream()),
          "b200_swiglu_fwd")


def swiglu_bwd(x12, dhidden, dx12) -> None:
    _req_cuda(x12, dhidden, dx12)
    T, H = dhidden.shape
    assert x12.shape == dx12.shape == (T, 2 * H)
    check(_L().b200_swiglu_bwd(x12.data_ptr(), x12.stride(0), dhidden.data_ptr(), dhidden.stride(0), T, H, dx12.data_ptr(),
                               dx12.stride(0), _stream()), "b200_swiglu_bwd")


def gather_rows(src, idx, out, Np: int = 0, N: int = 0, off: int = 0) -> None:
    """out[m] = src[map(idx[m])]; src f32 2-D; idx int64."""
    M, D = out.shape
    check(_L().b200_gather_rows(src.data_ptr(), src.stride(0), idx.data_ptr(), M, D, Np, N, off, out.data_ptr(),
                                out.stride(0), int(out.dtype == torch.bfloat16), _stream()), "b200_gather_rows")


def scatter_rows(inp, idx, dst, Np: int = 0, N: int = 0, off: int = 0, accumulate: bool = False, count_dev=None) -> None:
    M, D = inp.shape
    check(_L().b200_scatter_rows(inp.data_ptr(), inp.stride(0), int(inp.dtype == torch.bfloat16), _ptr(idx), M, D, Np, N,
                                 off, dst.data_ptr(), dst.stride(0), int(accumulate), _ptr(count_dev), _stream()),
          "b200_scatter_rows")


def l2norm_fwd(x, y, nrm, eps: float = 1e-12) -> None:
    R, D = x.shape
    check(_L().b200_l2norm_fwd(x.data_ptr(), x.stride(0), R, D, eps, y.data_ptr(), y.stride(0), nrm.data_ptr(), _stream()),
          "b200_l2norm_fwd")


def l2norm_bwd(dy, x, nrm, dx) -> None:
    R, D = x.shape
    check(_L().b200_l2norm_bwd(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), nrm.data_ptr(), R, D, dx.data_ptr(),
                               dx.stride(0), _stream()), "b200_l2norm_bwd")


def weightnorm_fwd(g, v, w_bf16, vnorm=None) -> None:
    O, I = v.shape
    check(_L().b200_weightnorm_fwd(g.data_ptr(), v.data_ptr(), O, I, w_bf16.data_ptr(), _ptr(vnorm), _stream()),
          "b200_weightnorm_fwd")


def weightnorm_bwd(dW, g, v, dg, dv) -> None:
    O, I = v.shape
    check(_L().b200_weightnorm_bwd(dW.data_ptr(), g.data_ptr(), v.data_ptr(), O, I, dg.data_ptr(), dv.data_ptr(), _stream()),
          "b200_weightnorm_bwd")


def small_matmul(A, Bm, Cm, a_trans: bool = False, accumulate: bool = False) -> None:
    """C[M,N] (+)= op(A) @ B ; B row-major [K,N]."""
    K, N = Bm.shape
    M = A.shape[1] if a_trans else A.shape[0]
    check(_L().b200_small_matmul(A.data_ptr(), A.stride(0), int(a_trans), Bm.data_ptr(), Bm.stride(0), M, N, K,
                                 Cm.data_ptr(), Cm.stride(0), int(accumulate), _stream()), "b200_small_matmul")


def cast_bf16(x, y) -> None:
    check(_L().b200_cast_bf16(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "b200_cast_bf16")


def fill_f32(x, v: float = 0.0) -> None:
    check(_L().b200_fill_f32(x.data_ptr(), x.numel(), v, _stream()), "b200_fill_f32")


def row_lse(x, colterm, scale: float, rowterm, scale_dev=None) -> None:
    R, K = x.shape
    _timed("row_lse", 2.0 * R * K, lambda: check(_L().b200_row_lse(x.data_ptr(), x.stride(0), R, K, _ptr(colterm), scale,
                                                                   _ptr(scale_dev), rowterm.data_ptr(), _stream()), "b200_row_lse"))


def col_reduce(x, out, rowvec=None, scale: float = 1.0, mode: int = 0, scale_dev=None) -> None:
    R, K = x.shape
    check(_L().b200_col_reduce(x.data_ptr(), x.stride(0), R, K, _ptr(rowvec), scale, _ptr(scale_dev), mode, out.data_ptr(),
                               _stream()), "b200_col_reduce")


def vec_op(y, x, a: float, b: float, op: int, a_dev=None) -> None:
    check(_L().b200_vec_op(y.data_ptr(), x.data_ptr(), y.numel(), a, b, op, _ptr(a_dev), _stream()), "b200_vec_op")


def dino_ce(s, t, colterm, t_rowterm, t_idx0, t_idx1, weight, s_scale: float, t_scale: float, loss_rows, ds=None,
            gscale: float = 1.0, t_scale_dev=None) -> None:
    """t_rowterm=None (single-teacher rows only): the kernel computes the teacher rows' log-sum-exp itself in its first
    pass (the teacher logits are then read from HBM once per step instead of twice)."""
    Rs, K = s.shape
    nbytes = ((4.0 if ds is not None else 2.0) + (2.0 if t_rowterm is None else 0.0)) * Rs * K
    _timed("dino_ce", nbytes, lambda: check(
        _L().b200_dino_ce(s.data_ptr(), s.stride(0), Rs, K, t.data_ptr(), t.stride(0), _ptr(colterm),
                          _ptr(t_rowterm), t_idx0.data_ptr(), _ptr(t_idx1), _ptr(weight), s_scale, t_scale, _ptr(t_scale_dev), gscale,
                          loss_rows.data_ptr(), _ptr(ds), ds.stride(0) if ds is not None else 0, _stream()), "b200_dino_ce"))


def segment_sum(x, offsets_i32, out, scale=None) -> None:
    check(_L().b200_segment_sum(x.data_ptr(), offsets_i32.data_ptr(), out.numel(), _ptr(scale), out.data_ptr(), _stream()),
          "b200_segment_sum")


def koleo(x, groups: int, n: int, loss_out, dx=None, gscale: float = 1.0, eps: float = 1e-8, bf16_sim: bool = True,
          nn_out=None) -> None:
    D = x.shape[1]
    scratch = None
    if (2 * n * D + 3 * n) * 4 > 220 * 1024 or n > 256:  # row-tiled path: normalised rows / gradient accumulator in a scratch
        scratch = torch.empty(groups * n * (2 * D + 1), device=x.device, dtype=torch.float32)
    check(_L().b200_koleo(x.data_ptr(), x.stride(0), groups, n, D, eps, int(bf16_sim), gscale, loss_out.data_ptr(),
                          _ptr(dx), dx.stride(0) if dx is not None else 0, _ptr(nn_out), _ptr(scratch),
                          scratch.numel() if scratch is not None else 0, _stream()), "b200_koleo")


def ema(teacher_flat, student_flat, m: float, teacher_bf16=None) -> None:
    _req_cuda(teacher_flat, student_flat)
    check(_L().b200_ema(teacher_flat.data_ptr(), student_flat.data_ptr(), teacher_flat.numel(), m, _ptr(teacher_bf16),
                        _stream()), "b200_ema")


def sumsq(x, out) -> None:
    _timed("sumsq", 4.0 * x.numel(), lambda: check(_L().b200_sumsq(x.data_ptr(), x.numel(), out.data_ptr(), _stream()), "b200_sumsq"))


def adamw_ema(args: AdamWArgs) -> None:
    nb = float(args.n) * ((20.0 + 20.0) if args.t else (16.0 + 14.0))
    _timed("adamw_ema", nb, lambda: check(_L().b200_adamw_ema(C.byref(args), _stream()), "b200_adamw_ema"))


def rope_apply(qkv, B: int, N: int, prefix: int, h: int, sin_tab, cos_tab) -> None:
    """In-place axial RoPE on q, k of the patch tokens; qkv bf16 [B*N, 3*h*64]; sin/cos f32 [N - prefix, 64]."""
    _req_cuda(qkv, sin_tab, cos_tab)
    assert sin_tab.dtype == cos_tab.dtype == torch.float32 and sin_tab.is_contiguous() and cos_tab.is_contiguous()
    assert sin_tab.shape == (N - prefix, 64) == cos_tab.shape
    check(_L().b200_rope_apply(qkv.data_ptr(), qkv.stride(0), B, N, prefix, h, 64, sin_tab.data_ptr(), cos_tab.data_ptr(),
                               _stream()), "b200_rope_apply")


def kl_rows(s, t, inv_temp: float, loss_rows, ds=None, gscale: float = 1.0) -> None:
    """loss_rows[r] = KL(softmax(t[r] * inv_temp) || softmax(s[r] * inv_temp)); ds = gscale * d loss_rows / d s.  f32 2-D."""
    _req_cuda(s, t, loss_rows, ds)
    R, K = s.shape
    assert s.dtype == t.dtype == torch.float32 and s.stride(1) == 1 and t.stride(1) == 1 and t.shape == s.shape
    check(_L().b200_kl_rows(s.data_ptr(), s.stride(0), t.data_ptr(), t.stride(0), R, K, inv_temp, gscale, loss_rows.data_ptr(),
                            _ptr(ds), ds.stride(0) if ds is not None else 0, _stream()), "b200_kl_rows")


def gather_samples(src, idx, dst) -> None:
    """dst[j] = src[idx[j]] for fp32 [*, N, D] sample blocks (batch-subset stochastic depth)."""
    _req_cuda(src, idx, dst)
    assert src.dtype == dst.dtype == torch.float32 and idx.dtype == torch.int64 and src.is_contiguous() and dst.is_contiguous()
    row = src[0].numel()
    check(_L().b200_copy_samples(src.data_ptr(), dst.data_ptr(), idx.data_ptr(), idx.numel(), row, 0, _stream()), "b200_copy_samples")


def scatter_samples(src, idx, dst) -> None:
    """dst[idx[j]] = src[j] (overwrite) for fp32 [*, N, D] sample blocks."""
    _req_cuda(src, idx, dst)
    assert src.dtype == dst.dtype == torch.float32 and idx.dtype == torch.int64 and src.is_contiguous() and dst.is_contiguous()
    row = dst[0].numel()
    check(_L().b200_copy_samples(src.data_ptr(), dst.data_ptr(), idx.data_ptr(), idx.numel(), row, 1, _stream()), "b200_copy_samples")


def mse(teacher, student, out, ds=None) -> None:
    """out[0] += mean((teacher - student)^2); ds = gradient wrt student.  Flat contiguous f32 tensors."""
    _req_cuda(teacher, student, out, ds)
    check(_L().b200_mse(teacher.data_ptr(), student.data_ptr(), student.numel(), out.data_ptr(), _ptr(ds), _stream()), "b200_mse")


def block_masks(targets_i32, H: int, W: int, max_patches: int, seed: int, masks_u8, step_dev=None, min_patches: int = 4,
                min_aspect: float = 0.3, max_aspect: float = 1 / 0.3) -> None:
    """Device block-wise masks: targets int32 [B] -> masks_u8 [B, H*W]."""
    _req_cuda(targets_i32, masks_u8, step_dev)
    B = targets_i32.numel()
    check(_L().b200_block_masks(targets_i32.data_ptr(), B, H, W, min_patches, max_patches, min_aspect, max_aspect, seed,
                                _ptr(step_dev), masks_u8.data_ptr(), _stream()), "b200_block_masks")


def random_subset(n: int, k: int, seed: int, counter, idx_i64) -> None:
    """idx_i64[:k] = uniformly random k-subset of range(n) in random order; counter (int64 [1], device) += 1."""
    _req_cuda(counter, idx_i64)
    check(_L().b200_random_subset(n, k, seed, counter.data_ptr(), idx_i64.data_ptr(), _stream()), "b200_random_subset")


def collate_masks(masks_u8, cap: int, idx_i64, weight, row_w, pad, m_valid) -> None:
    _req_cuda(masks_u8, idx_i64, weight, row_w, pad, m_valid)
    B, Np = masks_u8.shape
    check(_L().b200_collate_masks(masks_u8.data_ptr(), B, Np, cap, idx_i64.data_ptr(), weight.data_ptr(), row_w.data_ptr(),
                                  pad.data_ptr(), m_valid.data_ptr(), _stream()), "b200_collate_masks")
