"""Thin tensor-level wrappers over the C-ABI (include/b200dino.h).

torch is used for device memory and streams only; every op here enqueues hand-written sm_100a kernels from
libb200dino.so on the current CUDA stream.  No op has a CPU / PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_BF16, EPI_BIAS_GELU, EPI_DGELU, EPI_F32, EPI_F32_ATOMIC,  # noqa: F401
                   EPI_RESIDUAL, GemmArgs, check)


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
         splits: int = 1, block_n: int = 0) -> torch.Tensor:
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
    args.splits, args.epi, args.block_n, args.alpha = splits, epi, block_n, alpha
    args.C, args.ldc = out.data_ptr(), out.stride(0)
    if out2 is not None:
        args.C2, args.ldc2 = out2.data_ptr(), out2.stride(0)
    if aux is not None:
        args.aux, args.ldaux = aux.data_ptr(), aux.stride(0)
    args.bias = _ptr(bias)
    args.gamma = _ptr(gamma)
    args.rowscale = _ptr(rowscale)
    args.rows_per_scale = rows_per_scale
    check(_lib.lib().b200_gemm(C.byref(args), _stream()), "b200_gemm")
    return out
