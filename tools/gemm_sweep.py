"""Tile-width / schedule sweep of the cfg2 GEMM shapes (isolated launches, CUDA events, 20 iterations each)."""
import sys
import torch
sys.path.insert(0, ".")
from lightly_train_b200 import ops

dev = "cuda"


def bench(M, N, K, a_mn=False, b_mn=False, epi=ops.EPI_BF16, splits=1, block_n=0, ws_mode=0, iters=20):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    f32 = epi in (ops.EPI_F32, ops.EPI_F32_ATOMIC, ops.EPI_RESIDUAL)
    out = torch.zeros((M, N), device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    extra = {}
    if epi == ops.EPI_RESIDUAL:
        extra = dict(aux=torch.randn(M, N, device=dev), gamma=torch.randn(N, device=dev),
                     out2=torch.empty(M, N, device=dev, dtype=torch.bfloat16), bias=torch.randn(N, device=dev))
    kw = dict(a_mn=a_mn, b_mn=b_mn, epi=epi, splits=splits, block_n=block_n, ws_mode=ws_mode, **extra)
    try:
        for _ in range(3):
            ops.gemm(a, b, out, **kw)
    except Exception as e:  # unsupported combination
        print(f"M={M} N={N} K={K} epi={epi} splits={splits} bn={block_n} ws={ws_mode}: {type(e).__name__}")
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, b, out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} epi={epi} splits={splits} bn={block_n} ws={ws_mode}: "
          f"{ms * 1e3:6.1f} us  {2.0 * M * N * K / ms / 1e9:5.0f} TF", flush=True)


T = 25216
for bn in (128, 192, 256):
    for ws in (2, 1):
        bench(T, 1152, 384, block_n=bn, ws_mode=ws)
        bench(T, 1536, 384, block_n=bn, ws_mode=ws)
for bn in (0, 128, 192, 256):
    bench(384, 1536, T, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=0, block_n=bn)
    bench(1536, 384, T, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=0, block_n=bn)
    bench(1152, 384, T, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=0, block_n=bn)
    bench(384, 384, T, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=0, block_n=bn)
for bn in (128, 192, 256):
    bench(T, 384, 1536, b_mn=True, block_n=bn)   # dgrad fc1
    bench(T, 384, 1152, b_mn=True, block_n=bn)   # dgrad qkv
    bench(T, 384, 384, epi=ops.EPI_RESIDUAL, block_n=bn)
    bench(T, 384, 1536, epi=ops.EPI_RESIDUAL, block_n=bn)
