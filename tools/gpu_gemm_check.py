"""GPU bring-up check for b200_gemm (run under gpurun): correctness vs torch fp32 matmul + quick timing."""
import sys, time
import torch
sys.path.insert(0, ".")
from lightly_train_b200 import ops
from lightly_train_b200._lib import *  # noqa

torch.manual_seed(0)
dev = "cuda"
fails = 0


def ref_mm(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float().t() if b_mn else b.float()
    return A @ B.t()


def run(M, N, K, a_mn=False, b_mn=False, epi=EPI_BF16, block_n=0, splits=1, ws_mode=2):
    global fails
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    ref = ref_mm(a, b, a_mn, b_mn) / 8.0
    bias = torch.randn(N, device=dev)
    kw = dict(a_mn=a_mn, b_mn=b_mn, epi=epi, block_n=block_n, splits=splits, alpha=1 / 8.0, ws_mode=ws_mode)
    if epi == EPI_BF16:
        out = torch.full((M, N), 7.0, device=dev, dtype=torch.bfloat16)
        ops.gemm(a, b, out, bias=bias, **kw)
        want = (ref + bias).bfloat16().float(); got = out.float()
    elif epi == EPI_F32:
        out = torch.full((M, N), 7.0, device=dev)
        ops.gemm(a, b, out, **kw)
        want = ref; got = out
    elif epi == EPI_F32_ATOMIC:
        out = torch.ones((M, N), device=dev)
        ops.gemm(a, b, out, **kw)
        want = ref + 1; got = out
    elif epi == EPI_BIAS_GELU:
        out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
        out2 = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
        ops.gemm(a, b, out, bias=bias, out2=out2, **kw)
        u = (ref + bias).bfloat16()
        want = torch.nn.functional.gelu(u.float()).bfloat16().float(); got = out.float()
        e2 = (out2.float() - u.float()).abs().max().item()
        if e2 > 0.07: print("   pre-act mismatch", e2); fails += 1
    elif epi == EPI_RESIDUAL:
        x = torch.randn((M, N), device=dev)
        gamma = torch.randn(N, device=dev)
        rps = 16
        rs = torch.rand((M + rps - 1) // rps, device=dev)
        out = torch.empty((M, N), device=dev)
        out2 = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
        ops.gemm(a, b, out, bias=bias, out2=out2, aux=x, gamma=gamma, rowscale=rs, rows_per_scale=rps, **kw)
        o = (ref + bias).bfloat16().float()
        want = x + o * gamma * rs.repeat_interleave(rps)[:M, None]; got = out
    elif epi == EPI_DGELU:
        u = torch.randn((M, N), device=dev).bfloat16()
        out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
        ops.gemm(a, b, out, aux=u, **kw)
        uf = u.float().requires_grad_(True)
        torch.nn.functional.gelu(uf).backward(ref.bfloat16().float())
        want = uf.grad.bfloat16().float(); got = out.float()
    torch.cuda.synchronize()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    ok = err <= 0.02 * max(scale, 1.0)
    if not ok: fails += 1
    print(f"{'OK ' if ok else 'BAD'} M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} epi={epi} bn={block_n} splits={splits} ws={ws_mode} err={err:.4g} scale={scale:.3g}", flush=True)


import os
MODE = os.environ.get("CHECK", "all")
if MODE == "all":
    print(lib().b200_version())
    # basic K-major, all tile widths
    for bn in (128, 192, 256):
        run(256, 256, 64, block_n=bn)
        run(256, 512, 384, block_n=bn)
        run(1000, 1152, 384, block_n=bn)   # M tail
    run(128, 384, 64, epi=EPI_F32)
    run(130, 392, 72, epi=EPI_F32)         # M, N, K tails
    run(4403, 2048, 384)
    # majors
    for a_mn in (False, True):
        for b_mn in (False, True):
            run(384, 1152, 1000, a_mn=a_mn, b_mn=b_mn, epi=EPI_F32, block_n=192)
            run(256, 256, 128, a_mn=a_mn, b_mn=b_mn, epi=EPI_F32, block_n=128)
            run(256, 256, 128, a_mn=a_mn, b_mn=b_mn, epi=EPI_F32, block_n=256)
    # split-K atomic (wgrad-like)
    run(1152, 384, 25216, a_mn=True, b_mn=True, epi=EPI_F32_ATOMIC, splits=16, block_n=192)
    run(384, 1536, 5000, a_mn=True, b_mn=True, epi=EPI_F32_ATOMIC, splits=7)
    # weight-stationary schedule (auto for K<=384 and >= 8 M tiles; forced here on small/odd shapes too)
    for epi in (EPI_BF16, EPI_F32, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_DGELU):
        run(5000, 1152, 384, epi=epi, ws_mode=1)
        run(3000, 1536, 384, epi=epi, ws_mode=1, block_n=128)
    run(4403, 4096, 256, ws_mode=1)                      # K=256 -> 256-wide slab
    run(1300, 392, 200, epi=EPI_F32, ws_mode=1)          # tails in M, N, K
    run(2500, 1536, 384, b_mn=True, epi=EPI_DGELU, ws_mode=1)
    run(2504, 384, 384, a_mn=True, b_mn=True, epi=EPI_F32, ws_mode=1)
    run(25216, 1152, 384, ws_mode=0)
    # fused epilogues
    run(1000, 1536, 384, epi=EPI_BIAS_GELU)
    run(1000, 384, 1536, epi=EPI_RESIDUAL)
    run(1000, 1536, 384, epi=EPI_DGELU)
elif MODE == "2sm":
    # CTA-pair kernel (ws_mode=3): 256-row tiles, B split across the pair
    for bn in (128, 192, 256):
        run(256, 256, 64, block_n=bn, ws_mode=3)
        run(512, 512, 384, block_n=bn, ws_mode=3)
        run(1000, 1152, 384, block_n=bn, ws_mode=3)   # M tail inside the second CTA's half
        run(1100, 1152, 384, block_n=bn, ws_mode=3)   # M tail inside the first CTA's half (second half fully out of range)
        run(130, 392, 72, epi=EPI_F32, block_n=bn, ws_mode=3)
    for a_mn in (False, True):
        for b_mn in (False, True):
            run(384, 1152, 1000, a_mn=a_mn, b_mn=b_mn, epi=EPI_F32, block_n=256, ws_mode=3)
            run(512, 256, 128, a_mn=a_mn, b_mn=b_mn, epi=EPI_F32, block_n=128, ws_mode=3)
    run(1152, 384, 25216, a_mn=True, b_mn=True, epi=EPI_F32_ATOMIC, splits=16, block_n=128, ws_mode=3)
    run(384, 1536, 5000, a_mn=True, b_mn=True, epi=EPI_F32_ATOMIC, splits=7, block_n=256, ws_mode=3)
    for epi in (EPI_BF16, EPI_BIAS_GELU, EPI_RESIDUAL, EPI_DGELU):
        run(5000, 1536, 384, epi=epi, block_n=192, ws_mode=3)
        run(3000, 384, 1536, epi=epi, block_n=128, ws_mode=3)
    run(25216, 1152, 384, block_n=192, ws_mode=3)


def bench(M, N, K, a_mn=False, b_mn=False, epi=EPI_BF16, splits=1, block_n=0, iters=20, ws_mode=0):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    out = torch.zeros((M, N), device=dev, dtype=torch.float32 if epi in (EPI_F32, EPI_F32_ATOMIC, EPI_RESIDUAL) else torch.bfloat16)
    extra = {}
    if epi == EPI_RESIDUAL:
        extra = dict(aux=torch.randn(M, N, device=dev), gamma=torch.randn(N, device=dev), out2=torch.empty(M, N, device=dev, dtype=torch.bfloat16), bias=torch.randn(N, device=dev))
    if epi == EPI_BIAS_GELU:
        extra = dict(out2=torch.empty(M, N, device=dev, dtype=torch.bfloat16), bias=torch.randn(N, device=dev))
    if epi == EPI_DGELU:
        extra = dict(aux=torch.randn(M, N, device=dev).bfloat16())
    for _ in range(3): ops.gemm(a, b, out, a_mn=a_mn, b_mn=b_mn, epi=epi, splits=splits, block_n=block_n, ws_mode=ws_mode, **extra)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm(a, b, out, a_mn=a_mn, b_mn=b_mn, epi=epi, splits=splits, block_n=block_n, ws_mode=ws_mode, **extra)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    A = a.t() if a_mn else a
    B = b.t() if b_mn else b
    for _ in range(3): torch.matmul(A, B.t())
    e0.record()
    for _ in range(iters): torch.matmul(A, B.t())
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    print(f"bench M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} epi={epi} splits={splits} bn={block_n} ws={ws_mode}: {ms*1e3:.1f} us  {tf:.0f} TFLOP/s   (torch.matmul {ms_t*1e3:.1f} us, {2.0*M*N*K/ms_t/1e9:.0f} TF)", flush=True)


if fails == 0 and MODE == "2sm":
    T = 25216
    for bn in (128, 192, 256):
        bench(T, 1152, 384, block_n=bn, ws_mode=3)
        bench(T, 1536, 384, block_n=bn, ws_mode=3)
        bench(T, 384, 1536, b_mn=True, block_n=bn, ws_mode=3) if bn != 192 else None
        bench(T, 384, 384, epi=EPI_RESIDUAL, block_n=bn, ws_mode=3)
        bench(T, 384, 1536, epi=EPI_RESIDUAL, block_n=bn, ws_mode=3)
        bench(T, 1536, 384, epi=EPI_BIAS_GELU, block_n=bn, ws_mode=3)
    for bn, sp in ((256, 12), (256, 6), (128, 6), (128, 8)):
        bench(384, 1536, T, a_mn=True, b_mn=True, epi=EPI_F32_ATOMIC, splits=sp, block_n=bn, ws_mode=3)
        bench(1536, 384, T, a_mn=True, b_mn=True, epi=EPI_F32_ATOMIC, splits=sp, block_n=bn, ws_mode=3)
    bench(18944, 1152, 384, block_n=192, ws_mode=3)
    bench(18944, 1152, 384, block_n=192, ws_mode=2)
    bench(8192, 8192, 8192, block_n=256, ws_mode=3)
if fails == 0 and MODE == "all":
    for ws in (2, 0):
        bench(25216, 1152, 384, ws_mode=ws)
        bench(25216, 384, 384, ws_mode=ws)
        bench(25216, 1536, 384, ws_mode=ws)
        bench(25216, 384, 384, epi=EPI_RESIDUAL, ws_mode=ws)
        bench(25216, 1536, 384, epi=EPI_BIAS_GELU, ws_mode=ws)
        bench(25216, 1536, 384, b_mn=True, epi=EPI_DGELU, ws_mode=ws)
        bench(4403, 65536, 256, ws_mode=ws)
    bench(25216, 384, 1536, block_n=192)
    bench(25216, 384, 1536, epi=EPI_RESIDUAL)
    bench(8192, 8192, 8192)
    bench(1152, 384, 25216, a_mn=True, b_mn=True, epi=EPI_F32_ATOMIC, splits=8, block_n=192)
    bench(1152, 384, 25216, a_mn=True, b_mn=True, epi=EPI_F32_ATOMIC, splits=24, block_n=192)
print("FAILS", fails)
sys.exit(1 if fails else 0)
