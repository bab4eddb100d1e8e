"""Timing of LayerNorm + GEMM as two kernels vs the LayerNorm-prologue GEMM (b200_ln_gemm) at the cfg2 shapes."""
import sys
import torch
sys.path.insert(0, ".")
from lightly_train_b200 import ops

dev = "cuda"
torch.manual_seed(0)


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


for (M, K, N, epi, name) in [(25216, 384, 1152, ops.EPI_BF16, "qkv global"), (25216, 384, 1536, ops.EPI_BIAS_GELU, "fc1 global (teacher)"),
                             (25216, 384, 1536, ops.EPI_BIAS_GELU_DG, "fc1 global (student)"), (18944, 384, 1152, ops.EPI_BF16, "qkv local"),
                             (18944, 384, 1536, ops.EPI_BIAS_GELU_DG, "fc1 local (student)")]:
    x = torch.randn(M, K, device=dev)
    lw, lb = torch.ones(K, device=dev), torch.zeros(K, device=dev)
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(N, device=dev)
    xn = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if epi == ops.EPI_BIAS_GELU_DG else None
    save = epi != ops.EPI_BIAS_GELU

    def two():
        ops.layernorm_fwd(x, lw, lb, 1e-6, xn, mean if save else None, rstd if save else None)
        ops.gemm(xn, w, out, epi=epi, bias=bias, out2=out2)

    t2 = timeit(two)
    res = [f"{name:22s} M={M} N={N}: LN + GEMM {t2:6.1f} us"]
    for bn in (128, 192):
        def fused():
            ops.ln_gemm(x, lw, lb, 1e-6, w, out, epi=epi, bias=bias, out2=out2, xn_out=xn if save else None,
                        mean=mean if save else None, rstd=rstd if save else None, block_n=bn)
        res.append(f"fused bn={bn} {timeit(fused):6.1f} us")
    print(" | ".join(res), flush=True)
