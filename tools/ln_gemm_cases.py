"""Two launches of b200_ln_gemm per shape (local-crop and global-crop qkv) for `ncu -k regex:gemm_ln`."""
import sys
import torch
sys.path.insert(0, ".")
from lightly_train_b200 import ops

dev = "cuda"
torch.manual_seed(0)
for M in (18944, 25216):
    K, N = 384, 1152
    x = torch.randn(M, K, device=dev)
    lw, lb = torch.ones(K, device=dev), torch.zeros(K, device=dev)
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(N, device=dev)
    xn = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for _ in range(2):
        flush.zero_()
        ops.ln_gemm(x, lw, lb, 1e-6, w, out, bias=bias, xn_out=xn, mean=mean, rstd=rstd, block_n=128)
    torch.cuda.synchronize()
print("done")
