"""Plain qkv-shaped GEMM through the single-CTA and the CTA-pair kernels, one launch each, for `ncu --set full`."""
import sys
import torch
sys.path.insert(0, ".")
from lightly_train_b200 import ops

dev = "cuda"
T, D = 25216, 384
x = torch.randn(T, D, device=dev).bfloat16()
w = (torch.randn(3 * D, D, device=dev) * 0.02).bfloat16()
out = torch.empty(T, 3 * D, device=dev, dtype=torch.bfloat16)


def cases():
    ops.gemm(x, w, out, block_n=192, ws_mode=2)
    ops.gemm(x, w, out, block_n=192, ws_mode=3)
    ops.gemm(x, w, out, block_n=256, ws_mode=3)


for _ in range(3):
    cases()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
cases()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
