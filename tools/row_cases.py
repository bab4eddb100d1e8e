"""Row-kernel launches of the cfg2 step (LayerNorm fwd / fused LN+LayerScale bwd / column reductions) for `ncu --set full`."""
import sys
import torch
sys.path.insert(0, ".")
from lightly_train_b200 import ops

dev = "cuda"
T, D, H = 25216, 384, 1536
bf = torch.bfloat16
x = torch.randn(T, D, device=dev)
w, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
xn = torch.empty(T, D, device=dev, dtype=bf)
mean, rstd = torch.empty(T, device=dev), torch.empty(T, device=dev)
dy = torch.randn(T, D, device=dev).to(bf)
dx = torch.randn(T, D, device=dev)
dw, db, dg, dbias = (torch.zeros(D, device=dev) for _ in range(4))
o = torch.randn(T, D, device=dev).to(bf)
gamma = torch.randn(D, device=dev)
rowscale = torch.ones(128, device=dev)
dout = torch.empty(T, D, device=dev, dtype=bf)
du = torch.randn(T, H, device=dev).to(bf)
dbu = torch.zeros(H, device=dev)


def cases():
    ops.layernorm_fwd(x, w, b, 1e-6, xn, mean, rstd)
    ops.layernorm_bwd_ls(dy, x, w, mean, rstd, dx, True, dw, db, o, gamma, rowscale, 197, dout, dg, dbias)
    ops.col_reduce(du, dbu)


for _ in range(3):
    cases()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
cases()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
