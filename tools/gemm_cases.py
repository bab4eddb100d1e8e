"""Representative GEMM launches of the cfg2 step for `ncu --set full` (one launch per case after warm-up)."""
import sys
import torch
sys.path.insert(0, ".")
from lightly_train_b200 import ops

dev = "cuda"
T, D, H = 25216, 384, 1536
bf = torch.bfloat16
x = torch.randn(T, D, device=dev).to(bf)
hid = torch.randn(T, H, device=dev).to(bf)
wqkv = (torch.randn(3 * D, D, device=dev) * 0.02).to(bf)
wproj = (torch.randn(D, D, device=dev) * 0.02).to(bf)
w1 = (torch.randn(H, D, device=dev) * 0.02).to(bf)
w2 = (torch.randn(D, H, device=dev) * 0.02).to(bf)
b3, b1, bh = torch.randn(3 * D, device=dev), torch.randn(D, device=dev), torch.randn(H, device=dev)
gamma = torch.randn(D, device=dev)
res = torch.randn(T, D, device=dev)
out_qkv = torch.empty(T, 3 * D, device=dev, dtype=bf)
out_res, out_o = torch.empty(T, D, device=dev), torch.empty(T, D, device=dev, dtype=bf)
out_h, out_u = torch.empty(T, H, device=dev, dtype=bf), torch.empty(T, H, device=dev, dtype=bf)
out_du = torch.empty(T, H, device=dev, dtype=bf)


B_, N_, h_ = 128, 197, 6
qkv_att = torch.randn(B_ * N_, 3 * D, device=dev).to(bf)
att_out = torch.empty(B_ * N_, D, device=dev, dtype=bf)
att_lse = torch.empty(B_ * h_, N_, device=dev)
datt = torch.randn(B_ * N_, D, device=dev).to(bf)
dqkv = torch.empty_like(qkv_att)


import os
ATT = os.environ.get("CASES", "all") in ("all", "att")
GEMM_PLAIN = os.environ.get("CASES", "all") in ("all",)


def cases():
    if ATT:
        ops.attention_fwd(qkv_att, B_, N_, h_, att_out, att_lse, 0.125)
        ops.attention_bwd(qkv_att, att_out, datt, att_lse, B_, N_, h_, dqkv, 0.125)
    if os.environ.get("CASES", "all") == "att":
        return
    if GEMM_PLAIN:
        ops.gemm(x, wqkv, out_qkv, bias=b3, ws_mode=2)                                       # qkv generic schedule
    ops.gemm(x, wqkv, out_qkv, bias=b3)                                                      # qkv
    ops.gemm(x, wproj, out_res, epi=ops.EPI_RESIDUAL, bias=b1, out2=out_o, aux=res, gamma=gamma)  # proj
    ops.gemm(x, w1, out_h, epi=ops.EPI_BIAS_GELU, bias=bh, out2=out_u)                        # fc1
    ops.gemm(hid, w2, out_res, epi=ops.EPI_RESIDUAL, bias=b1, out2=out_o, aux=res, gamma=gamma)   # fc2
    if os.environ.get("CASES", "all") == "epi":
        ops.gemm(x, w1, out_h, epi=ops.EPI_BIAS_GELU_DG, bias=bh, out2=out_u)                 # fc1 student (h, gelu'(u))
        ops.gemm(x, w2, out_du, b_mn=True, epi=ops.EPI_MUL_AUX, aux=out_u)                    # dU = dH * gelu'(u)
        return
    ops.gemm(x, w2, out_du, b_mn=True, epi=ops.EPI_DGELU, aux=out_u)                          # dU dgrad


for _ in range(3):
    cases()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
cases()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
