"""Microbench of the loss-path kernels at the cfg2 shapes (K = 65536): fused-LSE CE over the masked-patch rows, two-teacher CE
over the cls rows, row_lse, and the optimizer sweep, timed with CUDA events between L2 flushes: python tools/loss_bench.py
(the `[variant N]` tag in profiles/r02_loss_bench*.log names the CE launch shapes that were A/B-timed with this tool; only the
fastest, variant 0 = 512 threads / one load in flight, is left in csrc/loss.cu)"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from lightly_train_b200 import ops  # noqa: E402

dev = "cuda"
K = 65536


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


g = torch.Generator().manual_seed(0)
M, n_crops, LB = 3721, 128, 512
t_logits = (torch.randn(n_crops + M, K, generator=g) * 0.3).to(dev, torch.bfloat16)
s_logits = (torch.randn(n_crops + LB + M, K, generator=g) * 0.3).to(dev, torch.bfloat16)
ds = torch.empty_like(s_logits)
colterm = torch.randn(K, generator=g).to(dev) * 0.1
rowterm = torch.empty(n_crops + M, device=dev)
loss_rows = torch.empty(s_logits.shape[0], device=dev)
nd = n_crops + LB
idx0 = torch.cat([torch.arange(n_crops), torch.arange(LB) % 64, torch.arange(M)]).to(dev, torch.int32)
idx1 = torch.cat([torch.full((n_crops,), -1), torch.arange(LB) % 64 + 64, torch.full((M,), -1)]).to(dev, torch.int32)
w = torch.rand(s_logits.shape[0], device=dev)
var = os.environ.get("B200_CE_VARIANT", "0")

us = timeit(lambda: ops.row_lse(t_logits[:n_crops], colterm, 25.0, rowterm[:n_crops]))
print(f"[variant {var}] row_lse {n_crops} rows: {us:.1f} us  {2 * n_crops * K / us / 1e3:.0f} GB/s")
us = timeit(lambda: ops.row_lse(t_logits[n_crops:], colterm, 25.0, rowterm[n_crops:]))
print(f"[variant {var}] row_lse {M} rows: {us:.1f} us  {2 * M * K / us / 1e3:.0f} GB/s")
us = timeit(lambda: ops.dino_ce(s_logits[:nd], t_logits[:n_crops], colterm, rowterm[:n_crops], idx0[:nd], idx1[:nd], w[:nd], 10.0, 25.0,
                                loss_rows[:nd], ds[:nd]))
print(f"[variant {var}] dino_ce cls rows ({nd}, 1-2 teachers from L2): {us:.1f} us  {4 * nd * K / us / 1e3:.0f} GB/s (student read + gradient write)")
us = timeit(lambda: ops.dino_ce(s_logits[nd:], t_logits[n_crops:], colterm, rowterm[n_crops:], idx0[nd:], None, w[nd:], 10.0, 25.0,
                                loss_rows[nd:], ds[nd:]))
print(f"[variant {var}] dino_ce masked rows ({M}), teacher LSE precomputed: {us:.1f} us  {6 * M * K / us / 1e3:.0f} GB/s (6 B/elem)")
us = timeit(lambda: ops.dino_ce(s_logits[nd:], t_logits[n_crops:], colterm, None, idx0[nd:], None, w[nd:], 10.0, 25.0, loss_rows[nd:], ds[nd:]))
print(f"[variant {var}] dino_ce masked rows ({M}), teacher LSE FUSED: {us:.1f} us  {6 * M * K / us / 1e3:.0f} GB/s (6 B/elem)")
