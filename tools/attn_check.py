"""GPU bring-up / A-B tool for the tcgen05 attention kernels (csrc/attention_tc.cu): numerics against an fp32 torch
statement and against the warp-level kernels, then CUDA-event timings of both implementations at the step's shape.

    python tools/attn_check.py fwd|bwd|time     (each mode in its own process: a device trap poisons the context)
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from lightly_train_b200 import ops  # noqa: E402

dev = "cuda"


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev, dtype)


def ref_fwd_bwd(qkv, do, B, N, h):
    D = h * 64
    q5 = qkv.float().view(B, N, 3, h, 64).permute(2, 0, 3, 1, 4).clone().requires_grad_(True)
    s = (q5[0] * 0.125) @ q5[1].transpose(-1, -2)
    o = (s.softmax(-1) @ q5[2]).transpose(1, 2).reshape(B * N, D)
    o.backward(do.float())
    return o.detach(), torch.logsumexp(s, -1).detach(), q5.grad.permute(1, 3, 0, 2, 4).reshape(B * N, 3 * D)


def check_fwd():
    for (B, N, h) in [(3, 197, 6), (2, 201, 3), (2, 256, 3), (2, 130, 2), (5, 37, 2), (7, 54, 3), (1, 16, 1), (64, 197, 6)]:
        D = h * 64
        qkv = rnd(B * N, 3 * D, dtype=torch.bfloat16, seed=10)
        out, ref = (torch.zeros(B * N, D, device=dev, dtype=torch.bfloat16) for _ in range(2))
        lse, lse_ref = (torch.zeros(B * h, N, device=dev) for _ in range(2))
        ops.attention_fwd_tc(qkv, B, N, h, out, lse, 0.125)
        torch.cuda.synchronize()
        ops.TC_ATTENTION_FWD = ops.TC_ATTENTION_PACKED = False  # reference: the warp-level kernel for every length
        ops.attention_fwd(qkv, B, N, h, ref, lse_ref, 0.125)
        o, l, _ = ref_fwd_bwd(qkv, torch.zeros(B * N, D, device=dev), B, N, h)
        print(f"fwd B={B} N={N} h={h}: |out-torch| {(out.float() - o).abs().max().item():.4f}  |out-warp| "
              f"{(out.float() - ref.float()).abs().max().item():.4f}  |lse-torch| {(lse.view(B, h, N) - l).abs().max().item():.4f}  "
              f"|lse-warp| {(lse - lse_ref).abs().max().item():.2e}", flush=True)


def check_bwd():
    for (B, N, h) in [(3, 197, 6), (2, 201, 3), (2, 130, 2), (5, 37, 2), (7, 54, 3), (2, 208, 1), (64, 197, 6)]:
        D = h * 64
        qkv = rnd(B * N, 3 * D, dtype=torch.bfloat16, seed=10)
        do = rnd(B * N, D, dtype=torch.bfloat16, seed=11)
        out = torch.zeros(B * N, D, device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(B * h, N, device=dev)
        ops.TC_ATTENTION_FWD = ops.TC_ATTENTION_BWD = ops.TC_ATTENTION_PACKED = False
        ops.attention_fwd(qkv, B, N, h, out, lse, 0.125)
        dq_w = torch.zeros_like(qkv)
        cs_w = torch.zeros(3 * D, device=dev)
        ops.attention_bwd(qkv, out, do, lse, B, N, h, dq_w, 0.125, colsum=cs_w)
        dq_t = torch.zeros_like(qkv)
        cs_t = torch.zeros(3 * D, device=dev)
        ops.attention_bwd_tc(qkv, out, do, lse, B, N, h, dq_t, 0.125, colsum=cs_t)
        torch.cuda.synchronize()
        _, _, want = ref_fwd_bwd(qkv, do, B, N, h)
        for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
            a, w, t = dq_t[:, sl].float(), dq_w[:, sl].float(), want[:, sl]
            print(f"bwd B={B} N={N} h={h} {name}: rel(tc,torch) {((a - t).norm() / t.norm()).item():.4f}  rel(warp,torch) "
                  f"{((w - t).norm() / t.norm()).item():.4f}  max|tc-warp| {(a - w).abs().max().item():.4f}", flush=True)
        ref_cs = dq_t.float().sum(0)
        print(f"   colsum: max|cs_tc - sum(dqkv_tc)| {(cs_t - ref_cs).abs().max().item():.4f} (scale {ref_cs.abs().max().item():.2f})  "
              f"max|cs_tc - cs_warp| {(cs_t - cs_w).abs().max().item():.4f}", flush=True)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def bench():
    for (B, N, h) in [(128, 197, 6), (512, 37, 6), (64, 197, 12), (32, 197, 16)]:
        D = h * 64
        qkv = rnd(B * N, 3 * D, dtype=torch.bfloat16, seed=10)
        do = rnd(B * N, D, dtype=torch.bfloat16, seed=11)
        out = torch.zeros(B * N, D, device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(B * h, N, device=dev)
        dq = torch.zeros_like(qkv)
        cs = torch.zeros(3 * D, device=dev)
        res = {}
        for tc in (False, True):
            ops.TC_ATTENTION_FWD = ops.TC_ATTENTION_BWD = ops.TC_ATTENTION_PACKED = tc
            res[("fwd", tc)] = timeit(lambda: ops.attention_fwd(qkv, B, N, h, out, lse, 0.125))
            res[("bwd", tc)] = timeit(lambda: ops.attention_bwd(qkv, out, do, lse, B, N, h, dq, 0.125, colsum=cs))
        fl = 4.0 * N * N * 64 * B * h
        print(f"time B={B} N={N} h={h} ({B * h} pairs): fwd warp {res[('fwd', False)]:.1f} us  tc {res[('fwd', True)]:.1f} us "
              f"({fl / res[('fwd', True)] / 1e6:.0f} TFLOP/s) | bwd warp {res[('bwd', False)]:.1f} us  tc {res[('bwd', True)]:.1f} us "
              f"({2.5 * fl / res[('bwd', True)] / 1e6:.0f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    {"fwd": check_fwd, "bwd": check_bwd, "time": bench}[sys.argv[1]]()
