#!/usr/bin/env python
"""bench.py -- images/sec of the full DINOv2 training step (ViT-S/16, 2x224^2 + 8x96^2 crops, bs 64/GPU).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (this repo's kernels)
    python bench.py --impl reference --steps K --warmup W     # reference arm: the reference algorithm on host cores

One "step" = teacher forward, student forward+backward (global + local crops), DINO/iBOT/KoLeo losses,
gradient all-reduce (N>1), clip + AdamW + EMA teacher.  Synthetic N(0,1) crops, random-init weights.
`value`  : whole-job images/s with inputs resident in HBM when the timed region starts.
`e2e`    : the same step through the public API with HOST (pinned) crops: H2D copy of every step's views and
           a D2H read of every step's loss inside the timed region (double-buffered on a copy stream).
`roofline`: tcgen05 GEMM launches of one step timed with CUDA events on the launch stream.
`cpu_baseline`: the oracle (CPU port of the reference path) timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

VIT_S16 = dict(img_size=224, patch_size=16, embed_dim=384, depth=12, num_heads=6, init_values=1e-5, drop_path_rate=0.1)
PER_GPU_BATCH = 64
N_LOCAL = 8


def make_views(batch: int, n_local: int, seed: int, device, pin: bool = False):
    g = torch.Generator().manual_seed(seed)
    views = [torch.randn(batch, 3, 224, 224, generator=g) for _ in range(2)]
    views += [torch.randn(batch, 3, 96, 96, generator=g) for _ in range(n_local)]
    if device is not None:
        return [v.to(device) for v in views]
    if pin:
        return [v.pin_memory() for v in views]
    return views


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index: int) -> None:
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._idx = index
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self) -> None:
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self._idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=5)

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------------- CPU port
def cpu_reference_step_time(batch: int, steps: int, warmup: int, threads: int):
    """The reference algorithm (oracle port: teacher fwd, student fwd+bwd via autograd, losses, clip, AdamW, EMA)
    on the host cores, fp32, ViT-S/16 + 65536-way heads, `batch` images per step."""
    from oracle import dinov2_oracle as O
    from tests.golden import recipes as R

    torch.set_num_threads(threads)
    vit = O.ViTConfig(embed_dim=384, depth=12, num_heads=6, patch_size=16, img_size=224, init_values=1e-5)
    head = O.HeadConfig(in_dim=384, hidden_dim=2048, bottleneck_dim=256, out_dim=65536)
    cfg = O.StepConfig(vit=vit, head=head)
    st = R.det_step_state(cfg, seed=7)
    student = {k: v.clone().requires_grad_(True) for k, v in st["student"].items()}
    teacher = st["teacher"]
    m_state = {k: torch.zeros_like(v) for k, v in student.items()}
    v_state = {k: torch.zeros_like(v) for k, v in student.items()}
    views = make_views(batch, N_LOCAL, 123, None)
    g = torch.Generator().manual_seed(5)
    masks = torch.rand(2 * batch, 196, generator=g) < 0.15
    masks[batch:] = False
    idx = masks.flatten().nonzero().flatten()
    w = O.masks_weight_from_masks(masks)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for p in student.values():
            p.grad = None
        out = O.training_step(cfg, student, teacher, st["centers"], views, masks, idx, w, teacher_temp=0.04)
        out["loss"].backward()
        with torch.no_grad():
            grads = [p.grad for p in student.values()]
            O.clip_grad_norm(grads, 3.0)
            for k, p in student.items():
                hp = O.param_hparams(k[len("backbone."):] if k.startswith("backbone.") else k, k.startswith("backbone."),
                                     1e-3, 0.04, 12)
                O.adamw_step(p, p.grad, m_state[k], v_state[k], it + 1, hp["lr"], hp["weight_decay"])
            O.update_ema(list(student.values()), [teacher[k] for k in student], 0.992)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def host_threads() -> int:
    """Threads the CPU arm uses: the cores this process may run on, capped at 16 (the oracle's small-matrix torch
    ops slow down badly when oversubscribed across a 128-core host)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 16))


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    batch = 4
    sec = cpu_reference_step_time(batch, args.steps, args.warmup, threads)
    value = batch / sec
    line = {
        "impl": "reference", "metric": "images/sec ViT-S/16 DINOv2 training step (2g+8l crops, bs64/GPU)", "value": value,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": "cfg2: ViT-S/16 DINOv2, 2x224^2 + 8x96^2 crops, bs=64/GPU, K=65536 shared DINO/iBOT head, "
                               "softmax centering, drop_path 0.1, full step incl. clip+AdamW+EMA",
                   "global_batch": 64 * args.gpus, "parallelism": f"dp{args.gpus}",
                   "sample": "each timed step is a bounded sample of that workload: %d images (of 64) through the same "
                             "step on the host cores; images/s is batch-size independent on the CPU" % batch},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"full step at bs={batch} (reference algorithm restated in oracle/, torch CPU fp32)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- B200 arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from the host instead of CUDA-graph replay")
    ap.add_argument("--gemm-profile", default="", help="write the per-shape GEMM timing table of one step to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist

    from lightly_train_b200 import _lib, ops
    from lightly_train_b200._methods.dinov2.dinov2 import DINOv2, DINOv2AdamWViTArgs, DINOv2Args

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version banner must not precede the JSON line on stdout
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    W = max(args.warmup, 3)
    K = args.steps
    B = args.batch

    import random
    random.seed(1000 + rank)
    torch.manual_seed(0)
    method = DINOv2(DINOv2Args(), DINOv2AdamWViTArgs(), VIT_S16, global_batch_size=B * world, max_steps=125_000, device=str(dev))
    method.use_cuda_graph = not args.eager
    if world > 1:  # identical initial weights on every rank
        dist.broadcast(method.s_arena.fp32, 0)
        dist.broadcast(method.t_arena.fp32, 0)
        method.s_arena.bf16_valid = method.t_arena.bf16_valid = False
    # two distinct resident batches (134 MB each, larger than L2) alternate between steps
    batches = [{"views": make_views(B, N_LOCAL, 1000 * rank + i, dev)} for i in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        method.train_step(batches[i % 2])
    barrier()
    l0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        h0 = time.perf_counter()
        for i in range(K):
            res = method.train_step(batches[i % 2])
        host_ms = (time.perf_counter() - h0) / K * 1e3  # host time to enqueue one step (must stay below ms_per_step)
        e1.record()
        barrier()
    launches = _lib.LAUNCHES - l0
    ms = e0.elapsed_time(e1) / K
    loss_val = float(res.loss)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    value = B * world / (ms / 1e3)

    # ---- e2e: host (pinned) crops -> device every step on a copy stream, loss read back every step
    e2e = None
    if not args.no_e2e:
        host = [make_views(B, N_LOCAL, 2000 * rank + i, None, pin=True) for i in range(2)]
        h2d_bytes = sum(v.numel() * 4 for v in host[0])
        copy_stream = torch.cuda.Stream(device=dev)
        dev_bufs = [[torch.empty_like(v, device=dev) for v in host[0]] for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.zeros(2).pin_memory()
        loss_done = [torch.cuda.Event() for _ in range(2)]
        losses: list = []

        def prefetch(i: int) -> None:
            slot = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                for d, h in zip(dev_bufs[slot], host[slot]):
                    d.copy_(h, non_blocking=True)
                ready[slot].record(copy_stream)

        def run_e2e(n: int) -> None:
            for s in range(2):
                consumed[s].record(torch.cuda.current_stream())
            prefetch(0)
            for i in range(n):
                slot = i % 2
                if i + 1 < n:
                    prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[slot])
                r = method.train_step({"views": dev_bufs[slot]})
                consumed[slot].record(torch.cuda.current_stream())
                # D2H of every step's loss into pinned memory; the host waits for it one step later (after the next
                # step has been enqueued), as a logging trainer does, so the device never idles on the read-back
                loss_host[slot:slot + 1].copy_(r.loss.reshape(1), non_blocking=True)
                loss_done[slot].record(torch.cuda.current_stream())
                if i >= 1:
                    loss_done[1 - slot].synchronize()
                    losses.append(float(loss_host[1 - slot]))
            loss_done[(n - 1) % 2].synchronize()
            losses.append(float(loss_host[(n - 1) % 2]))

        run_e2e(2)
        barrier()
        e0.record()
        run_e2e(K)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / K], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": B * world / (float(t[0]) / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 4}

    # ---- roofline: every tcgen05 GEMM launch of one step timed with CUDA events on the launch stream
    roofline = None
    # every rank runs this extra step (it contains the gradient all-reduce); rank 0 reports
    ops.GEMM_PROFILE = []
    method.use_cuda_graph = False  # events cannot be recorded inside a graph replay: time the eager schedule
    torch.cuda._sleep(int(3e8))    # ~150 ms head start for the host, so event pairs bracket kernels, not launch gaps
    # calibration: the same event pair around a one-CTA kernel of the library.  An event pair brackets the launch dispatch
    # (the front end cannot overlap it with the preceding kernel once an event sits in between) as well as the kernel.
    cal_buf = torch.zeros(32, device=dev)
    cal = []
    for _ in range(24):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(); ops.fill_f32(cal_buf, 0.0); c1.record()
        cal.append((c0, c1))
    method.train_step(batches[0])
    torch.cuda.synchronize()
    cal_us = sorted(a.elapsed_time(b) * 1e3 for a, b in cal)[len(cal) // 2]
    method.use_cuda_graph = not args.eager
    prof, ops.GEMM_PROFILE = ops.GEMM_PROFILE, None
    if rank == 0:
        if args.gemm_profile:
            table = {}
            for (f, a, b), key in zip(prof, ops.GEMM_PROFILE_KEYS):
                t = table.setdefault(key, [0, 0.0, 0.0])
                t[0] += 1; t[1] += a.elapsed_time(b); t[2] += f
            rows = sorted(((k, v) for k, v in table.items()), key=lambda kv: -kv[1][1])
            with open(args.gemm_profile, "w") as fh:
                fh.write("M,N,K,a_mn,b_mn,epi,splits,launches,total_ms,TFLOP/s\n")
                for k, v in rows:
                    fh.write(",".join(map(str, k)) + f",{v[0]},{v[1]:.4f},{v[2] / (v[1] * 1e-3) / 1e12:.1f}\n")
        flops = sum(f for f, _, _ in prof)
        gemm_ms_events = sum(a.elapsed_time(b) for _, a, b in prof)
        # per-launch dispatch overhead inside an event pair = pair time of the one-CTA kernel minus its own ~2 us run time
        overhead_us = max(0.0, cal_us - 2.0)
        gemm_ms = gemm_ms_events - len(prof) * overhead_us * 1e-3
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        ach = flops / (gemm_ms * 1e-3) / 1e12
        traffic = None
        tf = ROOT / "profiles" / "r01_gemm_traffic.json"  # dram bytes per launch from the committed `ncu --set full` capture
        if tf.exists():
            traffic = json.loads(tf.read_text()).get("mean_dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all launches of one step)", "achieved": ach,
                    "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                    "traffic_source": "profiles/r01_gemm_traffic.json: mean dram read+write bytes per launch over the 5 block GEMMs",
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
                    "timed": "CUDA events around every b200_gemm launch of one eagerly launched step (same kernels as the graph "
                             "replay), minus the launch-dispatch overhead an event pair adds, calibrated in the same step on a "
                             "one-CTA kernel (event_pair_empty_us - 2 us of run time) per launch",
                    "gemm_launches": len(prof), "gemm_ms_per_step": gemm_ms, "gemm_ms_per_step_events_raw": gemm_ms_events,
                    "event_pair_empty_us": cal_us, "achieved_raw": flops / (gemm_ms_events * 1e-3) / 1e12,
                    "gemm_tflop_per_step": flops / 1e12}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N=1 only (the other ranks would idle)
        threads = host_threads()
        sec = cpu_reference_step_time(2, 1, 1, threads)
        cpu_baseline = {"value": 2 / sec, "unit": "images/s", "cores": threads, "kind": "port",
                        "sample": "full step (fwd+bwd+AdamW+EMA) of ViT-S/16 + K=65536 heads at bs=2, oracle/ on torch CPU fp32, 1 warm-up + 1 timed"}

    if rank == 0:
        line = {
            "metric": "images/sec ViT-S/16 DINOv2 training step (2g+8l crops, bs64/GPU)", "value": value, "unit": "images/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "cfg2: ViT-S/16 DINOv2, 2x224^2 + 8x96^2 crops, bs=%d/GPU, K=65536 shared DINO/iBOT head, "
                                   "softmax centering, drop_path 0.1, full step incl. clip+AdamW+EMA" % B,
                       "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2_policy": "inputs (2 alternating 134 MB batches) and activations (>8 GB/step) exceed the 126 MB L2"},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": launches, "gpu_launches_scope": "library kernel launches of rank 0 (graph replays counted per captured launch)", "host_ms_per_step": round(host_ms, 3), "loss": loss_val,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
