"""Multi-rank (NCCL) numerics of the training step against the SINGLE-PROCESS GLOBAL-BATCH oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/ddp_check.py [out.json]

Checks (rank 0 writes the JSON verdict, every rank asserts):
  1. all-reduced gradients / world  ==  gradients of the oracle's loss on the concatenated (global) batch, with the
     centering statistics reduced over ranks (softmax centers; Sinkhorn-Knopp prototype sums).  KoLeo is switched off for
     this comparison: it is a per-rank nearest-neighbour term in the reference too (no feature gather), so DDP differs from
     the global batch there by construction.
  2. the DDP loss (mean over ranks) == the oracle's global-batch loss.
  3. after 5 optimisation steps (eager + CUDA-graph schedules, overlapped two-part all-reduce, deterministic grad-norm)
     the fp32 student / teacher arenas, AdamW moments and loss centers are BIT-IDENTICAL on every rank.
"""
import json
import os
import random
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from lightly_train_b200._methods.dinov2.dinov2 import DINOv2, DINOv2AdamWViTArgs, DINOv2Args  # noqa: E402
from oracle import dinov2_oracle as O  # noqa: E402
from tests.golden import recipes as R  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
out_path = sys.argv[1] if len(sys.argv) > 1 else ""
result = {"world": world}


def build(center_method: str, koleo_w: float, drop_path: float = 0.0) -> DINOv2:
    cfg = R.step_config(center_method, center_method != "softmax")
    st = R.det_step_state(cfg, seed=41)
    margs = DINOv2Args(ibot_separate_head=cfg.ibot_separate_head, hidden_dim=cfg.head.hidden_dim,
                       dino_bottleneck_dim=cfg.head.bottleneck_dim, output_dim=cfg.head.out_dim, center_method=center_method,
                       warmup_steps=2, student_freeze_last_layer_steps=1, koleo_loss_weight=koleo_w,
                       teacher_temp_start=0.05, teacher_temp_end=0.05)
    mk = dict(img_size=224, patch_size=16, embed_dim=cfg.vit.embed_dim, depth=cfg.vit.depth, num_heads=cfg.vit.num_heads,
              init_values=cfg.vit.init_values, drop_path_rate=drop_path)
    m = DINOv2(margs, DINOv2AdamWViTArgs(), mk, global_batch_size=3 * world, max_steps=100, device=str(dev))
    m.s_arena.load_from(st["student"]); m.t_arena.load_from(st["teacher"])
    m.dino_loss.center.copy_(st["centers"]["dino"]); m.ibot_loss.center.copy_(st["centers"]["ibot"])
    return m, cfg, st


def rank_inputs(cfg, r: int):
    g = torch.Generator().manual_seed(500 + r)
    B = 3
    views = [torch.randn(B, 3, 224, 224, generator=g) for _ in range(2)] + [torch.randn(B, 3, 96, 96, generator=g) for _ in range(2)]
    masks = torch.rand(2 * B, cfg.vit.num_patches, generator=g) < 0.25
    masks[1] = False
    return views, masks


for center_method in ("softmax", "sinkhorn_knopp"):
    m, cfg, st = build(center_method, koleo_w=0.0)
    views, masks = rank_inputs(cfg, rank)
    idx = masks.flatten().nonzero().flatten()
    w = O.masks_weight_from_masks(masks)
    res = m.training_step_impl({"views": [v.to(dev) for v in views],
                                "masks": {"collated_masks": masks, "mask_indices_list": idx, "masks_weight": w}}, 0)
    m._finish_grad_allreduce()  # the product's own bucketed reduction (head | upper backbone | rest)
    torch.cuda.synchronize()
    loss = res.loss.detach().clone()
    dist.all_reduce(loss)
    loss /= world
    # ---- global-batch oracle on the host (every rank computes it: identical)
    allv = [rank_inputs(cfg, r) for r in range(world)]
    B = 3
    gviews = [torch.cat([allv[r][0][i] for r in range(world)]) for i in range(4)]
    gmasks = torch.cat([allv[r][1][:B] for r in range(world)] + [allv[r][1][B:] for r in range(world)])
    gidx = gmasks.flatten().nonzero().flatten()
    gw = O.masks_weight_from_masks(gmasks)
    student = {k: v.clone().requires_grad_(True) for k, v in st["student"].items()}
    import dataclasses
    ocfg = dataclasses.replace(cfg, koleo_loss_weight=0.0)
    torch.set_num_threads(8)
    out = O.training_step(ocfg, student, st["teacher"], st["centers"], gviews, gmasks, gidx, gw, teacher_temp=0.05, autocast=False)
    out["loss"].backward()
    worst = ("", 0.0)
    for k, p in student.items():
        g = m.s_arena.g(k).float().cpu() / world
        e = (g - p.grad).norm().item() / (p.grad.norm().item() + 1e-12)
        if p.grad.norm().item() > 1e-9 and e > worst[1]:
            worst = (k, e)
    dl = abs(float(loss) - float(out["loss"]))
    result[center_method] = {"ddp_loss": float(loss), "oracle_global_batch_loss": float(out["loss"]), "loss_delta": dl,
                             "worst_grad_rel_err": worst}
    assert dl < 5e-3 * max(1.0, abs(float(out["loss"]))), (center_method, float(loss), float(out["loss"]))
    assert worst[1] < 6e-2, worst

# ---- replica identity after 5 real steps (eager, then graph replay; drop-path on, KoLeo on)
m, cfg, st = build("softmax", koleo_w=0.1, drop_path=0.1)
random.seed(7 + rank)
torch.manual_seed(3 + rank)  # DIFFERENT stochastic-depth draws per rank: only the all-reduce couples the replicas
views, _ = rank_inputs(cfg, rank)
vd = [v.to(dev) for v in views]
for i in range(5):
    m.use_cuda_graph = i >= 2
    m.train_step({"views": vd})
m.dino_loss.apply_center_update(); m.ibot_loss.apply_center_update()
torch.cuda.synchronize()
same = {}
for name, t in (("student_fp32", m.s_arena.fp32), ("teacher_fp32", m.t_arena.fp32), ("exp_avg", m.s_arena.exp_avg),
                ("exp_avg_sq", m.s_arena.exp_avg_sq), ("student_bf16", m.s_arena.bf16.view(torch.int16).float()),
                ("dino_center", m.dino_loss.center), ("ibot_center", m.ibot_loss.center)):
    ref = t.clone()
    dist.broadcast(ref, 0)
    same[name] = bool(torch.equal(ref, t))
flags = torch.tensor([int(all(same.values()))], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
result["replicas_bit_identical_after_5_steps"] = bool(flags.item())
result["per_tensor_rank%d" % rank] = same
assert flags.item() == 1, same
# ---- Sinkhorn-Knopp with the [K] all-reduces captured into the step graph (B200_GRAPH_NCCL=1): graph replay == eager
from lightly_train_b200._methods.dinov2 import dinov2 as _dm  # noqa: E402
if _dm.GRAPH_NCCL:
    losses = {}
    for graphed in (False, True):
        m, cfg, st = build("sinkhorn_knopp", koleo_w=0.1, drop_path=0.0)
        m.use_cuda_graph = graphed
        assert m._graph_ok() == graphed
        random.seed(11 + rank)
        torch.manual_seed(5 + rank)
        ls = []
        for i in range(4):
            ls.append(float(m.train_step({"views": vd}).loss))
        torch.cuda.synchronize()
        losses[graphed] = ls
        if graphed:
            ref = m.s_arena.fp32.clone()
            dist.broadcast(ref, 0)
            ok = torch.tensor([int(torch.equal(ref, m.s_arena.fp32))], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            result["sinkhorn_graph_nccl_replicas_identical"] = bool(ok.item())
            assert ok.item() == 1
            torch.cuda.synchronize()
            m.release_graphs()  # the communicator cannot be destroyed while a graph holds captured NCCL kernels
    result["sinkhorn_graph_nccl_losses_rank%d" % rank] = losses
    for a, b in zip(losses[False], losses[True]):
        assert abs(a - b) < 2e-3 * max(1.0, abs(a)), losses
if rank == 0:
    print("DDP_CHECK", json.dumps(result), flush=True)
    if out_path:
        Path(out_path).write_text(json.dumps(result, indent=1))
dist.barrier()
dist.destroy_process_group()
