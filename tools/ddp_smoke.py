"""2-rank NCCL smoke of the training step (tiny model): eager steps, then CUDA-graph steps, with progress prints.
Launch: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/ddp_smoke.py"""
import os, random, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from lightly_train_b200._methods.dinov2.dinov2 import DINOv2, DINOv2AdamWViTArgs, DINOv2Args

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)


def log(*a):
    print(f"[rank {rank} t={time.time() % 1000:.1f}]", *a, flush=True)


random.seed(100 + rank)
torch.manual_seed(0)
margs = DINOv2Args(hidden_dim=256, dino_bottleneck_dim=64, output_dim=512)
mk = dict(img_size=224, patch_size=16, embed_dim=128, depth=2, num_heads=2, init_values=1e-5, drop_path_rate=0.1)
m = DINOv2(margs, DINOv2AdamWViTArgs(), mk, global_batch_size=4 * world, max_steps=100, device=str(dev))
dist.broadcast(m.s_arena.fp32, 0); dist.broadcast(m.t_arena.fp32, 0)
m.s_arena.bf16_valid = m.t_arena.bf16_valid = False
g = torch.Generator().manual_seed(rank)
views = [torch.randn(4, 3, 224, 224, generator=g).to(dev) for _ in range(2)] + [torch.randn(4, 3, 96, 96, generator=g).to(dev) for _ in range(2)]
log("init done")
for mode in ("eager", "graph"):
    m.use_cuda_graph = mode == "graph"
    for i in range(4):
        res = m.train_step({"views": views})
        torch.cuda.synchronize()
        log(mode, "step", i, "loss", float(res.loss))
# replicas must stay identical: compare a checksum of the student weights across ranks
cs = m.s_arena.fp32.double().sum().reshape(1)
lst = [torch.zeros_like(cs) for _ in range(world)]
dist.all_gather(lst, cs)
log("weight checksums", [float(x) for x in lst])
assert all(abs(float(x) - float(lst[0])) < 1e-6 * abs(float(lst[0])) for x in lst), "replicas diverged"
dist.destroy_process_group()
log("OK")
