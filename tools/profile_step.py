"""Run warm-up steps then ONE training step inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`).

    python tools/profile_step.py [batch] [cfg2|cfg3|cfg5]      (default: cfg2 shapes, the config's batch size)"""
import random
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from lightly_train_b200._methods.dinov2.dinov2 import DINOv2, DINOv2AdamWViTArgs, DINOv2Args  # noqa: E402

cfg = bench.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "cfg2"]
B = int(sys.argv[1]) if len(sys.argv) > 1 and int(sys.argv[1]) > 0 else cfg["batch"]
random.seed(0)
torch.manual_seed(0)
m = DINOv2(DINOv2Args(**cfg["method"]), DINOv2AdamWViTArgs(), dict(cfg["vit"]), global_batch_size=B, device="cuda:0")
if cfg["ckpt"]:
    m.student_embedding_model.wrapped_model.set_activation_checkpointing(True)
batch = {"views": bench.make_views(B, cfg["n_local"], 0, torch.device("cuda:0"), local=cfg["local"])}
for _ in range(2):
    m.train_step(batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m.train_step(batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
