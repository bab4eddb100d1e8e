"""Run warm-up steps then ONE eagerly launched training step inside a cudaProfilerStart/Stop range (for
`ncu --profile-from-start off`): cfg2 shapes, bs 64."""
import random
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from lightly_train_b200._methods.dinov2.dinov2 import DINOv2, DINOv2AdamWViTArgs, DINOv2Args  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
random.seed(0)
torch.manual_seed(0)
m = DINOv2(DINOv2Args(), DINOv2AdamWViTArgs(), dict(bench.CONFIGS["cfg2"]["vit"]), global_batch_size=B, device="cuda:0")
batch = {"views": bench.make_views(B, 8, 0, torch.device("cuda:0"))}
for _ in range(2):
    m.train_step(batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m.train_step(batch)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
