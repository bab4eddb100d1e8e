"""-m gpu: device-side block-wise mask generation (csrc/masks.cu) -- statistical parity with the
reference's host generator (same algorithm, different RNG) and exact structural invariants; then a full step driven by
device masks (eager and CUDA-graph replay)."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs CUDA", allow_module_level=True)

from lightly_train_b200 import ops  # noqa: E402
from lightly_train_b200._methods.dinov2.utils import MaskingGenerator  # noqa: E402

dev = "cuda"


def _host_masks(targets, H, W):
    gen = MaskingGenerator(input_size=(H, W), max_num_patches=int(0.5 * H * W))
    return np.stack([gen(t) for t in targets]).reshape(len(targets), -1)


def test_device_masks_statistics_match_host_generator():
    H = W = 14
    Np = H * W
    random.seed(0)
    n = 4096
    targets = [int(Np * random.uniform(0.1, 0.5)) if i % 2 == 0 else 0 for i in range(n)]
    t = torch.tensor(targets, dtype=torch.int32, device=dev)
    masks = torch.empty(n, Np, device=dev, dtype=torch.uint8)
    ops.block_masks(t, H, W, int(0.5 * Np), 1234, masks)
    dm = masks.cpu().numpy().astype(bool)
    hm = _host_masks(targets, H, W)
    cnt_d, cnt_h, tg = dm.sum(1), hm.sum(1), np.array(targets)
    assert (cnt_d <= tg).all() and (cnt_d[tg == 0] == 0).all()
    # the generator stops short only when 10 proposals in a row fail: equally rare on both sides
    assert abs((cnt_d == tg).mean() - (cnt_h == tg).mean()) < 0.02 and (cnt_d == tg).mean() > 0.9
    assert abs(cnt_d.sum() / cnt_h.sum() - 1) < 0.01
    # per-cell masking probability map (centre cells are masked more often than corners: rectangles must fit)
    pd, ph = dm[tg > 0].mean(0), hm[tg > 0].mean(0)
    assert np.abs(pd - ph).max() < 0.07, np.abs(pd - ph).max()  # 2 x 2048 crops: the max over 196 cells of a difference with sigma 0.016
    assert np.abs(pd - ph).mean() < 0.015
    # blockiness: fraction of horizontally adjacent cell pairs that differ (random dots would give ~2p(1-p) = 0.4)
    def edges(m):
        g = m.reshape(-1, H, W)
        return (g[:, :, 1:] != g[:, :, :-1]).mean()
    assert abs(edges(dm[tg > 0]) - edges(hm[tg > 0])) < 0.01
    # different (seed, step) -> different masks; same -> identical
    m2, m3 = torch.empty_like(masks), torch.empty_like(masks)
    ops.block_masks(t, H, W, int(0.5 * Np), 1234, m2)
    step = torch.ones(1, device=dev, dtype=torch.int32)
    ops.block_masks(t, H, W, int(0.5 * Np), 1234, m3, step_dev=step)
    assert torch.equal(m2, masks) and not torch.equal(m3, masks)


def test_collate_matches_reference_collation():
    B, Np, cap = 12, 196, 1024
    g = torch.Generator().manual_seed(0)
    masks = (torch.rand(B, Np, generator=g) < 0.2)
    masks[3] = False
    mu8 = masks.to(dev, torch.uint8)
    idx = torch.full((cap,), -1, device=dev, dtype=torch.int64)
    w, rw, pad = (torch.full((cap,), 7.0, device=dev) for _ in range(3))
    mv = torch.zeros(1, device=dev, dtype=torch.int32)
    ops.collate_masks(mu8, cap, idx, w, rw, pad, mv)
    M = int(masks.sum())
    want_idx = masks.flatten().nonzero().flatten()
    want_w = (1 / masks.sum(-1).clamp(min=1.0)).unsqueeze(-1).expand_as(masks)[masks]  # utils.py:141-149
    assert int(mv) == M
    assert torch.equal(idx[:M].cpu(), want_idx) and (idx[M:] == 0).all()
    torch.testing.assert_close(w[:M].cpu(), want_w.float())
    assert (w[M:] == 0).all() and torch.allclose(rw[:M], torch.full((M,), 1.0 / M, device=dev)) and (rw[M:] == 0).all()
    assert (pad[:M] == 0).all() and (pad[M:] < -1e29).all()


@pytest.mark.parametrize("graph", [False, True])
def test_step_with_device_masks(graph):
    from lightly_train_b200._methods.dinov2.dinov2 import DINOv2, DINOv2AdamWViTArgs, DINOv2Args
    from tests.golden import recipes as R

    cfg = R.step_config("softmax", False)
    margs = DINOv2Args(hidden_dim=cfg.head.hidden_dim, dino_bottleneck_dim=cfg.head.bottleneck_dim, output_dim=cfg.head.out_dim)
    mk = dict(img_size=224, patch_size=16, embed_dim=cfg.vit.embed_dim, depth=cfg.vit.depth, num_heads=cfg.vit.num_heads,
              init_values=cfg.vit.init_values, drop_path_rate=0.0)
    m = DINOv2(margs, DINOv2AdamWViTArgs(), mk, 8, 3, max_steps=10, device=dev)
    m.mask_source = "device"
    m.use_cuda_graph = graph
    views, _, _, _ = R.step_case_inputs(cfg, batch=4)
    random.seed(1)
    losses = [float(m.train_step({"views": [v.to(dev) for v in views]}).loss) for _ in range(3)]
    assert all(np.isfinite(losses))
    assert int(m._mask_step_dev) == 3  # one RNG stream per step
