"""CPU: dry run of the host-side schedules with every C-ABI call stubbed out (no arithmetic, CPU tensors): the python
control flow of one optimisation step -- eager and 'graph replay', split backward with the overlapped all-reduce buckets,
batch-subset stochastic depth with and without activation checkpointing, Sinkhorn-Knopp, device masks, the Lightning-shaped
hook loop, checkpoint save / load, the DINOv3 teacher and the DistillationV3 step -- must run without shape, argument or
name errors.  (The arithmetic itself is checked on the GPU: tests/test_*_gpu.py.)"""
import contextlib
import random

import pytest
import torch


class _FakeLib:
    def __init__(self, calls):
        self.calls = calls

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls[name] = self.calls.get(name, 0) + 1
            return 0
        return f


class _S:
    cuda_stream = 0

    def wait_stream(self, o): pass

    def wait_event(self, e): pass


class _E:
    def record(self, s=None): pass


class _G:
    def replay(self): pass

    def reset(self): pass

    def pool(self): return 0


@pytest.fixture
def stubbed(monkeypatch):
    from lightly_train_b200 import _lib, ops
    calls = {}
    fake = _FakeLib(calls)
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    monkeypatch.setattr(ops, "_L", lambda: fake)
    monkeypatch.setattr(ops, "_req_cuda", lambda *a: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    # the one stubbed call whose OUTPUT steers host-side indexing: keep its indices in range
    monkeypatch.setattr(ops, "random_subset", lambda n, k, seed, counter, idx: idx.copy_(torch.arange(k)))
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _S())
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: _S())
    monkeypatch.setattr(torch.cuda, "Event", lambda *a, **k: _E())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", lambda *a, **k: _G())
    monkeypatch.setattr(torch.cuda, "graph", lambda *a, **k: contextlib.nullcontext())
    return calls


def _method(depth, dpr, uniform=False, ckpt=False, center="softmax", sep=False, mask_source="host", graph=False):
    from lightly_train_b200._methods.dinov2.dinov2 import DINOv2, DINOv2AdamWViTArgs, DINOv2Args
    m = DINOv2(DINOv2Args(output_dim=512, hidden_dim=256, dino_bottleneck_dim=64, ibot_bottleneck_dim=64, center_method=center,
                          ibot_separate_head=sep), DINOv2AdamWViTArgs(),
               dict(img_size=64, patch_size=16, embed_dim=128, depth=depth, num_heads=2, init_values=1e-5, drop_path_rate=dpr,
                    drop_path_uniform=uniform), 8, 3, max_steps=10, device="cpu")
    m._side_stream, m._comm_stream, m._head_ready, m._mid_ready = _S(), _S(), _E(), _E()
    if ckpt:
        m.student_embedding_model.wrapped_model.set_activation_checkpointing(True)
    m.mask_source = mask_source
    m.use_cuda_graph = graph
    m.force_backbone_split = depth // 2 if depth >= 4 else None  # exercise the cut backward on one rank too
    return m


CASES = [dict(depth=2, dpr=0.0), dict(depth=4, dpr=0.3), dict(depth=4, dpr=0.3, uniform=True, ckpt=True),
         dict(depth=6, dpr=0.1, center="sinkhorn_knopp", sep=True), dict(depth=4, dpr=0.0, mask_source="device"),
         dict(depth=4, dpr=0.0, graph=True), dict(depth=6, dpr=0.3, graph=True, mask_source="device", ckpt=True),
         dict(depth=4, dpr=0.0, graph=True, center="sinkhorn_knopp", sep=True)]


@pytest.mark.parametrize("kw", CASES, ids=lambda k: "-".join(f"{a}{b}" for a, b in k.items()))
def test_step_schedules_run(stubbed, kw):
    m = _method(**kw)
    views = [torch.randn(4, 3, 64, 64) for _ in range(2)] + [torch.randn(4, 3, 32, 32) for _ in range(2)]
    random.seed(0)
    for _ in range(2):
        m.train_step({"views": views})
    # the reference's hook names in Lightning's order
    (opt,), (sch,) = m.configure_optimizers()
    m.training_step({"views": views}, 0)
    m.on_before_optimizer_step(opt)
    m.configure_gradient_clipping(opt)
    opt.step()
    sch["scheduler"].step()
    m.trainer.global_step += 1
    m.on_train_batch_end(None, None, 0)
    assert m.trainer.global_step == 3 and "train_loss" in m.logged
    ck = m.checkpoint()
    m.load_checkpoint(ck)
    assert not m.s_arena.bf16_valid and not m.t_arena.bf16_valid
    assert stubbed.get("b200_gemm", 0) > 0 and stubbed.get("b200_adamw_ema", 0) == 3
    if kw.get("dpr", 0) > 0.1:
        assert stubbed.get("b200_copy_samples", 0) > 0      # compact batch-subset schedule
    if kw.get("mask_source") == "device":
        assert stubbed.get("b200_block_masks", 0) == 3 and stubbed.get("b200_collate_masks", 0) == 3
    if kw.get("depth", 0) >= 4 and not kw.get("graph"):
        pass  # split backward exercised (depth >= 4 cuts at depth / 2)


def test_layernorm_prologue_gemm_route(stubbed, monkeypatch):
    """B200_LN_GEMM=1: norm1 -> qkv and norm2 -> fc1 go through b200_ln_gemm (embed dim 128 <= 384, multiple of 64) in the
    teacher and the dense student blocks; the policy helper refuses unsupported widths and the "auto" multi-wave case."""
    from lightly_train_b200._models import dinov2_vit as V
    monkeypatch.setattr(V, "LN_GEMM", "1")
    monkeypatch.setattr(V, "_ln_gemm_wanted", lambda T, D, dev: V.LN_GEMM != "0" and D % 64 == 0 and D <= 384)  # cpu tensors here
    m = _method(depth=2, dpr=0.0)
    views = [torch.randn(4, 3, 64, 64) for _ in range(2)] + [torch.randn(4, 3, 32, 32) for _ in range(2)]
    random.seed(0)
    m.train_step({"views": views})
    # per block 2 fused launches; teacher (global) + student global + student local = 3 passes x 2 blocks
    assert stubbed.get("b200_ln_gemm", 0) == 3 * 2 * 2
    monkeypatch.undo()
    assert V._ln_gemm_wanted(1000, 768, torch.device("cpu")) is False


def test_release_graphs_and_teardown_hook(stubbed, monkeypatch):
    """release_graphs() drops the captured graphs; the process-group teardown hook releases every registered holder first."""
    from lightly_train_b200._methods.dinov2 import dinov2 as Dm
    m = _method(depth=2, dpr=0.0, graph=True)
    views = [torch.randn(4, 3, 64, 64) for _ in range(2)] + [torch.randn(4, 3, 32, 32) for _ in range(2)]
    random.seed(0)
    m.train_step({"views": views})
    assert m._static is not None and len(m._static["graphs"]) == 1
    Dm._NCCL_GRAPH_HOLDERS.add(m)
    Dm._release_all_nccl_graphs()
    assert m._static is None
    m.train_step({"views": views})  # re-captures on demand
    assert m._static is not None


def test_world_two_allreduce_buckets(stubbed, monkeypatch):
    """With two ranks the gradient arena goes out in three buckets: heads, blocks >= depth/2 + norm, the rest."""
    import torch.distributed as dist
    slices = []

    class _W:
        def wait(self): pass

    def fake_all_reduce(t, async_op=False, op=None):
        slices.append((t.data_ptr(), t.numel()))
        return _W() if async_op else None

    monkeypatch.setattr(dist, "is_available", lambda: True)
    monkeypatch.setattr(dist, "is_initialized", lambda: True)
    monkeypatch.setattr(dist, "get_world_size", lambda *a, **k: 2)
    monkeypatch.setattr(dist, "all_reduce", fake_all_reduce)
    m = _method(depth=4, dpr=0.0)
    views = [torch.randn(4, 3, 64, 64) for _ in range(2)] + [torch.randn(4, 3, 32, 32) for _ in range(2)]
    random.seed(0)
    m.training_step_impl({"views": views}, 0)
    g = m.s_arena.grad
    base = g.data_ptr()
    covered = sorted(((p - base) // 4, n) for p, n in slices if base <= p < base + g.numel() * 4 and n > 1024)
    assert len(covered) == 2                                  # head bucket + upper-backbone bucket are in flight
    slices.clear()
    m.optimizer_step()
    covered += sorted(((p - base) // 4, n) for p, n in slices if base <= p < base + g.numel() * 4 and n >= 128)
    covered.sort()
    pos = 0
    for off, n in covered:                                     # the buckets tile the arena exactly once
        assert off == pos, (covered, pos)
        pos = off + n
    assert pos == g.numel()


def test_distillation_and_dinov3_schedules_run(stubbed):
    import torchvision

    from lightly_train_b200._methods.distillationv3.distillationv3 import DistillationV3, DistillationV3Args
    from lightly_train_b200._models.dinov3_vit import DinoV3VisionTransformer, DINOv3ViTModelWrapper
    from lightly_train_b200._models.torchvision_resnet import EmbeddingModel, ResNetModelWrapper
    t = DinoV3VisionTransformer(img_size=64, patch_size=16, embed_dim=128, depth=2, num_heads=2, layerscale_init=1e-5,
                                norm_layer="layernormbf16", n_storage_tokens=4, mask_k_bias=True, device="cpu")
    assert {"rope_embed.periods", "blocks.0.attn.qkv.bias_mask", "storage_tokens"} <= set(t.state_dict())
    assert float(t.blocks[0].attn.qkv.bias_mask[128:256].abs().sum()) == 0.0
    student = EmbeddingModel(ResNetModelWrapper(torchvision.models.resnet18()))
    m = DistillationV3(DistillationV3Args(queue_size=16), None, student, 4, 3, teacher_embedding_model=DINOv3ViTModelWrapper(t))
    res = m.training_step_impl({"views": [torch.randn(4, 3, 64, 64)]}, 0)
    res.loss.backward()
    assert m.student_projection_head_global.weight.grad is not None
    assert stubbed.get("b200_rope_apply", 0) == 2 and stubbed.get("b200_kl_rows", 0) == 2
    assert m.teacher_queue.shape == (16, 128)
