"""CPU: the drop-in boundary against the reference's OWN method class (through oracle/ref_full.py) without touching a GPU:
reference constructor (pre-built embedding model), state_dict names / shapes both ways, optimizer param groups, and the
per-step lr / weight-decay / freeze values the hooks produce (dinov2.py:550-639)."""
import pytest
import torch

from oracle import ref_full

pytestmark = pytest.mark.skipif(not ref_full.available(), reason="reference source not present")


def _pair(vit_kw, method_kw, max_steps=20):
    from lightly_train_b200._methods.dinov2.dinov2 import DINOv2, DINOv2AdamWViTArgs, DINOv2Args
    torch.manual_seed(0)
    ref, ropt, rsched = ref_full.build_dinov2(dict(vit_kw, block_chunks=0), dict(method_kw), global_batch_size=64, max_steps=max_steps)
    # a FRESH reference backbone (with its stochastic depth) is what a caller hands to the constructor
    ref_full.install()
    from lightly_train._models.dinov2_vit.dinov2_vit import DINOv2ViTModelWrapper  # type: ignore
    from lightly_train._models.dinov2_vit.dinov2_vit_src.models.vision_transformer import DinoVisionTransformer  # type: ignore
    from lightly_train._models.embedding_model import EmbeddingModel  # type: ignore
    emb = EmbeddingModel(wrapped_model=DINOv2ViTModelWrapper(DinoVisionTransformer(**dict(vit_kw, block_chunks=0))))
    mine = DINOv2(DINOv2Args(**method_kw), DINOv2AdamWViTArgs(), emb, 64, 3, max_steps=max_steps, device="cpu")
    return ref, ropt, rsched, mine, emb


@pytest.mark.parametrize("vit_kw,method_kw", [
    (dict(img_size=224, patch_size=16, embed_dim=128, depth=4, num_heads=2, init_values=1e-5, drop_path_rate=0.3), dict(output_dim=512, hidden_dim=256)),
    (dict(img_size=224, patch_size=14, embed_dim=128, depth=2, num_heads=2, init_values=1e-5, drop_path_rate=0.2, drop_path_uniform=True,
          ffn_layer="swiglufused", num_register_tokens=4, interpolate_antialias=True, interpolate_offset=0.0),
     dict(output_dim=512, hidden_dim=256, ibot_separate_head=True, center_method="sinkhorn_knopp")),
])
def test_reference_constructor_and_state_dict(vit_kw, method_kw):
    ref, _, _, mine, emb = _pair(vit_kw, method_kw)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())                      # same names, same ORDER
    assert all(a[k].shape == b[k].shape for k in a)
    # architecture read back from the module, incl. the stochastic-depth schedule of the student (teacher: none)
    s = mine.s_vit
    want = [float(x) for x in ([vit_kw["drop_path_rate"]] * vit_kw["depth"] if vit_kw.get("drop_path_uniform")
                               else torch.linspace(0, vit_kw["drop_path_rate"], vit_kw["depth"]).tolist())]
    assert s.dpr == pytest.approx(want) and mine.t_vit.dpr == [0.0] * vit_kw["depth"]
    assert s.swiglu == ("ffn_layer" in vit_kw) and s.num_register_tokens == vit_kw.get("num_register_tokens", 0)
    # the passed backbone's weights initialise BOTH sides (reference: student = deepcopy(teacher), dinov2.py:197-198)
    w = emb.wrapped_model.get_model().state_dict()["blocks.1.attn.qkv.weight"]
    assert torch.equal(mine.state_dict()["teacher_embedding_model.wrapped_model._model.blocks.1.attn.qkv.weight"], w)
    assert torch.equal(mine.state_dict()["student_embedding_model.wrapped_model._model.blocks.1.attn.qkv.weight"], w)
    # checkpoints flow both ways
    assert not mine.load_state_dict(a, strict=True).missing_keys
    res = ref.load_state_dict(b, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_param_groups_and_hook_schedules_match_reference():
    """Same group names, lr multipliers and weight-decay switches as get_optimizer_with_decay + get_fused_param_groups, and the
    same per-step lr (warm-up + cosine), weight decay (cosine) and freeze decisions as the reference's hooks, step by step."""
    vit_kw = dict(img_size=224, patch_size=16, embed_dim=128, depth=4, num_heads=2, init_values=1e-5, drop_path_rate=0.0)
    mk = dict(output_dim=512, hidden_dim=256, warmup_steps=3, student_freeze_last_layer_steps=2, student_freeze_backbone_steps=1)
    ref, ropt, rsched, mine, _ = _pair(vit_kw, mk, max_steps=8)
    (opt,), (sch,) = mine.configure_optimizers()
    rg = {g["name"]: g for g in ropt.param_groups}
    mg = {g["name"]: g for g in opt.param_groups}
    assert list(rg) == list(mg)
    assert all(sum(p.numel() for p in rg[n]["params"]) == sum(p.numel() for p in mg[n]["params"]) for n in rg)
    for step in range(8):
        ref.on_before_optimizer_step(ropt)
        mine.on_before_optimizer_step(opt)
        for n in rg:
            assert mg[n]["lr"] == pytest.approx(rg[n]["lr"], rel=1e-6, abs=1e-12), (step, n)
            assert mg[n]["weight_decay"] == pytest.approx(rg[n]["weight_decay"], rel=1e-6), (step, n)
        assert opt.freeze_backbone == (step < 1) and opt.freeze_last_layer == (step < 2)
        # (no optimizer.step(): the reference's would need gradients; the schedules only depend on the counters)
        ropt.step(); rsched.step(); ref.trainer.global_step += 1
        sch["scheduler"].step(); mine.trainer.global_step += 1
