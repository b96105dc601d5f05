"""Host logic on CPU: the engine's forward/backward schedule, the model mirror's forward() and the state_dict
schema, with every C-ABI op replaced by its torch restatement (tests/shadow_ops.py), against the golden vectors
of the unmodified reference.  No GPU, no CUDA library involved -- this checks the composition, not the kernels."""
import os

import pytest
import torch

import shadow_ops
import timesformer_oracle as O
from procedurevrl_b200 import functional as PF
from procedurevrl_b200 import ops as real_ops
from procedurevrl_b200.lib.config import get_cfg
from procedurevrl_b200.lib.models import MODEL_REGISTRY
from procedurevrl_b200.lib.models.vit import VisionTransformer

torch.set_num_threads(max(1, os.cpu_count() or 1))


@pytest.fixture
def shadow(monkeypatch):
    for n in shadow_ops.ALL:
        monkeypatch.setattr(real_ops, n, getattr(shadow_ops, n))
    monkeypatch.setattr(VisionTransformer, "_require_cuda", False)


def coin_cfg(gold_dir, depth, T, precision):
    c = get_cfg()
    c.merge_from_list(["DEV.ENABLE", True, "DEV.MATCH_LANG_EMB", True, "DEV.TEST_LANG_EMB",
                       os.path.join(gold_dir, "clip_step_emb_coin.pt"), "TRAIN.DATASET", "howto100m_develop",
                       "TRAIN.LINEAR", True, "MODEL.MODEL_NAME", "vit_base_patch16_224_develop", "MODEL.NUM_CLASSES", 778,
                       "MODEL.ARCH", "vit", "MODEL.PRETRAINED", False, "MODEL.DROP_PATH", 0.0, "TIMESFORMER.DEPTH", depth,
                       "DATA.NUM_FRAMES", T, "DATA.TEST_CROP_SIZE", 224, "B200.PRECISION", precision])
    return c


def pretrain_cfg(gold_dir, depth, precision):
    c = coin_cfg(gold_dir, depth, 8, precision)
    c.merge_from_list(["DEV.ORDER_PRETRAIN_ENABLED", True, "DEV.ORDER_TFM_LAYERS", 4, "TRAIN.LINEAR", False,
                       "TRAIN.LABEL_EMB", os.path.join(gold_dir, "clip_step_emb_coin.pt"), "TRAIN.TOPK", 5,
                       "MODEL.LOSS_FUNC", "kldiv", "MODEL.TEXT_MODEL", "clip_vit_b_16"])
    return c


def check_grads(named, gold, rtol):
    from golden_checks import check_grads as _cg
    _cg(named, gold, rtol)


def finetune_cfg(gold_dir, c, precision):
    """configs/COIN/step_classification.yaml / configs/EK/egocentric_action_classification.yaml without MATCH_LANG_EMB."""
    cfg = coin_cfg(gold_dir, c["depth"], c["T"], precision)
    cfg.merge_from_list(["DEV.MATCH_LANG_EMB", False, "TRAIN.DATASET", c["dataset"], "MODEL.NUM_CLASSES", c["n_classes"]])
    return cfg


def finetune_state(g):
    c = g["cfg"]
    st = O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"])
    st.update(g["extra_state"])
    return st


def test_state_dict_schema(gold_dir):
    """identical names and shapes, so reference checkpoints load with strict=True."""
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(pretrain_cfg(gold_dir, 2, "bf16"))
    st = O.seeded_state(depth=2, frames=8, seed=1, with_order=True)
    own = m.state_dict()
    assert set(own) == set(st), (set(own) ^ set(st))
    for k in own:
        assert own[k].shape == st[k].shape, k
    m.load_state_dict(st, strict=True)
    # fresh init mirrors the reference quirk: temporal_fc zero in every block (vit.py:273-281)
    m2 = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(coin_cfg(gold_dir, 2, 8, "bf16"))
    assert all(b.temporal_fc.weight.abs().sum() == 0 for b in m2.model.blocks)
    assert not m2.model.head.weight.requires_grad          # fine-tune branch freezes the head (vit.py:241)
    names = [n for n, _ in m.named_parameters()]
    assert any("order" in n for n in names) and any(n.startswith("model.head.") for n in names)


@pytest.mark.parametrize("name", ["coin_d2_b2.pt", "coin_d2_t4.pt"])
def test_matchlang_forward_backward(shadow, gold_dir, name):
    g = torch.load(os.path.join(gold_dir, name))
    c = g["cfg"]
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(coin_cfg(gold_dir, c["depth"], c["T"], "bf16x3"))
    m.load_state_dict(O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"]), strict=True)
    for p in m.parameters():
        p.requires_grad_(True)
    x = O.synthetic_clips(c["B"], 3, c["T"], 224, 224, seed=c["clip_seed"])
    m.train()
    logits = m(x)
    torch.testing.assert_close(logits, g["logits"], rtol=1e-3, atol=2e-3)
    assert torch.equal(logits.argmax(1), g["logits"].argmax(1))
    loss = torch.nn.functional.cross_entropy(logits, g["labels"])
    assert abs(loss.item() - g["loss"]) < 1e-3
    loss.backward()
    check_grads({k: p.grad for k, p in m.named_parameters()}, g["grads"], rtol=5e-3)
    m.eval()
    with torch.no_grad():
        probs = m(x)
    torch.testing.assert_close(probs, g["probs"], rtol=5e-3, atol=1e-6)


def test_droppath(shadow, gold_dir):
    g = torch.load(os.path.join(gold_dir, "droppath_d2.pt"))
    c = g["cfg"]
    cfg = coin_cfg(gold_dir, c["depth"], c["T"], "bf16x3")
    cfg.merge_from_list(["MODEL.DROP_PATH", c["rate"]])
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(cfg)
    m.load_state_dict(O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"]), strict=True)
    keep = 1.0 - m.model.dpr[1]
    sc = [torch.floor(keep + t.flatten()) / keep for t in g["rands"]]
    m.model.fixed_drop_scales = [None, {"temporal": sc[0], "spatial": sc[1], "mlp": sc[2]}]
    m.train()
    x = O.synthetic_clips(c["B"], 3, c["T"], 224, 224, seed=c["clip_seed"])
    with torch.no_grad():
        logits = m(x)
    torch.testing.assert_close(logits, g["logits"], rtol=1e-3, atol=2e-3)
    # the module's own draws have the reference's shapes and values in {0, 1/keep}
    m.model.fixed_drop_scales = None
    ds = m.model._drop_scales(c["B"], c["T"], 196, x.device)
    assert ds[0] is None and ds[1]["temporal"].shape == (c["B"] * 196,) and ds[1]["spatial"].shape == (c["B"] * c["T"],)
    assert set((ds[1]["temporal"] * keep).round().tolist()) <= {0.0, 1.0}


@pytest.mark.parametrize("name", ["pretrain_d2_v2.pt"])
def test_pretrain_step(shadow, gold_dir, name):
    """(The depth-12 step on the shipped HowTo100M bank, pretrain_d12_ht100m.pt, is held to its golden by the oracle test
    on CPU and by the product path on the GPU; through the torch restatements of the ops it takes minutes on CPU.)"""
    g = torch.load(os.path.join(gold_dir, name))
    c = g["cfg"]
    Bv = c["Bv"]
    cfg = pretrain_cfg(gold_dir, c["depth"], "bf16x3")
    cfg.merge_from_list(["TRAIN.LABEL_EMB", os.path.join(gold_dir, c.get("bank", "clip_step_emb_coin.pth").replace(".pth", ".pt")),
                         "MODEL.NUM_CLASSES", c.get("n_classes", 778)])
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(cfg)
    m.load_state_dict(O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"], with_order=True), strict=True)
    gen = torch.Generator().manual_seed(c["emb_seed"])
    text_emb = 0.4 * torch.randn(Bv * 9, 512, generator=gen)
    vis_emb = 0.4 * torch.randn(Bv * 9, 512, generator=gen)
    frames = O.synthetic_clips(Bv, 9, 3, c["T"], 224, 224, seed=c["clip_seed"])
    d = g["draws"]
    m.model.order_tfm.fixed_draws = (d["mask_inds"], d["pad_start"], d["noise"])
    m.model.fixed_rand_inds = d["rand_inds"]
    m.train()
    pred, teacher, mse = m([frames, {"clip_text_emb": text_emb, "clip_vis_feat": vis_emb}])
    torch.testing.assert_close(pred, g["pred"], rtol=1e-3, atol=3e-3)
    torch.testing.assert_close(teacher, g["teacher"], rtol=1e-4, atol=5e-4)
    torch.testing.assert_close(mse[0], g["mse0"], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(mse[1], g["mse1"], rtol=1e-3, atol=1e-3)
    loss, l1, l2 = PF.pretrain_loss(pred, teacher, mse, topk=5)
    assert abs(l1.item() - g["loss1"]) < 2e-3 and abs(l2.item() - g["loss2"]) < 2e-3
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert len(grads) == g["n_trainable_with_grad"]
    check_grads(grads, g["grads"], rtol=1e-2)
    # device-side draws of the order transformer: shapes / ranges of the reference's draws (tfm_model.py:145,279-284)
    m.model.order_tfm.fixed_draws = None
    with torch.no_grad():
        den, mask_inds, pair, inter = m.model.order_tfm(torch.randn(Bv * 9, 512), is_pretrain=True)
    assert den.shape == (Bv, 512) and inter.shape == (4 * Bv, 512) and int(mask_inds.max()) < 9


@pytest.mark.parametrize("name", ["finetune_headcls_d2.pt", "finetune_ek_d2.pt"])
def test_finetune_heads(shadow, gold_dir, name):
    """SURVEY 8a row A14 (vit.py:308-322): the path every shipped COIN / EK fine-tuning YAML takes -- frozen `head`, then
    `head_cls`, or the EPIC-Kitchens (verb, noun) tuple, which also skips the eval-mode softmax."""
    g = torch.load(os.path.join(gold_dir, name))
    c = g["cfg"]
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(finetune_cfg(gold_dir, c, "bf16x3"))
    m.load_state_dict(finetune_state(g), strict=True)
    assert sorted(k for k, p in m.named_parameters() if not p.requires_grad) == g["frozen"]
    x = O.synthetic_clips(c["B"], 3, c["T"], 224, 224, seed=c["clip_seed"])
    m.train()
    out = m(x)
    assert isinstance(out, tuple) == g["is_tuple"]
    outs = list(out) if g["is_tuple"] else [out]
    for o, ref in zip(outs, g["outputs"]):
        torch.testing.assert_close(o, ref, rtol=1e-3, atol=2e-3)
        assert torch.equal(o.argmax(1), ref.argmax(1))
    loss = sum(torch.nn.functional.cross_entropy(o, l) for o, l in zip(outs, g["labels"]))
    assert abs(loss.item() - g["loss"]) < 2e-3
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert len(grads) == g["n_with_grad"]
    check_grads(grads, g["grads"], rtol=5e-3)
    m.eval()
    with torch.no_grad():
        ev = m(x)
    for o, ref in zip(list(ev) if g["is_tuple"] else [ev], g["eval_outputs"]):
        torch.testing.assert_close(o, ref, rtol=5e-3, atol=1e-6 if not g["is_tuple"] else 2e-3)


def test_forecast(shadow, gold_dir):
    g = torch.load(os.path.join(gold_dir, "forecast_d2.pt"))
    c = g["cfg"]
    cfg = coin_cfg(gold_dir, c["depth"], c["T"], "bf16x3")
    cfg.merge_from_list(["MODEL.NUM_SEG", c["num_seg"], "MODEL.DROP_E", 0.0])
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(cfg)
    m.load_state_dict(O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"], with_order=True), strict=True)
    x = O.synthetic_clips(c["B"], 3, c["num_seg"] * c["T"], 224, 224, seed=c["clip_seed"])
    m.eval()
    with torch.no_grad():
        probs = m(x)
    torch.testing.assert_close(probs, g["probs"], rtol=5e-3, atol=1e-6)


@pytest.mark.parametrize("name", ["coin_d2_spaceonly.pt", "coin_d2_joint.pt"])
def test_attention_type_variants(shadow, gold_dir, name):
    """TIMESFORMER.ATTENTION_TYPE space_only / joint_space_time (vit.py:124-127, :414-416) through the
    plain-ViT schedule of the engine: logits vs the reference golden, gradients vs the oracle's autograd."""
    g = torch.load(os.path.join(gold_dir, name))
    c = g["cfg"]
    cfg = coin_cfg(gold_dir, c["depth"], c["T"], "bf16x3")
    cfg.merge_from_list(["TIMESFORMER.ATTENTION_TYPE", c["attention_type"]])
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(cfg)
    st = O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"])
    own = m.state_dict()
    assert not any("temporal" in k for k in own)                   # the reference builds no temporal modules here
    m.load_state_dict({k: v for k, v in st.items() if k in own}, strict=True)
    for p in m.parameters():
        p.requires_grad_(True)
    x = O.synthetic_clips(c["B"], 3, c["T"], 224, 224, seed=c["clip_seed"])
    m.train()
    logits = m(x)
    torch.testing.assert_close(logits, g["logits"], rtol=1e-3, atol=2e-3)
    labels = torch.tensor([5, 300])[:c["B"]]
    torch.nn.functional.cross_entropy(logits, labels).backward()
    p = {k: v.clone().requires_grad_(True) for k, v in st.items()}
    e = torch.load(os.path.join(gold_dir, "clip_step_emb_coin.pt"))
    ref = O.match_lang_forward(p, x, e / e.norm(dim=1, keepdim=True), depth=c["depth"], attention_type=c["attention_type"])
    torch.nn.functional.cross_entropy(ref, labels).backward()
    for k, prm in m.named_parameters():
        if prm.grad is None or p[k].grad is None:
            continue
        rn = p[k].grad.norm().item()
        assert abs(prm.grad.norm().item() - rn) <= 5e-3 * rn + 1e-7, k
        torch.testing.assert_close(prm.grad, p[k].grad, rtol=5e-3, atol=5e-3 * rn / prm.numel() ** 0.5 + 1e-7, msg=k)


def test_uint8_frames_equal_normalised_float_frames(shadow, gold_dir):
    """uint8 frames with the loader's normalisation (datasets/utils.py:309-326) fused into the im2col."""
    cfg = coin_cfg(gold_dir, 2, 8, "bf16x3")
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(cfg)
    m.load_state_dict(O.seeded_state(depth=2, frames=8, seed=3), strict=True)
    m.eval()
    u8 = torch.randint(0, 256, (2, 3, 8, 224, 224), generator=torch.Generator().manual_seed(1), dtype=torch.uint8)
    xf = (u8.float() / 255.0 - 0.45) / 0.225
    with torch.no_grad():
        torch.testing.assert_close(m(u8), m(xf), rtol=1e-4, atol=1e-6)
