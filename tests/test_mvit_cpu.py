"""MViTv2 path (BASELINE config 5), host logic on CPU: the model mirror (procedurevrl_b200/lib/models/mvit.py) and the
autograd Functions of mvit_functional / tc_functional with every C-ABI op replaced by its torch restatement
(tests/shadow_ops.py), against the golden vectors of the UNMODIFIED reference (tests/golden/mvit_*.pt): state_dict schema
incl. the shipped 16 x 224 model, logits, per-block cls rows, loss, every parameter gradient.  No GPU involved -- this
checks the composition (geometry, pooling strides, relative-position tables, skip paths, gradient routing), the kernels
themselves are checked in tests/test_mvit_gpu.py."""
import json
import os

import pytest
import torch

import mvit_oracle as MO
import shadow_ops
from procedurevrl_b200 import ops as real_ops
from procedurevrl_b200.lib.config import get_cfg
from procedurevrl_b200.lib.models import MODEL_REGISTRY
from procedurevrl_b200.lib.models import mvit as MV
from procedurevrl_b200 import tc_functional as TC

torch.set_num_threads(max(1, os.cpu_count() or 1))
CASES = ["d4_t4_c64", "d3_t8_c96"]


@pytest.fixture
def shadow(monkeypatch):
    for n in shadow_ops.ALL:
        monkeypatch.setattr(real_ops, n, getattr(shadow_ops, n))
    monkeypatch.setattr(MV.MViT_encoder, "_require_cuda", False)
    TC._WCACHE.clear()


def mvit_cfg(gold_dir, mvit_keys, frames, crop, precision, droppath=0.0):
    c = get_cfg()
    c.merge_from_list(["DEV.ENABLE", True, "DEV.MATCH_LANG_EMB", True, "DEV.TEST_LANG_EMB",
                       os.path.join(gold_dir, "clip_step_emb_coin.pt"), "TRAIN.LABEL_EMB", "", "MODEL.TEXT_MODEL", "",
                       "MODEL.MODEL_NAME", "MViT", "MODEL.ARCH", "mvit", "MODEL.NUM_CLASSES", 778, "MODEL.PRETRAINED", False,
                       "DATA.NUM_FRAMES", frames, "DATA.TRAIN_CROP_SIZE", crop, "DATA.TEST_CROP_SIZE", crop,
                       "DATA.INPUT_CHANNEL_NUM", [3], "B200.PRECISION", precision])
    for k, v in mvit_keys.items():
        c.MVIT[k] = v
    c.MVIT.DROPPATH_RATE = droppath
    return c


def build(gold_dir, g, precision):
    c = g["cfg"]
    m = MODEL_REGISTRY.get("MViT")(mvit_cfg(gold_dir, c["mvit"], c["frames"], c["crop"], precision))
    m.load_state_dict(MO.seeded_state(g["shapes"], c["seed"]), strict=True)      # identical schema or this raises
    for p in m.parameters():
        p.requires_grad_(True)
    return m.train()


def test_full_size_schema_and_geometry(gold_dir):
    """The shipped MViTv2-S 16 x 224 (procedurevrl_mvitv2_adamw.yaml): parameter names / shapes equal the reference's."""
    with open(os.path.join(gold_dir, "mvit_full_geometry.json")) as f:
        g = json.load(f)
    enc = MV.MViT_encoder(mvit_cfg(gold_dir, g["mvit"], g["frames"], g["crop"], "bf16"))
    own = {MO.PRE + k: list(v.shape) for k, v in enc.state_dict().items()}
    assert own == g["shapes"]
    assert sum(v.numel() for v in enc.state_dict().values()) == g["n_params"]
    assert enc.patch_dims == g["patch_dims"] and enc.geo["out_dim"] == g["out_dim"]
    with pytest.raises(NotImplementedError):
        MV.mvit_geometry(dict(g["mvit"], POOL_FIRST=True), 16, 224)


def test_no_cpu_fallback(gold_dir):
    g = torch.load(os.path.join(gold_dir, "mvit_d4_t4_c64.pt"))
    m = build(gold_dir, g, "bf16")
    with pytest.raises((RuntimeError, AssertionError)):
        m(MO.synthetic_clips(1, 4, 64, 3))


@pytest.mark.parametrize("case", CASES)
def test_forward_backward_match_reference(shadow, gold_dir, case):
    g = torch.load(os.path.join(gold_dir, f"mvit_{case}.pt"))
    c = g["cfg"]
    m = build(gold_dir, g, "bf16x3")
    x = MO.synthetic_clips(c["B"], c["frames"], c["crop"], c["seed"] + 1)
    taps = []
    enc = m.model.video_encoder
    orig = enc.forward
    enc.forward = lambda clips: orig(clips, taps=taps)
    logits = m(x)
    torch.testing.assert_close(logits.detach(), g["logits"], rtol=1e-3, atol=5e-3)      # the north star's logits tolerance
    assert torch.equal(logits.argmax(1), g["logits"].argmax(1))
    assert len(taps) == len(g["taps"])
    for t, ref in zip(taps, g["taps"]):
        assert tuple(t.shape) == tuple(ref["shape"])
        torch.testing.assert_close(t[:, 0].detach(), ref["cls"], rtol=1e-3, atol=1e-4)
        assert abs(t.norm().item() - ref["norm"]) <= 1e-3 * ref["norm"]
    loss = torch.nn.functional.cross_entropy(logits, g["labels"])
    assert abs(loss.item() - g["loss"]) <= 1e-3 * abs(g["loss"])
    loss.backward()
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"]) and len(got) == g["n_grads"]
    for k, ref in g["grads"].items():
        assert abs(got[k].norm().item() - ref["norm"]) <= 1e-2 * ref["norm"] + 1e-7, k
        torch.testing.assert_close(got[k].flatten()[:32], ref["head"], rtol=1e-2, atol=1e-6 + 1e-3 * ref["norm"],
                                   msg=lambda s, k=k: f"{k}: {s}")


def pretrain_model(gold_dir, g, precision):
    """The config-5 pre-training model (TRAIN.LABEL_EMB set, order transformer, text side pre-extracted) with the golden's state."""
    import timesformer_oracle as TO
    c = g["cfg"]
    cfg = mvit_cfg(gold_dir, c["mvit"], c["frames"], c["crop"], precision)
    cfg.merge_from_list(["DEV.ORDER_PRETRAIN_ENABLED", True, "DEV.ORDER_TFM_LAYERS", 4, "TRAIN.LABEL_EMB",
                         os.path.join(gold_dir, "clip_step_emb_coin.pt"), "TRAIN.TOPK", c["topk"], "MODEL.LOSS_FUNC", "kldiv",
                         "MODEL.TEXT_MODEL", "clip_vit_b_16"])
    m = MODEL_REGISTRY.get("MViT")(cfg)
    shapes = {k: v for k, v in g["shapes"].items() if k.startswith(MO.PRE) or k.startswith("model.head.")}
    state = MO.seeded_state(shapes, c["seed"])
    state.update({k: v for k, v in TO.seeded_state(depth=1, frames=8, seed=c["seed"] + 1, with_order=True).items()
                  if k.startswith("model.order_tfm.")})
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v) for k, v in g["shapes"].items()}
    m.load_state_dict(state, strict=True)
    gen = torch.Generator().manual_seed(c["seed"] + 2)
    text_emb = 0.4 * torch.randn(c["Bv"] * 9, 512, generator=gen)
    vis_emb = 0.4 * torch.randn(c["Bv"] * 9, 512, generator=gen)
    x = MO.synthetic_clips(c["Bv"] * 9, c["frames"], c["crop"], c["seed"] + 3).reshape(c["Bv"], 9, 3, c["frames"], c["crop"], c["crop"])
    d = g["draws"]
    m.model.order_tfm.fixed_draws = (d["mask_inds"], d["pad_start"], d["noise"])
    m.model.fixed_rand_inds = d["rand_inds"]
    return m.train(), x, {"clip_text_emb": text_emb, "clip_vis_feat": vis_emb}


def check_pretrain_step(m, x, meta, g, dev="cpu"):
    from procedurevrl_b200 import functional as PF
    pred, teacher, mse = m([x.to(dev), {k: v.to(dev) for k, v in meta.items()}])
    torch.testing.assert_close(pred.detach().cpu(), g["pred"], rtol=1e-3, atol=5e-3)
    torch.testing.assert_close(teacher.cpu(), g["teacher"], rtol=1e-4, atol=5e-4)
    torch.testing.assert_close(mse[0].detach().cpu(), g["mse0"], rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(mse[1].detach().cpu(), g["mse1"], rtol=1e-3, atol=1e-3)
    loss, l1, l2 = PF.pretrain_loss(pred, teacher, mse, topk=g["cfg"]["topk"])
    assert abs(l1.item() - g["loss1"]) < 2e-3 and abs(l2.item() - g["loss2"]) < 2e-3
    loss.backward()
    got = {k: p.grad for k, p in m.named_parameters() if p.grad is not None}
    assert set(got) == set(g["grads"]) and len(got) == g["n_grads"]
    for k, ref in g["grads"].items():
        assert abs(got[k].norm().item() - ref["norm"]) <= 1e-2 * ref["norm"] + 1e-7, k
        torch.testing.assert_close(got[k].flatten()[:32].cpu(), ref["head"], rtol=1e-2, atol=1e-6 + 1e-3 * ref["norm"],
                                   msg=lambda s, k=k: f"{k}: {s}")


def test_pretrain_step_matches_reference(shadow, gold_dir):
    """BASELINE config 5's step as tools/train_net.py:146-162 runs it: model([frames, meta]) -> (pred, teacher, mse) ->
    KL(top-k) + MSE -> backward, against the unmodified reference (its random draws replayed)."""
    g = torch.load(os.path.join(gold_dir, "mvit_pretrain_d4_t4_c64.pt"))
    m, x, meta = pretrain_model(gold_dir, g, "bf16x3")
    check_pretrain_step(m, x, meta, g)


def test_throughput_mode_and_droppath(shadow, gold_dir):
    """bf16 activations / operands: logits stay within the bf16 budget of the reference's, DropPath rows scale whole clips."""
    g = torch.load(os.path.join(gold_dir, "mvit_d4_t4_c64.pt"))
    c = g["cfg"]
    m = build(gold_dir, g, "bf16")
    x = MO.synthetic_clips(c["B"], c["frames"], c["crop"], c["seed"] + 1)
    logits = m(x)
    assert (logits.detach() - g["logits"]).abs().max().item() < 1.0          # cosine / 0.02: 0.02 in cosine units
    logits.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.model.video_encoder.parameters())
    enc = m.model.video_encoder
    enc.fixed_drop_scales = [None] + [torch.tensor([0.0, 2.0])] * (len(enc.blocks) - 1)
    out = m(x)
    assert torch.isfinite(out).all() and not torch.allclose(out, logits)
    m.eval()
    enc.fixed_drop_scales = None
    probs = m(x)
    torch.testing.assert_close(probs.sum(1), torch.ones(c["B"]), rtol=1e-4, atol=1e-4)  # eval returns softmax (mvit.py:183-184)
