"""Pins the CPU oracle (oracle/timesformer_oracle.py) against golden vectors produced by the
unmodified reference (oracle/make_golden.py).  CPU only; the whole file runs in ~1-2 min."""
import os

import pytest
import torch
import torch.nn.functional as F

import timesformer_oracle as O

torch.set_num_threads(max(1, os.cpu_count() or 1))
RTOL, ATOL = 1e-4, 2e-4          # fp32 restatement vs fp32 reference on the same CPU: reassociation only


def _load(gold_dir, name):
    return torch.load(os.path.join(gold_dir, name))


def _check_grads(grads, gold_grads, rtol=2e-3):
    from golden_checks import check_grads
    check_grads(grads, gold_grads, rtol)


def _bank(gold_dir, cfg):
    """The text bank a pretrain golden was made with (cfg['bank'], default the COIN fixture), rows L2-normalised."""
    name = cfg.get("bank", "clip_step_emb_coin.pth").replace(".pth", ".pt")
    e = torch.load(os.path.join(gold_dir, name))
    return e / e.norm(dim=1, keepdim=True)


@pytest.mark.parametrize("name", ["coin_d2_b2.pt", "coin_d2_t4.pt", "coin_d12_b4.pt"])
def test_matchlang_logits_taps_grads(gold_dir, coin_label_emb, name):
    g = _load(gold_dir, name)
    c = g["cfg"]
    p = O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"])
    for v in p.values():
        v.requires_grad_(True)
    x = O.synthetic_clips(c["B"], 3, c["T"], 224, 224, seed=c["clip_seed"])
    taps = {}
    logits = O.match_lang_forward(p, x, coin_label_emb, depth=c["depth"], taps=taps)
    torch.testing.assert_close(logits, g["logits"], rtol=RTOL, atol=ATOL)
    assert torch.equal(logits.argmax(1), g["logits"].argmax(1))
    for k, v in g["taps"].items():
        torch.testing.assert_close(taps[k][:, :6, :16], v, rtol=RTOL, atol=ATOL)
    loss = F.cross_entropy(logits, g["labels"])
    assert abs(loss.item() - g["loss"]) < 1e-4
    loss.backward()
    _check_grads({k: v.grad for k, v in p.items()}, g["grads"])
    with torch.no_grad():
        probs = O.match_lang_forward(p, x, coin_label_emb, depth=c["depth"], training=False)
    torch.testing.assert_close(probs, g["probs"], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("name", ["coin_d2_joint.pt", "coin_d2_spaceonly.pt"])
def test_attention_type_variants(gold_dir, coin_label_emb, name):
    g = _load(gold_dir, name)
    c = g["cfg"]
    p = O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"])
    x = O.synthetic_clips(c["B"], 3, c["T"], 224, 224, seed=c["clip_seed"])
    with torch.no_grad():
        logits = O.match_lang_forward(p, x, coin_label_emb, depth=c["depth"], attention_type=c["attention_type"])
    torch.testing.assert_close(logits, g["logits"], rtol=RTOL, atol=ATOL)


def test_droppath_semantics(gold_dir, coin_label_emb):
    g = _load(gold_dir, "droppath_d2.pt")
    c = g["cfg"]
    p = O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"])
    x = O.synthetic_clips(c["B"], 3, c["T"], 224, 224, seed=c["clip_seed"])
    rates = torch.linspace(0, c["rate"], c["depth"]).tolist()            # vit.py:220
    keep = 1.0 - rates[1]
    r = [t.flatten() for t in g["rands"]]
    sc = [torch.floor(keep + t) / keep for t in r]                       # vit_utils.py:150-154
    scales = [None, {"temporal": sc[0], "spatial": sc[1], "mlp": sc[2]}]
    with torch.no_grad():
        logits = O.match_lang_forward(p, x, coin_label_emb, depth=c["depth"], drop_scales=scales)
    torch.testing.assert_close(logits, g["logits"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", ["pretrain_d2_v2.pt", "pretrain_d12_v1.pt", "pretrain_d12_ht100m.pt"])
def test_pretrain_step(gold_dir, name):
    """pretrain_d12_ht100m.pt: the shipped HowTo100M verb-phrase bank (K = 9871), the one procedurevrl_adamw.yaml trains on."""
    g = _load(gold_dir, name)
    c = g["cfg"]
    coin_label_emb = _bank(gold_dir, c)
    Bv = c["Bv"]
    p = O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"], with_order=True)
    for v in p.values():
        v.requires_grad_(True)
    gen = torch.Generator().manual_seed(c["emb_seed"])
    text_emb = 0.4 * torch.randn(Bv * 9, 512, generator=gen)
    vis_emb = 0.4 * torch.randn(Bv * 9, 512, generator=gen)
    frames = O.synthetic_clips(Bv, 9, 3, c["T"], 224, 224, seed=c["clip_seed"])
    d = g["draws"]
    draws = O.OrderDraws(d["mask_inds"], d["pad_start"], d["noise"], d["rand_inds"])
    pred, teacher, mse = O.pretrain_forward(p, frames, text_emb, vis_emb, coin_label_emb, draws, depth=c["depth"])
    torch.testing.assert_close(pred, g["pred"], rtol=RTOL, atol=5e-4)
    torch.testing.assert_close(teacher, g["teacher"], rtol=RTOL, atol=5e-4)
    torch.testing.assert_close(mse[0], g["mse0"], rtol=RTOL, atol=1e-5)
    torch.testing.assert_close(mse[1], g["mse1"], rtol=RTOL, atol=1e-4)
    torch.testing.assert_close(O.topk_teacher(teacher), g["teacher_topk"], rtol=1e-3, atol=1e-6)
    loss, l1, l2 = O.pretrain_loss(pred, teacher, mse)
    assert abs(l1.item() - g["loss1"]) < 2e-4 and abs(l2.item() - g["loss2"]) < 2e-4
    loss.backward()
    grads = {k: v.grad for k, v in p.items() if v.grad is not None}
    assert len(grads) == g["n_trainable_with_grad"]
    _check_grads(grads, g["grads"])


def test_forecast(gold_dir, coin_label_emb):
    g = _load(gold_dir, "forecast_d2.pt")
    c = g["cfg"]
    p = O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"], with_order=True)
    x = O.synthetic_clips(c["B"], 3, c["num_seg"] * c["T"], 224, 224, seed=c["clip_seed"])
    with torch.no_grad():
        S, T = c["num_seg"], c["T"]
        xs = x.reshape(c["B"], 3, S, T, 224, 224).permute(0, 2, 1, 3, 4, 5).reshape(c["B"] * S, 3, T, 224, 224)  # vit.py:293
        feat = O.forward_features(p, xs, c["depth"])
        emb = O.video_embedding(p, feat)
        f = O.l2_normalize(O.order_tfm_forecast(p, emb, S))                # vit.py:304-306
        probs = O.similarity_logits(f, coin_label_emb).softmax(1)
    torch.testing.assert_close(probs, g["probs"], rtol=2e-3, atol=1e-6)


@pytest.mark.parametrize("name", ["finetune_headcls_d2.pt", "finetune_ek_d2.pt"])
def test_finetune_heads(gold_dir, name):
    """SURVEY 8a row A14 (vit.py:308-322): `head_cls` without DEV.MATCH_LANG_EMB, and the EPIC-Kitchens (verb, noun) tuple."""
    g = _load(gold_dir, name)
    c = g["cfg"]
    p = O.seeded_state(depth=c["depth"], frames=c["T"], seed=c["state_seed"])
    p.update({k: v.clone() for k, v in g["extra_state"].items()})
    for v in p.values():
        v.requires_grad_(True)
    x = O.synthetic_clips(c["B"], 3, c["T"], 224, 224, seed=c["clip_seed"])
    fwd = O.finetune_ek_forward if g["is_tuple"] else O.finetune_cls_forward
    out = fwd(p, x, depth=c["depth"])
    outs = list(out) if g["is_tuple"] else [out]
    assert len(outs) == len(g["outputs"])
    for o, ref in zip(outs, g["outputs"]):
        torch.testing.assert_close(o, ref, rtol=RTOL, atol=ATOL)
        assert torch.equal(o.argmax(1), ref.argmax(1))
    loss = sum(F.cross_entropy(o, l) for o, l in zip(outs, g["labels"]))
    assert abs(loss.item() - g["loss"]) < 2e-4
    loss.backward()
    _check_grads({k: v.grad for k, v in p.items()}, g["grads"])
    with torch.no_grad():
        ev = fwd(p, x, depth=c["depth"], training=False)
    for o, ref in zip(list(ev) if g["is_tuple"] else [ev], g["eval_outputs"]):
        torch.testing.assert_close(o, ref, rtol=1e-3, atol=1e-6)
