"""Golden-vector generator: runs the UNMODIFIED reference (imported from /root/reference through
`oracle/ref_shims.py`) on seeded synthetic inputs and writes small fixtures to `tests/golden/`.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not exist on the GPU
box):

    python oracle/make_golden.py            # regenerates every tests/golden/*.pt

Weights come from `timesformer_oracle.seeded_state` (same generator the tests use, so only the
seed travels, never 480 MB of weights) and are loaded into the reference module with
`load_state_dict(strict=True)`, which also pins the state_dict schema.
Random draws the reference makes internally (mask positions, pad starts, diffusion noise,
randperm, DropPath uniforms) are recorded by wrapping `torch.randint/randn_like/randperm/rand`
for the duration of the call and stored next to the outputs.
"""
import contextlib
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import timesformer_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
COIN_EMB = os.path.join(GOLD, "clip_step_emb_coin.pt")


@contextlib.contextmanager
def record_rng(log):
    names = ["randint", "randn_like", "randperm", "rand"]
    orig = {n: getattr(torch, n) for n in names}

    def wrap(n):
        def f(*a, **k):
            out = orig[n](*a, **k)
            log.append((n, out.detach().clone().cpu()))
            return out
        return f
    for n in names:
        setattr(torch, n, wrap(n))
    try:
        yield
    finally:
        for n in names:
            setattr(torch, n, orig[n])


def build_ref(yaml_rel, overrides, state):
    vit, build, _ = ref_shims.load_reference()
    cwd = os.getcwd()
    os.chdir(ref_shims.REFERENCE_ROOT)      # the YAMLs use ./data/... relative paths
    try:
        cfg = ref_shims.reference_cfg(yaml_rel, overrides)
        m = build.build_model(cfg)
    finally:
        os.chdir(cwd)
    sd = {k: v for k, v in state.items()}
    own = m.state_dict()
    missing = [k for k in own if k not in sd and not k.startswith("model.text_model.")]
    extra = [k for k in sd if k not in own]
    assert not missing and not extra, (missing[:5], extra[:5])
    for k in own:                                   # text_model buffers (clip stub table) stay
        if k.startswith("model.text_model."):
            sd[k] = own[k]
    m.load_state_dict(sd, strict=True)
    # GPU semantics of check_device_norm (vit.py:435-440): rows L2-normalised on first device move.
    if isinstance(m.model.label_emb, torch.Tensor):
        m.model.label_emb = m.model.label_emb / m.model.label_emb.norm(dim=1, keepdim=True)
    return m, cfg


def block_taps(m):
    taps = {}
    hooks = []
    for i, blk in enumerate(m.model.blocks):
        hooks.append(blk.register_forward_hook(
            lambda mod, inp, out, i=i: taps.__setitem__(f"block{i}", out.detach()[:, :6, :16].clone())))
    return taps, hooks


def grad_windows(g, n_win=4, width=64):
    """`n_win` windows of `width` consecutive elements at evenly spaced offsets past the head (the last one ends at the
    tensor's end), so that an error anywhere in a large dW -- a permuted tile, a dropped split-K slice -- is seen."""
    f = g.flatten()
    n = f.numel()
    if n <= width:
        return torch.zeros(0, width), []
    offs = sorted({min(n - width, (i + 1) * (n - width) // n_win) for i in range(n_win)})
    return torch.stack([f[o:o + width] for o in offs]).clone(), offs


def grad_summary(named_grads):
    out = {}
    for k, g in named_grads.items():
        win, offs = grad_windows(g)
        out[k] = {"norm": g.norm().item(), "head": g.flatten()[:64].clone(), "sum": g.double().sum().item(),
                  "abs_sum": g.double().abs().sum().item(), "win": win, "win_offsets": offs}
    return out


GRAD_KEYS = ["model.cls_token", "model.pos_embed", "model.time_embed", "model.patch_embed.proj.weight",
             "model.patch_embed.proj.bias", "model.norm.weight", "model.norm.bias"]
BLOCK_GRAD = ["norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
              "attn.proj.bias", "temporal_norm1.weight", "temporal_norm1.bias", "temporal_attn.qkv.weight",
              "temporal_attn.qkv.bias", "temporal_attn.proj.weight", "temporal_attn.proj.bias",
              "temporal_fc.weight", "temporal_fc.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight",
              "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"]


def gold_coin_matchlang(depth, B, T, name, attention_type="divided_space_time", with_grads=True):
    """COIN-shape parity config: step_classification.yaml + DEV.MATCH_LANG_EMB."""
    ov = ["DEV.MATCH_LANG_EMB", True, "MODEL.DROP_PATH", 0.0, "TIMESFORMER.DEPTH", depth,
          "DATA.NUM_FRAMES", T, "TIMESFORMER.ATTENTION_TYPE", attention_type]
    st = O.seeded_state(depth=depth, frames=T, seed=depth * 100 + T)
    if attention_type != "divided_space_time":
        st = {k: v for k, v in st.items() if "temporal" not in k}
        if attention_type == "space_only":
            st.pop("model.time_embed")
    m, cfg = build_ref("configs/COIN/step_classification.yaml", ov, st)
    x = O.synthetic_clips(B, 3, T, 224, 224, seed=7)
    m.train()
    for p_ in m.parameters():
        p_.requires_grad_(True)                      # head is frozen by the reference (vit.py:241); we want its grad too
    taps, hooks = block_taps(m)
    logits = m(x)
    for h in hooks:
        h.remove()
    out = {"cfg": {"depth": depth, "B": B, "T": T, "attention_type": attention_type,
                   "state_seed": depth * 100 + T, "clip_seed": 7},
           "logits": logits.detach().clone(), "taps": taps}
    if with_grads:
        labels = torch.arange(B) * 37 % logits.shape[1]
        loss = F.cross_entropy(logits, labels)
        loss.backward()
        grads = {k: p_.grad for k, p_ in m.named_parameters() if p_.grad is not None}
        keys = [k for k in GRAD_KEYS if k in grads] + ["model.head.weight", "model.head.bias"]
        for i in sorted({0, depth - 1}):
            keys += [f"model.blocks.{i}.{s}" for s in BLOCK_GRAD if f"model.blocks.{i}.{s}" in grads]
        out["labels"] = labels
        out["loss"] = loss.item()
        out["grads"] = grad_summary({k: grads[k] for k in keys})
    m.eval()
    with torch.no_grad():
        out["probs"] = m(x).clone()
    torch.save(out, os.path.join(GOLD, name))
    print(name, "logits", tuple(logits.shape), "loss", out.get("loss"))


def gold_droppath(name):
    """DropPath semantics (vit_utils.py:140-155; per-row-of-view masks, rates linspace vit.py:220)."""
    depth, B, T, rate = 2, 2, 8, 0.5
    st = O.seeded_state(depth=depth, frames=T, seed=31)
    m, cfg = build_ref("configs/COIN/step_classification.yaml",
                       ["DEV.MATCH_LANG_EMB", True, "MODEL.DROP_PATH", rate, "TIMESFORMER.DEPTH", depth], st)
    x = O.synthetic_clips(B, 3, T, 224, 224, seed=8)
    m.train()
    torch.manual_seed(5)
    log = []
    with record_rng(log), torch.no_grad():
        logits = m(x)
    rands = [t for n, t in log if n == "rand"]
    torch.save({"cfg": {"depth": depth, "B": B, "T": T, "rate": rate, "state_seed": 31, "clip_seed": 8},
                "rands": rands, "logits": logits.clone()}, os.path.join(GOLD, name))
    print(name, "n_rand", len(rands), [tuple(r.shape) for r in rands])


def gold_pretrain(depth, Bv, name, bank="./data/clip_step_emb_coin.pth", n_classes=778):
    """HowTo100M stage-2 pretrain step (procedurevrl_adamw.yaml) with the COIN step bank (small fixture) or the shipped
    HowTo100M verb-phrase bank (K = 9871, the one the YAML names) as LABEL_EMB, and the CLIP text tower replaced by
    pre-extracted text embeddings."""
    T, max_len = 8, 9
    st = O.seeded_state(depth=depth, frames=T, seed=900 + depth, with_order=True)
    g = torch.Generator().manual_seed(77)
    text_emb = 0.4 * torch.randn(Bv * max_len, 512, generator=g)
    vis_emb = 0.4 * torch.randn(Bv * max_len, 512, generator=g)
    ov = ["MODEL.DROP_PATH", 0.0, "TIMESFORMER.DEPTH", depth, "TRAIN.LABEL_EMB", bank,
          "MODEL.NUM_CLASSES", n_classes, "TRAIN.TEXT", "preextracted"]
    m, cfg = build_ref("configs/HowTo100M/procedurevrl_adamw.yaml", ov, st)
    # pre-extracted text embeddings stand in for the frozen tower's output (north star)
    m.model.text_model.encode_text = lambda ids: text_emb
    frames = O.synthetic_clips(Bv, max_len, 3, T, 224, 224, seed=9)
    meta = {"clip_text_ids": torch.zeros(Bv * max_len, 77, dtype=torch.long), "clip_vis_feat": vis_emb}
    m.train()
    m.model.text_model.eval()
    torch.manual_seed(11)
    log = []
    with record_rng(log):
        pred, teacher, mse = m([frames, meta])
    ints = [t for n, t in log if n == "randint"]
    mask_inds = ints[0]
    pad_start = torch.full((Bv,), max_len, dtype=torch.long)
    j = 1
    for i in range(Bv):
        if int(mask_inds[i]) + 1 != max_len:
            pad_start[i] = int(ints[j])
            j += 1
    noise = torch.stack([t for n, t in log if n == "randn_like"])
    rand_inds = [t for n, t in log if n == "randperm"][0]
    # loss exactly as tools/train_net.py:152-162
    with torch.no_grad():
        tp = F.softmax(teacher, 1)
        tp = (tp.unsqueeze(1) * (tp.unsqueeze(1) == tp.topk(k=cfg.TRAIN.TOPK, dim=1)[0].unsqueeze(2)).float()).sum(1)
        tp = tp / tp.sum(1, keepdim=True)
    loss1 = torch.nn.KLDivLoss(reduction="batchmean")(F.log_softmax(pred, dim=1), tp)
    loss2 = torch.nn.MSELoss(reduction="mean")(mse[0], mse[1])
    loss = loss1 + loss2
    loss.backward()
    grads = {k: p_.grad for k, p_ in m.named_parameters() if p_.grad is not None}
    keys = GRAD_KEYS + ["model.head.weight", "model.head.bias"]
    for i in sorted({0, depth - 1}):
        keys += [f"model.blocks.{i}.{s}" for s in BLOCK_GRAD]
    keys += [k for k in grads if k.startswith("model.order_tfm.")]
    torch.save({"cfg": {"depth": depth, "Bv": Bv, "T": T, "state_seed": 900 + depth, "clip_seed": 9, "emb_seed": 77,
                        "bank": os.path.basename(bank), "n_classes": n_classes},
                "draws": {"mask_inds": mask_inds, "pad_start": pad_start, "noise": noise, "rand_inds": rand_inds},
                "pred": pred.detach().clone(), "teacher": teacher.detach().clone(),
                "mse0": mse[0].detach().clone(), "mse1": mse[1].detach().clone(),
                "teacher_topk": tp.clone(),
                "loss": loss.item(), "loss1": loss1.item(), "loss2": loss2.item(),
                "n_trainable_with_grad": len(grads),
                "grads": grad_summary({k: grads[k] for k in keys})}, os.path.join(GOLD, name))
    print(name, "pred", tuple(pred.shape), "loss", loss.item(), loss1.item(), loss2.item(),
          "mask", mask_inds.tolist(), "pad", pad_start.tolist())


def gold_forecast(name):
    """Zero-shot step forecasting (step_forecasting.yaml + MATCH_LANG_EMB, MODEL.NUM_SEG 8): vit.py:292-307."""
    depth, B, T, S = 2, 1, 8, 8
    st = O.seeded_state(depth=depth, frames=T, seed=55, with_order=True)
    m, cfg = build_ref("configs/COIN/step_forecasting.yaml",
                       ["DEV.MATCH_LANG_EMB", True, "MODEL.DROP_PATH", 0.0, "TIMESFORMER.DEPTH", depth,
                        "MODEL.DROP_E", 0.0], st)
    x = O.synthetic_clips(B, 3, S * T, 224, 224, seed=10)
    m.eval()
    with torch.no_grad():
        probs = m(x)
    torch.save({"cfg": {"depth": depth, "B": B, "T": T, "num_seg": S, "state_seed": 55, "clip_seed": 10},
                "probs": probs.clone()}, os.path.join(GOLD, name))
    print(name, tuple(probs.shape), probs.max().item())


def gold_finetune_heads(name, dataset):
    """SURVEY 8a row A14, vit.py:308-322: the fine-tuning path WITHOUT DEV.MATCH_LANG_EMB -- frozen `head` (768 -> 512),
    L2 normalisation, then `head_cls` (512 -> NUM_CLASSES) / temp, or for TRAIN.DATASET Epickitchens the tuple
    (`head_v` 512 -> 97, `head_n` 512 -> 300).  This is what every shipped COIN / EK fine-tuning YAML runs."""
    depth, B, T = 2, 2, 8
    ek = dataset == "Epickitchens"
    yaml_rel = "configs/EK/egocentric_action_classification.yaml" if ek else "configs/COIN/step_classification.yaml"
    ov = ["MODEL.DROP_PATH", 0.0, "TIMESFORMER.DEPTH", depth, "DATA.NUM_FRAMES", T, "DEV.MATCH_LANG_EMB", False]
    vit, build, _ = ref_shims.load_reference()
    cwd = os.getcwd()
    os.chdir(ref_shims.REFERENCE_ROOT)
    try:
        cfg = ref_shims.reference_cfg(yaml_rel, ov)
        m = build.build_model(cfg)
    finally:
        os.chdir(cwd)
    st = O.seeded_state(depth=depth, frames=T, seed=640 + int(ek))
    g = torch.Generator().manual_seed(641 + int(ek))
    own = m.state_dict()
    for k in own:                                        # the fine-tune heads are not part of seeded_state
        if k not in st:
            st[k] = 0.05 * torch.randn(own[k].shape, generator=g)
    st = {k: v for k, v in st.items() if k in own}
    m.load_state_dict(st, strict=True)
    frozen = sorted(k for k, p_ in m.named_parameters() if not p_.requires_grad)
    x = O.synthetic_clips(B, 3, T, 224, 224, seed=12)
    m.train()
    out = m(x)
    outs = list(out) if isinstance(out, tuple) else [out]
    labels = [torch.arange(B) * 29 % o.shape[1] for o in outs]
    loss = sum(F.cross_entropy(o, l) for o, l in zip(outs, labels))
    loss.backward()
    grads = {k: p_.grad for k, p_ in m.named_parameters() if p_.grad is not None}
    keys = [k for k in grads if "head" in k] + [k for k in GRAD_KEYS if k in grads]
    keys += [f"model.blocks.{i}.{s_}" for i in (0, depth - 1) for s_ in BLOCK_GRAD if f"model.blocks.{i}.{s_}" in grads]
    m.eval()
    with torch.no_grad():
        ev = m(x)
    evs = list(ev) if isinstance(ev, tuple) else [ev]
    torch.save({"cfg": {"depth": depth, "B": B, "T": T, "dataset": dataset, "state_seed": 640 + int(ek),
                        "head_seed": 641 + int(ek), "clip_seed": 12, "n_classes": cfg.MODEL.NUM_CLASSES},
                "extra_state": {k: v for k, v in st.items() if "head_" in k},
                "frozen": frozen, "is_tuple": isinstance(out, tuple),
                "outputs": [o.detach().clone() for o in outs], "eval_outputs": [o.clone() for o in evs],
                "labels": labels, "loss": loss.item(), "grads": grad_summary({k: grads[k] for k in keys}),
                "n_with_grad": len(grads)}, os.path.join(GOLD, name))
    print(name, [tuple(o.shape) for o in outs], "loss", loss.item(), "frozen", frozen, "n_with_grad", len(grads))


HT100M_EMB = os.path.join(GOLD, "clip_step_emb_ht100m_vbphrase.pt")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])                 # optional: names of the fixtures to (re)generate

    def want(name):
        return not only or name in only
    if not os.path.exists(COIN_EMB):
        e = torch.load(os.path.join(ref_shims.REFERENCE_ROOT, "data", "clip_step_emb_coin.pth"))
        torch.save(e.clone(), COIN_EMB)     # shipped fixture: fp32 [778,512], un-normalised
    if not os.path.exists(HT100M_EMB):
        e = torch.load(os.path.join(ref_shims.REFERENCE_ROOT, "data", "clip_step_emb_ht100m_vbphrase.pth"))
        torch.save(e.clone(), HT100M_EMB)   # shipped fixture: fp32 [9871,512], the bank procedurevrl_adamw.yaml trains on
    jobs = [("coin_d2_b2.pt", lambda n: gold_coin_matchlang(2, 2, 8, n)),
            ("coin_d12_b4.pt", lambda n: gold_coin_matchlang(12, 4, 8, n)),
            ("coin_d2_t4.pt", lambda n: gold_coin_matchlang(2, 1, 4, n)),
            ("coin_d2_joint.pt", lambda n: gold_coin_matchlang(2, 2, 8, n, attention_type="joint_space_time", with_grads=False)),
            ("coin_d2_spaceonly.pt", lambda n: gold_coin_matchlang(2, 2, 8, n, attention_type="space_only", with_grads=False)),
            ("droppath_d2.pt", gold_droppath),
            ("pretrain_d2_v2.pt", lambda n: gold_pretrain(2, 2, n)),
            ("pretrain_d12_v1.pt", lambda n: gold_pretrain(12, 1, n)),
            ("pretrain_d12_ht100m.pt", lambda n: gold_pretrain(12, 1, n, "./data/clip_step_emb_ht100m_vbphrase.pth", 9871)),
            ("forecast_d2.pt", gold_forecast),
            ("finetune_headcls_d2.pt", lambda n: gold_finetune_heads(n, "howto100m_develop")),
            ("finetune_ek_d2.pt", lambda n: gold_finetune_heads(n, "Epickitchens"))]
    for name, fn in jobs:
        if want(name):
            fn(name)
    return


if __name__ == "__main__":
    main()
