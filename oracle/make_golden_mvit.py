"""Golden vectors for the MViTv2 encoder (BASELINE config 5), produced by the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY (runs in the build container, where /root/reference exists).  For each case below it builds the
reference `MViT` (lib/models/mvit.py -> slowfast_mvit/mvit.py) through `oracle/ref_shims.py` from the shipped
procedurevrl_mvitv2_adamw.yaml with a reduced geometry, loads `mvit_oracle.seeded_state` into it, runs the
DEV.MATCH_LANG_EMB forward on seeded clips, a cross-entropy loss and its backward, and writes
tests/golden/mvit_<case>.pt: the MVIT.* keys, the parameter schema (names + shapes), logits, the encoder's cls
feature, the cls row + norm of every block output, the loss and a summary of every parameter gradient.
The full-size geometry of the shipped YAML (16 x 224, depth 16: widths, heads, strides, token counts) is recorded in
tests/golden/mvit_full_geometry.json for `mvit_oracle.geometry`.

    python oracle/make_golden_mvit.py
"""
import importlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import mvit_oracle as MO  # noqa: E402
import ref_shims  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")
YAML = "configs/HowTo100M/procedurevrl_mvitv2_adamw.yaml"
COIN_EMB = os.path.join(GOLD, "clip_step_emb_coin.pt")

CASES = {
    # depth 4: Q stride at blocks 1 and 3 (with width / head doubling), KV strides 4 -> 2 -> 2 -> 1, grid 2 x 16 x 16
    "d4_t4_c64": dict(frames=4, crop=64, B=2, seed=101, ov=[
        "MVIT.DEPTH", 4, "MVIT.DIM_MUL", [[1, 2.0], [3, 2.0]], "MVIT.HEAD_MUL", [[1, 2.0], [3, 2.0]],
        "MVIT.POOL_Q_STRIDE", [[0, 1, 1, 1], [1, 1, 2, 2], [2, 1, 1, 1], [3, 1, 2, 2]],
        "MVIT.POOL_KV_STRIDE_ADAPTIVE", [1, 4, 4]]),
    # depth 3, longer clip, rectangular ratio between Q and KV grids (KV stride 8 on a 24 x 24 grid), one block without
    # any Q entry (no Q pooling conv at all)
    "d3_t8_c96": dict(frames=8, crop=96, B=1, seed=202, ov=[
        "MVIT.DEPTH", 3, "MVIT.DIM_MUL", [[1, 2.0]], "MVIT.HEAD_MUL", [[1, 2.0]],
        "MVIT.POOL_Q_STRIDE", [[0, 1, 1, 1], [1, 1, 2, 2]], "MVIT.POOL_KV_STRIDE_ADAPTIVE", [1, 8, 8]]),
}


def build_reference(frames, crop, ov):
    ref_shims.install()
    cwd = os.getcwd()
    os.chdir(ref_shims.REFERENCE_ROOT)
    try:
        cfg = ref_shims.reference_cfg(YAML, [
            "DATA.NUM_FRAMES", frames, "DATA.TRAIN_CROP_SIZE", crop, "DATA.TEST_CROP_SIZE", crop,
            "TRAIN.LABEL_EMB", "", "DEV.MATCH_LANG_EMB", True, "DEV.TEST_LANG_EMB", "data/clip_step_emb_coin.pth",
            "MODEL.TEXT_MODEL", "", "DEV.ORDER_PRETRAIN_ENABLED", False, "MODEL.NUM_CLASSES", 778] + ov)
        mv = importlib.import_module("lib.models.mvit")
        m = mv.MViT(cfg)
    finally:
        os.chdir(cwd)
    return m, cfg


def mvit_keys(cfg):
    keep = ("DEPTH", "NUM_HEADS", "EMBED_DIM", "PATCH_KERNEL", "PATCH_STRIDE", "PATCH_PADDING", "MLP_RATIO", "DIM_MUL",
            "HEAD_MUL", "POOL_KVQ_KERNEL", "POOL_KV_STRIDE_ADAPTIVE", "POOL_Q_STRIDE", "MODE", "POOL_FIRST", "SEPARATE_QKV",
            "CLS_EMBED_ON", "USE_ABS_POS", "REL_POS_SPATIAL", "REL_POS_TEMPORAL", "RESIDUAL_POOLING", "DIM_MUL_IN_ATT", "NORM")
    def plain(v):
        return [plain(x) for x in v] if isinstance(v, (list, tuple)) else v
    return {k: plain(cfg.MVIT[k]) for k in keep}


def one_case(name, frames, crop, B, seed, ov):
    m, cfg = build_reference(frames, crop, ov)
    mv = mvit_keys(cfg)
    geo = MO.geometry(mv, frames, crop)
    own = m.state_dict()
    enc_shapes = {k: tuple(v.shape) for k, v in own.items() if k.startswith(MO.PRE)}
    assert enc_shapes == {k: tuple(v) for k, v in MO.param_shapes(geo).items()}, "oracle schema != reference schema"
    shapes = dict(enc_shapes)
    shapes["model.head.weight"], shapes["model.head.bias"] = tuple(own["model.head.weight"].shape), tuple(own["model.head.bias"].shape)
    state = MO.seeded_state(shapes, seed)
    m.load_state_dict(state, strict=True)
    for p in m.parameters():
        p.requires_grad_(True)
    m.model.label_emb = m.model.label_emb / m.model.label_emb.norm(dim=1, keepdim=True)   # GPU semantics of check_device_norm
    m.train()
    taps, hooks = [], []
    for blk in m.model.video_encoder.blocks:
        hooks.append(blk.register_forward_hook(lambda mod, inp, out: taps.append(out[0].detach())))
    feat = []
    hooks.append(m.model.video_encoder.register_forward_hook(lambda mod, inp, out: feat.append(out.detach().clone())))
    x = MO.synthetic_clips(B, frames, crop, seed + 1)
    logits = m(x)
    labels = (torch.arange(B) * 131 + 7) % logits.shape[1]
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    for h in hooks:
        h.remove()
    grads = {k: {"norm": p.grad.norm().item(), "sum": p.grad.double().sum().item(), "head": p.grad.flatten()[:32].clone()}
             for k, p in m.named_parameters() if p.grad is not None}
    out = {"cfg": {"frames": frames, "crop": crop, "B": B, "seed": seed, "mvit": mv}, "shapes": shapes,
           "logits": logits.detach(), "labels": labels, "loss": loss.item(), "feat": feat[0],
           "taps": [{"cls": t[:, 0].clone(), "norm": t.norm().item(), "shape": tuple(t.shape)} for t in taps],
           "grads": grads, "n_grads": len(grads)}
    path = os.path.join(GOLD, f"mvit_{name}.pt")
    torch.save(out, path)
    print(name, "logits", tuple(logits.shape), "loss", round(loss.item(), 5), "grads", len(grads), os.path.getsize(path), "bytes")


def pretrain_case(name="pretrain_d4_t4_c64", case="d4_t4_c64", Bv=1, seed=303):
    """BASELINE config 5 as the driver runs it (tools/train_net.py:146-162 on procedurevrl_mvitv2_adamw.yaml): one video of
    9 clips through the MViT encoder, head, order transformer, teacher and the KL(top-k) + MSE loss, and its backward --
    reduced geometry, COIN step bank as LABEL_EMB, the CLIP text tower replaced by pre-extracted embeddings (north star).
    Records the reference's random draws so that the mirror can replay them."""
    import make_golden as MG
    import timesformer_oracle as TO
    import torch.nn.functional as F
    c = CASES[case]
    frames, crop, max_len = c["frames"], c["crop"], 9
    ref_shims.install()
    cwd = os.getcwd()
    os.chdir(ref_shims.REFERENCE_ROOT)
    try:
        cfg = ref_shims.reference_cfg(YAML, [
            "DATA.NUM_FRAMES", frames, "DATA.TRAIN_CROP_SIZE", crop, "DATA.TEST_CROP_SIZE", crop,
            "TRAIN.LABEL_EMB", "./data/clip_step_emb_coin.pth", "MODEL.NUM_CLASSES", 778, "TRAIN.TEXT", "preextracted"] + c["ov"])
        mv = importlib.import_module("lib.models.mvit")
        m = mv.MViT(cfg)
    finally:
        os.chdir(cwd)
    own = m.state_dict()
    shapes = {k: tuple(v.shape) for k, v in own.items() if k.startswith(MO.PRE) or k.startswith("model.head.")}
    state = MO.seeded_state(shapes, seed)
    order = {k: v for k, v in TO.seeded_state(depth=1, frames=8, seed=seed + 1, with_order=True).items()
             if k.startswith("model.order_tfm.")}
    assert {k: tuple(v.shape) for k, v in order.items()} == {k: tuple(v.shape) for k, v in own.items() if k.startswith("model.order_tfm.")}
    state.update(order)
    for k in own:
        if k.startswith("model.text_model."):
            state[k] = own[k]
    m.load_state_dict(state, strict=True)
    m.model.label_emb = m.model.label_emb / m.model.label_emb.norm(dim=1, keepdim=True)
    g = torch.Generator().manual_seed(seed + 2)
    text_emb = 0.4 * torch.randn(Bv * max_len, 512, generator=g)
    vis_emb = 0.4 * torch.randn(Bv * max_len, 512, generator=g)
    m.model.text_model.encode_text = lambda ids: text_emb
    x = MO.synthetic_clips(Bv * max_len, frames, crop, seed + 3).reshape(Bv, max_len, 3, frames, crop, crop)
    meta = {"clip_text_ids": torch.zeros(Bv * max_len, 77, dtype=torch.long), "clip_vis_feat": vis_emb}
    m.train()
    m.model.text_model.eval()
    torch.manual_seed(11)
    log = []
    with MG.record_rng(log):
        pred, teacher, mse = m([x, meta])
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
    with torch.no_grad():                                       # tools/train_net.py:152-162
        tp = F.softmax(teacher, 1)
        tp = (tp.unsqueeze(1) * (tp.unsqueeze(1) == tp.topk(k=cfg.TRAIN.TOPK, dim=1)[0].unsqueeze(2)).float()).sum(1)
        tp = tp / tp.sum(1, keepdim=True)
    loss1 = torch.nn.KLDivLoss(reduction="batchmean")(F.log_softmax(pred, dim=1), tp)
    loss2 = torch.nn.MSELoss(reduction="mean")(mse[0], mse[1])
    (loss1 + loss2).backward()
    grads = {k: {"norm": p.grad.norm().item(), "sum": p.grad.double().sum().item(), "head": p.grad.flatten()[:32].clone()}
             for k, p in m.named_parameters() if p.grad is not None}
    out = {"cfg": {"case": case, "frames": frames, "crop": crop, "Bv": Bv, "seed": seed, "mvit": mvit_keys(cfg), "topk": cfg.TRAIN.TOPK},
           "shapes": {k: tuple(v.shape) for k, v in own.items() if not k.startswith("model.text_model.")},
           "draws": {"mask_inds": mask_inds, "pad_start": pad_start, "noise": noise, "rand_inds": rand_inds},
           "pred": pred.detach().clone(), "teacher": teacher.detach().clone(), "mse0": mse[0].detach().clone(),
           "mse1": mse[1].detach().clone(), "loss1": loss1.item(), "loss2": loss2.item(), "grads": grads, "n_grads": len(grads)}
    path = os.path.join(GOLD, f"mvit_{name}.pt")
    torch.save(out, path)
    print(name, "pred", tuple(pred.shape), "loss", round(loss1.item(), 5), round(loss2.item(), 5), "grads", len(grads),
          os.path.getsize(path), "bytes")


def full_geometry():
    """The shipped 16 x 224 model: only its construction is needed (no forward)."""
    m, cfg = build_reference(16, 224, [])
    enc = m.model.video_encoder
    blocks = [dict(dim=b.dim, dim_out=b.attn.dim_out, heads=b.attn.num_heads,
                   pool_q=None if b.attn.pool_q is None else [list(b.attn.pool_q.kernel_size), list(b.attn.pool_q.stride)],
                   pool_kv=None if b.attn.pool_k is None else [list(b.attn.pool_k.kernel_size), list(b.attn.pool_k.stride)],
                   pool_skip=None if b.pool_skip is None else [list(b.pool_skip.kernel_size), list(b.pool_skip.stride)],
                   rel_sp=b.attn.rel_pos_h.shape[0], rel_t=b.attn.rel_pos_t.shape[0], has_proj=hasattr(b, "proj"))
              for b in enc.blocks]
    shapes = {k: list(v.shape) for k, v in m.state_dict().items() if k.startswith(MO.PRE)}
    out = {"mvit": mvit_keys(cfg), "frames": 16, "crop": 224, "patch_dims": list(enc.patch_dims), "blocks": blocks,
           "out_dim": enc.norm.weight.shape[0], "n_params": sum(v.numel() for k, v in m.state_dict().items() if k.startswith(MO.PRE)),
           "shapes": shapes}
    with open(os.path.join(GOLD, "mvit_full_geometry.json"), "w") as f:
        json.dump(out, f)
    print("full geometry:", out["patch_dims"], "->", out["out_dim"], out["n_params"], "encoder parameters")


def main():
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    for name, c in CASES.items():
        one_case(name, **c)
    pretrain_case()
    full_geometry()


if __name__ == "__main__":
    main()
