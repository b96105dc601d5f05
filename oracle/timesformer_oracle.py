"""CPU oracle: a plain-PyTorch fp32 restatement of the ProcedureVRL TimeSformer hot path.

TEST INFRASTRUCTURE ONLY -- never imported by the product package `procedurevrl_b200`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it, and only as the checker / the CPU baseline.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the *unmodified reference itself*, generated in the build
container by `oracle/make_golden.py` (reference imported through `oracle/ref_shims.py`) and
committed under `tests/golden/`; `tests/test_oracle_golden.py` re-checks them on every run.

Every function cites the reference lines it restates (paths relative to the reference root).
State is a flat dict keyed exactly like the reference `state_dict()`, all math is
fp32 torch on whatever device the tensors live on (CPU in tests), written functionally (no
nn.Module) so that autograd of the same code yields the reference gradients.

GPU semantics of `check_device_norm` (vit.py:435-440) are assumed throughout: `label_emb` rows are
L2-normalised (on a CPU-only run the reference skips that; see SURVEY.md §7 hard parts).
"""
import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

P = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps=1e-6):
    """nn.LayerNorm(768, eps=1e-6): vit.py:102,108,115,225 with eps from vit.py:488."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x):
    """nn.GELU() default (exact erf form): vit.py:45,50,56."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def linear(x, w, b=None):
    y = x @ w.t()
    return y if b is None else y + b


def attention(p: P, pre: str, x, num_heads=12):
    """Attention.forward, vit.py:75-92 (with_qkv=True, dropouts p=0)."""
    B, N, C = x.shape
    d = C // num_heads
    qkv = linear(x, p[pre + "qkv.weight"], p.get(pre + "qkv.bias"))
    qkv = qkv.reshape(B, N, 3, num_heads, d).permute(2, 0, 3, 1, 4)      # vit.py:78
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (d ** -0.5)                        # vit.py:84 (scale :67)
    attn = attn.softmax(dim=-1)                                           # vit.py:85
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)                     # vit.py:88
    return linear(out, p[pre + "proj.weight"], p[pre + "proj.bias"])      # vit.py:90


def apply_drop_path(x, scale):
    """drop_path, vit_utils.py:140-155: `x / keep * mask` with a per-row (dim 0) mask.
    `scale` is the already materialised per-row factor mask/keep (None = identity: eval / p=0)."""
    if scale is None:
        return x
    return x * scale.view(-1, *([1] * (x.dim() - 1)))


def mlp(p: P, pre: str, x):
    """Mlp.forward, vit.py:54-60 (drop p=0)."""
    return linear(gelu_erf(linear(x, p[pre + "fc1.weight"], p[pre + "fc1.bias"])),
                  p[pre + "fc2.weight"], p[pre + "fc2.bias"])


def block_divided(p: P, pre: str, x, B, T, W, dp=None):
    """Block.forward divided_space_time branch, vit.py:128-158.
    x: [B, 1 + H*W*T, D] with token order (h w t), t fastest.
    dp: optional dict of DropPath row scales {'temporal': [B*H*W], 'spatial': [B*T], 'mlp': [B]}."""
    D = x.shape[-1]
    L = x.shape[1] - 1
    HW = L // T
    dp = dp or {}
    # temporal (vit.py:130-135)
    xt = x[:, 1:, :]
    xt_ = xt.reshape(B, HW, T, D).reshape(B * HW, T, D)                   # 'b (h w t) m -> (b h w) t m'
    res_t = attention(p, pre + "temporal_attn.",
                      layer_norm(xt_, p[pre + "temporal_norm1.weight"], p[pre + "temporal_norm1.bias"]))
    res_t = apply_drop_path(res_t, dp.get("temporal"))
    res_t = res_t.reshape(B, HW * T, D)                                   # '(b h w) t m -> b (h w t) m'
    res_t = linear(res_t, p[pre + "temporal_fc.weight"], p[pre + "temporal_fc.bias"])
    xt = x[:, 1:, :] + res_t
    # spatial (vit.py:138-153)
    init_cls = x[:, 0, :].unsqueeze(1)                                    # [B,1,D]
    cls_tok = init_cls.repeat(1, T, 1).reshape(B * T, 1, D)
    xs = xt.reshape(B, HW, T, D).permute(0, 2, 1, 3).reshape(B * T, HW, D)  # 'b (h w t) m -> (b t) (h w) m'
    xs = torch.cat((cls_tok, xs), 1)
    res_s = attention(p, pre + "attn.", layer_norm(xs, p[pre + "norm1.weight"], p[pre + "norm1.bias"]))
    res_s = apply_drop_path(res_s, dp.get("spatial"))
    cls_out = res_s[:, 0, :].reshape(B, T, D).mean(1, keepdim=True)       # vit.py:147-149
    res_s = res_s[:, 1:, :].reshape(B, T, HW, D).permute(0, 2, 1, 3).reshape(B, HW * T, D)
    # merge + MLP (vit.py:156-157)
    x = torch.cat((init_cls, xt), 1) + torch.cat((cls_out, res_s), 1)
    y = mlp(p, pre + "mlp.", layer_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"]))
    return x + apply_drop_path(y, dp.get("mlp"))


def block_joint(p: P, pre: str, x, dp=None):
    """Block.forward space_only / joint_space_time branch, vit.py:124-127."""
    dp = dp or {}
    a = attention(p, pre + "attn.", layer_norm(x, p[pre + "norm1.weight"], p[pre + "norm1.bias"]))
    x = x + apply_drop_path(a, dp.get("attn"))
    y = mlp(p, pre + "mlp.", layer_norm(x, p[pre + "norm2.weight"], p[pre + "norm2.bias"]))
    return x + apply_drop_path(y, dp.get("mlp"))


def patch_embed(p: P, x, patch=16):
    """PatchEmbed.forward, vit.py:174-180: 'b c t h w -> (b t) c h w', Conv2d(k=s=16), flatten.
    Because stride == kernel the conv is an im2col GEMM with K index (c, kh, kw)."""
    B, C, T, H, Wd = x.shape
    x = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, Wd)
    x = F.conv2d(x, p["model.patch_embed.proj.weight"], p["model.patch_embed.proj.bias"], stride=patch)
    W = x.size(-1)
    return x.flatten(2).transpose(1, 2), T, W                             # [(B T), N, D]


def embed_tokens(p: P, x, B, T, W, attention_type="divided_space_time"):
    """forward_features embeds, vit.py:371-407 (pos_drop/time_drop p=0)."""
    D = x.shape[-1]
    cls = p["model.cls_token"].expand(x.size(0), -1, -1)
    x = torch.cat((cls, x), dim=1)
    pos = p["model.pos_embed"]
    if x.size(1) != pos.size(1):                                          # vit.py:375-386 nearest resize
        cls_pos = pos[0, 0, :].unsqueeze(0).unsqueeze(1)
        other = pos[0, 1:, :].unsqueeze(0).transpose(1, 2)
        Pp = int(other.size(2) ** 0.5)
        H = x.size(1) // W
        other = F.interpolate(other.reshape(1, x.size(2), Pp, Pp), size=(H, W), mode="nearest")
        pos = torch.cat((cls_pos, other.flatten(2).transpose(1, 2)), 1)
    x = x + pos
    if attention_type != "space_only":                                    # vit.py:393-407
        cls_tokens = x[:B, 0, :].unsqueeze(1)
        x = x[:, 1:]
        N = x.shape[1]
        x = x.reshape(B, T, N, D).permute(0, 2, 1, 3).reshape(B * N, T, D)  # '(b t) n m -> (b n) t m'
        te = p["model.time_embed"]
        if T != te.size(1):                                               # vit.py:398-402
            te = F.interpolate(te.transpose(1, 2), size=(T), mode="nearest").transpose(1, 2)
        x = x + te
        x = x.reshape(B, N * T, D)                                        # '(b n) t m -> b (n t) m'
        x = torch.cat((cls_tokens, x), dim=1)
    return x


def forward_features(p: P, x, depth, attention_type="divided_space_time", drop_scales=None,
                     taps: Optional[dict] = None):
    """VisionTransformer.forward_features, vit.py:365-423 -> cls feature [B, D].
    drop_scales: optional list (len depth) of per-block DropPath scale dicts."""
    B = x.shape[0]
    x, T, W = patch_embed(p, x)
    if taps is not None:
        taps["patch"] = x
    x = embed_tokens(p, x, B, T, W, attention_type)
    if taps is not None:
        taps["embed"] = x
    for i in range(depth):
        dp = drop_scales[i] if drop_scales is not None else None
        pre = f"model.blocks.{i}."
        if attention_type == "divided_space_time":
            x = block_divided(p, pre, x, B, T, W, dp)
        else:
            x = block_joint(p, pre, x, dp)
        if taps is not None:
            taps[f"block{i}"] = x
    if attention_type == "space_only":                                    # vit.py:414-416
        x = x.reshape(B, T, x.shape[1], x.shape[2]).mean(1)
    x = layer_norm(x, p["model.norm.weight"], p["model.norm.bias"])       # vit.py:418
    return x[:, 0]                                                        # vit.py:421


def l2_normalize(x):
    """x / x.norm(dim=1, keepdim=True): vit.py:302,306,311,316,333,340,431."""
    return x / x.norm(dim=1, keepdim=True)


def normalized_label_emb(label_emb):
    """check_device_norm with norm=True on first GPU use, vit.py:435-440."""
    return label_emb / label_emb.norm(dim=1, keepdim=True)


def video_embedding(p: P, feat):
    """head + L2 norm, vit.py:301-303."""
    return l2_normalize(linear(feat, p["model.head.weight"], p["model.head.bias"]))


def similarity_logits(emb, label_emb_n, temp=0.02):
    """x @ label_emb.t() / temp, vit.py:307,334,341,432 (temp = DEV.TEMP, defaults.py:53)."""
    return emb @ label_emb_n.t() / temp


def match_lang_forward(p: P, x, label_emb_n, depth=12, temp=0.02, training=True,
                       attention_type="divided_space_time", drop_scales=None, taps=None):
    """VisionTransformer.forward for DEV.MATCH_LANG_EMB with no text (zero-shot / COIN-shape
    parity config): vit.py:296-307, 355-358.  training -> raw logits, eval -> softmax probs."""
    feat = forward_features(p, x, depth, attention_type, drop_scales, taps)
    if taps is not None:
        taps["feat"] = feat
    logits = similarity_logits(video_embedding(p, feat), label_emb_n, temp)
    return logits if training else logits.softmax(dim=1)


def finetune_cls_forward(p: P, x, depth=12, temp=0.02, training=True, drop_scales=None):
    """VisionTransformer.forward, classification fine-tune branch: vit.py:315-322, 355-358."""
    feat = forward_features(p, x, depth, drop_scales=drop_scales)
    e = l2_normalize(linear(feat, p["model.head.weight"], p["model.head.bias"]))
    logits = linear(e, p["model.head_cls.weight"], p["model.head_cls.bias"]) / temp
    return logits if training else logits.softmax(dim=1)


def finetune_ek_forward(p: P, x, depth=12, temp=0.02, training=True, drop_scales=None):
    """VisionTransformer.forward, EPIC-Kitchens fine-tune branch (TRAIN.DATASET Epickitchens): vit.py:308-314 -- frozen
    `head`, L2 normalisation, then the verb (97) and noun (300) heads, each / temp; returned as the tuple (v, n) --
    in eval mode too: the early `return (v, n)` at vit.py:314 bypasses the test-time softmax of vit.py:355-356."""
    feat = forward_features(p, x, depth, drop_scales=drop_scales)
    e = l2_normalize(linear(feat, p["model.head.weight"], p["model.head.bias"]))
    v = linear(e, p["model.head_v.weight"], p["model.head_v.bias"]) / temp
    n = linear(e, p["model.head_n.weight"], p["model.head_n.bias"]) / temp
    return (v, n)


# ----------------------------------------------------------------------------------------------
# clip-level order / diffusion transformer (tfm_model.py) -- SURVEY.md §8f-1, stays in PyTorch
# ----------------------------------------------------------------------------------------------
@dataclass
class OrderDraws:
    """The random draws of one pretrain forward, in the order the reference makes them:
    mask_inds  (tfm_model.py:145), pad_start per sample (tfm_model.py:280-283; == max_len when the
    mask is the last token), noise per level (tfm_model.py:180), rand_inds (vit.py:345)."""
    mask_inds: torch.Tensor        # [Bv] int64
    pad_start: torch.Tensor        # [Bv] int64
    noise: torch.Tensor            # [levels, Bv, C]
    rand_inds: torch.Tensor        # [Bv*max_len] int64 (first Bv*ORDER_RECOG_BATCH kept)


def diffusion_schedule(levels=4):
    """configure_diffusion, tfm_model.py:106-127 with linear_beta_schedule diffusion_model.py:328-331."""
    betas = torch.linspace(0.0001, 0.02, levels)
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    return torch.sqrt(alphas_cumprod), torch.sqrt(1.0 - alphas_cumprod)


def sinusoidal_embedding(t, dim):
    """SinusoidalPositionEmbeddings.forward, diffusion_model.py:39-46."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, device=t.device) * -e)
    e = t[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def _mha(p: P, pre: str, x, heads, pad_mask):
    """nn.MultiheadAttention(x,x,x, key_padding_mask) as called in tfm_model.py:46-48. x: [S,B,C]."""
    S, B, C = x.shape
    d = C // heads
    qkv = linear(x, p[pre + "in_proj_weight"], p[pre + "in_proj_bias"])
    q, k, v = qkv.split(C, dim=-1)
    def sh(t):
        return t.reshape(S, B * heads, d).transpose(0, 1)                # [B*h, S, d]
    q, k, v = sh(q), sh(k), sh(v)
    att = (q * (d ** -0.5)) @ k.transpose(1, 2)                           # [B*h,S,S]
    if pad_mask is not None:
        m = pad_mask.view(B, 1, 1, S).expand(B, heads, 1, S).reshape(B * heads, 1, S)
        att = att.masked_fill(m, float("-inf"))
    att = att.softmax(dim=-1)
    out = (att @ v).transpose(0, 1).reshape(S, B, C)
    return linear(out, p[pre + "out_proj.weight"], p[pre + "out_proj.bias"])


def _resblock(p: P, pre: str, x, heads, pad_mask):
    """ResidualAttentionBlock.forward, tfm_model.py:50-53 (LN eps 1e-5 default, QuickGELU :27-29)."""
    h = layer_norm(x, p[pre + "ln_1.weight"], p[pre + "ln_1.bias"], eps=1e-5)
    x = x + _mha(p, pre + "attn.", h, heads, pad_mask)
    h = layer_norm(x, p[pre + "ln_2.weight"], p[pre + "ln_2.bias"], eps=1e-5)
    h = linear(h, p[pre + "mlp.c_fc.weight"], p[pre + "mlp.c_fc.bias"])
    h = h * torch.sigmoid(1.702 * h)
    return x + linear(h, p[pre + "mlp.c_proj.weight"], p[pre + "mlp.c_proj.bias"])


def _time_mlp(p: P, pre: str, t, hidden):
    """time_mlp, tfm_model.py:89-94: sinusoid(hidden/4) -> Linear -> GELU -> Linear."""
    e = sinusoidal_embedding(t.float() if t.dtype != torch.float32 else t, hidden // 4)
    e = linear(e, p[pre + "time_mlp.1.weight"], p[pre + "time_mlp.1.bias"])
    return linear(gelu_erf(e), p[pre + "time_mlp.3.weight"], p[pre + "time_mlp.3.bias"])


def _order_level(p, pre, clip_feats, mask_inds, bs_inds, t_index, pad_mask, layers, heads, max_len):
    """One denoising level, tfm_model.py:187-196 (shared by training :172-197 and forecast :222-244)."""
    Bv, C = clip_feats.shape[1], clip_feats.shape[2]
    dev = clip_feats.device
    type_emb = p[pre + "type_embedding.weight"][0].expand(max_len, Bv, C).clone()
    type_emb[mask_inds, bs_inds] = p[pre + "type_embedding.weight"][1]
    temp_emb = p[pre + "temporalEmbedding.weight"][:max_len].unsqueeze(1).expand(max_len, Bv, C)
    t = torch.full((Bv,), t_index, device=dev, dtype=torch.long)
    h = clip_feats + type_emb + temp_emb
    h = h + _time_mlp(p, pre, t, C).unsqueeze(0)                          # 't c -> b t c' broadcast over seq
    for i in range(layers):
        h = _resblock(p, f"{pre}temporalModelling.resblocks.{i}.", h, heads, pad_mask)
    return h[mask_inds, bs_inds]


def order_tfm_pretrain(p: P, video_emb, draws: OrderDraws, max_len=9, layers=4, heads=8,
                       pre="model.order_tfm."):
    """DiffusionTransformer.forward(is_pretrain=True), tfm_model.py:137-156 + 165-204 + 272-289.
    Returns (denoised_final [Bv,C], mask_inds, [x0_target, all_denoised] each [levels*Bv,C], all_denoised)."""
    C = video_emb.shape[1]
    clip_feats = video_emb.reshape(-1, max_len, C).transpose(0, 1)        # '(b t) c -> t b c'
    Bv = clip_feats.shape[1]
    dev = video_emb.device
    bs_inds = torch.arange(Bv, device=dev)
    mask_inds = draws.mask_inds.to(dev)
    x0 = clip_feats[mask_inds, bs_inds]                                   # tfm_model.py:148
    # pad_sequence (tfm_model.py:272-289): tokens at positions >= pad_start -> pad embedding (in place)
    pos = torch.arange(max_len, device=dev).unsqueeze(1)                  # [S,1]
    pad = pos >= draws.pad_start.to(dev).unsqueeze(0)                     # [S,Bv]
    clip_feats = torch.where(pad.unsqueeze(-1), p[pre + "pad_embedding.weight"][0].expand(max_len, Bv, C),
                             clip_feats)
    pad_mask = pad.t()                                                    # [Bv,S] True = ignore
    sqrt_ac, sqrt_1mac = diffusion_schedule(layers)
    sqrt_ac, sqrt_1mac = sqrt_ac.to(dev), sqrt_1mac.to(dev)
    outs: List[torch.Tensor] = []
    denoised = None
    for lvl in range(layers):                                             # tfm_model.py:172-197
        t_index = layers - 1 - lvl
        src = x0 if lvl == 0 else denoised
        noisy = sqrt_ac[t_index] * src.detach() + sqrt_1mac[t_index] * draws.noise[lvl].to(dev)  # ennoise :291-302
        feats = clip_feats.clone()
        feats[mask_inds, bs_inds] = noisy
        denoised = _order_level(p, pre, feats, mask_inds, bs_inds, t_index, pad_mask, layers, heads, max_len)
        outs.append(denoised)
    x0_target = x0.unsqueeze(0).expand(layers, -1, -1).reshape(-1, C)     # tfm_model.py:201
    all_denoised = torch.cat(outs)                                        # tfm_model.py:202
    return denoised, mask_inds, [x0_target, all_denoised], all_denoised


def order_tfm_forecast(p: P, video_emb, num_seg, max_len=9, layers=4, heads=8, pre="model.order_tfm."):
    """DiffusionTransformer.diffusion_signal_forecast, tfm_model.py:206-249 (noise is zeros, :217)."""
    C = video_emb.shape[1]
    feats0 = video_emb.reshape(-1, num_seg, C).transpose(0, 1)
    Bv = feats0.shape[1]
    dev = video_emb.device
    bs_inds = torch.arange(Bv, device=dev)
    mask_inds = torch.full((Bv,), max_len - 1, device=dev, dtype=torch.long)
    orig = torch.cat((feats0, torch.zeros(1, Bv, C, device=dev)), dim=0)
    sqrt_ac, _ = diffusion_schedule(layers)
    feats = orig.clone()
    for lvl in range(layers):
        t_index = layers - 1 - lvl
        if lvl != 0:
            feats = feats.clone()
            feats[mask_inds, bs_inds] = sqrt_ac[t_index].to(dev) * den.detach()   # + sqrt(1-ac)*0
        den = _order_level(p, pre, feats, mask_inds, bs_inds, t_index, None, layers, heads, max_len)
        feats = orig.clone()
        feats[mask_inds, bs_inds] = den
    return feats[mask_inds, bs_inds]


# ----------------------------------------------------------------------------------------------
# pretrain forward + loss
# ----------------------------------------------------------------------------------------------
def pseudo_labels(text_emb, vis_emb, label_emb_n, temp=0.02):
    """get_pseudo_labels, vit.py:425-433, with the frozen CLIP text tower's output `text_emb`
    supplied pre-extracted (north star; the tower itself is out of scope)."""
    e = (text_emb + vis_emb) / 2.0
    return similarity_logits(l2_normalize(e), label_emb_n, temp)


def pretrain_forward(p: P, frames, text_emb, vis_emb, label_emb_n, draws: OrderDraws, depth=12,
                     temp=0.02, max_len=9, order_layers=4, order_recog_batch=9, drop_scales=None,
                     taps=None):
    """VisionTransformer.forward, order-pretraining branch: vit.py:285-352.
    frames [Bv, max_len, 3, T, H, W] -> (pred [Bv*R + L*Bv, K], teacher [same], [x0_target, denoised])."""
    Bv = frames.shape[0]
    x = frames.reshape(Bv * max_len, *frames.shape[2:])                   # vit.py:291
    feat = forward_features(p, x, depth, drop_scales=drop_scales, taps=taps)
    video_emb = video_embedding(p, feat)                                  # vit.py:301-303
    logits = similarity_logits(video_emb, label_emb_n, temp)              # vit.py:307
    teacher = pseudo_labels(text_emb, vis_emb, label_emb_n, temp)         # vit.py:327
    pred_emb, mask_inds, mse_pair, inter = order_tfm_pretrain(p, video_emb, draws, max_len, order_layers)
    # vit.py:333-334 computes mask_pred from pred_emb; it is not returned (dead value) -> skipped.
    masked_teacher = teacher.reshape(Bv, max_len, -1)[torch.arange(Bv), mask_inds]      # vit.py:337,360-363
    inter_pred = similarity_logits(l2_normalize(inter), label_emb_n, temp)             # vit.py:340-341
    inter_teacher = masked_teacher.unsqueeze(0).expand(order_layers, -1, -1).reshape(-1, masked_teacher.size(-1))
    keep = draws.rand_inds[: Bv * order_recog_batch]                      # vit.py:345-347
    pred = torch.cat((logits[keep], inter_pred), dim=0)                   # vit.py:350
    teach = torch.cat((teacher[keep], inter_teacher), dim=0)              # vit.py:351
    if taps is not None:
        taps.update(feat=feat, video_emb=video_emb, logits=logits)
    return pred, teach, mse_pair


def topk_teacher(teacher_logits, topk=5):
    """train_net.py:153-158 (no_grad): softmax, keep entries equal to one of the row's top-k values
    (literal broadcast-compare, so exact ties count once per matching top-k slot), renormalise."""
    t = F.softmax(teacher_logits, 1)
    if topk != 0:
        tv = t.topk(k=topk, dim=1)[0]
        t = (t.unsqueeze(1) * (t.unsqueeze(1) == tv.unsqueeze(2)).float()).sum(1)
        t = t / t.sum(1, keepdim=True)
    return t


def pretrain_loss(pred, teacher_logits, mse_pair, topk=5):
    """train_net.py:131-133,152-162: KLDivLoss(batchmean)(log_softmax(pred), teacher_topk) + MSE(mean)."""
    with torch.no_grad():
        t = topk_teacher(teacher_logits, topk)
    loss1 = F.kl_div(F.log_softmax(pred, dim=1), t, reduction="batchmean")
    loss2 = F.mse_loss(mse_pair[0], mse_pair[1], reduction="mean")
    return loss1 + loss2, loss1, loss2


# ----------------------------------------------------------------------------------------------
# seeded state / inputs shared by the golden generator, the tests and the bench
# ----------------------------------------------------------------------------------------------
def seeded_state(depth=12, frames=8, embed_dim=768, emb_dim=512, mlp_ratio=4, patch=16, img=224,
                 seed=0, with_order=False, order_layers=4, max_len=9, head_cls=0) -> P:
    """Deterministic synthetic weights with the reference state_dict schema.
    NOT the reference initialiser: per SURVEY.md §3.3 a fresh reference model has all-zero
    temporal_fc / time_embed, which would hide the temporal branch, so every tensor gets
    non-trivial values (trunc-normal-ish N(0, .02) weights, small biases, LN weight ~ 1)."""
    g = torch.Generator().manual_seed(seed)
    D, H = embed_dim, embed_dim * mlp_ratio
    n_patches = (img // patch) ** 2

    def w(*s, std=0.02):
        return (torch.randn(*s, generator=g) * std).clamp_(-2 * std, 2 * std)

    p: P = {}
    p["model.cls_token"] = w(1, 1, D)
    p["model.pos_embed"] = w(1, n_patches + 1, D)
    p["model.time_embed"] = w(1, frames, D)
    p["model.patch_embed.proj.weight"] = w(D, 3, patch, patch)
    p["model.patch_embed.proj.bias"] = w(D)
    for i in range(depth):
        b = f"model.blocks.{i}."
        for n in ("norm1", "temporal_norm1", "norm2"):
            p[b + n + ".weight"] = 1.0 + w(D, std=0.1)
            p[b + n + ".bias"] = w(D, std=0.05)
        for a in ("attn", "temporal_attn"):
            p[b + a + ".qkv.weight"] = w(3 * D, D, std=0.04)
            p[b + a + ".qkv.bias"] = w(3 * D)
            p[b + a + ".proj.weight"] = w(D, D)
            p[b + a + ".proj.bias"] = w(D)
        p[b + "temporal_fc.weight"] = w(D, D)
        p[b + "temporal_fc.bias"] = w(D)
        p[b + "mlp.fc1.weight"] = w(H, D)
        p[b + "mlp.fc1.bias"] = w(H)
        p[b + "mlp.fc2.weight"] = w(D, H)
        p[b + "mlp.fc2.bias"] = w(D)
    p["model.norm.weight"] = 1.0 + w(D, std=0.1)
    p["model.norm.bias"] = w(D, std=0.05)
    p["model.head.weight"] = w(emb_dim, D, std=0.04)
    p["model.head.bias"] = w(emb_dim)
    if head_cls:
        p["model.head_cls.weight"] = w(head_cls, emb_dim, std=0.04)
        p["model.head_cls.bias"] = w(head_cls)
    if with_order:
        o = "model.order_tfm."
        C = emb_dim
        p[o + "pad_embedding.weight"] = w(1, C, std=0.01)
        p[o + "type_embedding.weight"] = w(2, C, std=0.5)
        p[o + "temporalEmbedding.weight"] = w(max_len, C, std=0.01)
        for i in range(order_layers):
            r = f"{o}temporalModelling.resblocks.{i}."
            p[r + "attn.in_proj_weight"] = w(3 * C, C, std=C ** -0.5)
            p[r + "attn.in_proj_bias"] = w(3 * C)
            p[r + "attn.out_proj.weight"] = w(C, C, std=(C ** -0.5) * ((2 * order_layers) ** -0.5))
            p[r + "attn.out_proj.bias"] = w(C)
            p[r + "ln_1.weight"] = 1.0 + w(C, std=0.1)
            p[r + "ln_1.bias"] = w(C, std=0.05)
            p[r + "mlp.c_fc.weight"] = w(4 * C, C, std=(2 * C) ** -0.5)
            p[r + "mlp.c_fc.bias"] = w(4 * C)
            p[r + "mlp.c_proj.weight"] = w(C, 4 * C, std=(C ** -0.5) * ((2 * order_layers) ** -0.5))
            p[r + "mlp.c_proj.bias"] = w(C)
            p[r + "ln_2.weight"] = 1.0 + w(C, std=0.1)
            p[r + "ln_2.bias"] = w(C, std=0.05)
        p[o + "time_mlp.1.weight"] = w(C, C // 4, std=0.05)
        p[o + "time_mlp.1.bias"] = w(C)
        p[o + "time_mlp.3.weight"] = w(C, C, std=0.04)
        p[o + "time_mlp.3.bias"] = w(C)
    return p


def synthetic_clips(*shape, seed=0):
    """SURVEY.md §8d synthetic frames: uint8 U[0,255] -> /255 -> (x-0.45)/0.225 (defaults.py:510,516)."""
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    return (u8.float() / 255.0 - 0.45) / 0.225


def synthetic_draws(Bv, max_len=9, levels=4, C=512, seed=0) -> OrderDraws:
    g = torch.Generator().manual_seed(seed)
    mask = torch.randint(0, max_len, (Bv,), generator=g)
    pad = torch.empty(Bv, dtype=torch.long)
    for i in range(Bv):
        m = int(mask[i])
        pad[i] = max_len if m + 1 == max_len else int(torch.randint(m + 1, max_len, (1,), generator=g))
    noise = torch.randn(levels, Bv, C, generator=g)
    perm = torch.randperm(Bv * max_len, generator=g)
    return OrderDraws(mask, pad, noise, perm)
