"""Drop-in mirror of the reference's MViTv2 model (BASELINE config 5: reference lib/models/mvit.py wrapper +
lib/models/slowfast_mvit/{mvit,attention,stem_helper,common}.py encoder), running on libpvrl_sm100.so.

Same surface as the reference: registry entry `MViT(cfg)` (MODEL.MODEL_NAME: MViT) whose `.model` has
`.video_encoder` (cls_token, patch_embed.proj, blocks.N.{norm1, attn.{qkv, proj, pool_q/k/v, norm_q/k/v, rel_pos_h/w/t},
norm2, mlp.fc1/fc2, proj}, norm), `.head`, `.order_tfm`, `.text_model` -- identical `state_dict()` schema
(tests/golden/mvit_full_geometry.json, written by the unmodified reference), identical `forward` signatures / returns.
The nn.Linear / nn.LayerNorm / nn.Conv3d sub-modules are *parameter containers* (the reference's names, shapes and
initialisers); the math is scheduled below over

    tc_functional.tc_linear / tc_mlp     every Linear (75 % of the FLOPs) on the tcgen05 GEMM, GELU fused in its epilogues
    mvit_functional.*                    LayerNorm, Q/K/V pooling, pooled attention with relative-position bias and residual
                                         pooling, MaxPool3d skip, Conv3d stem rows (csrc/mvit.cu)

and autograd strings the backward together (each Function's backward is again one or two kernels).  torch itself is left
with glue only: the residual additions, `cat` of the cls token, DropPath row scaling and the three small relative-
position einsums (mvit_functional.rel_pos_projections).  Only what the shipped MViT YAMLs use is built -- MODE conv,
POOL_FIRST False, SEPARATE_QKV False, CLS_EMBED_ON True, USE_ABS_POS False, REL_POS_SPATIAL / TEMPORAL True, RESIDUAL_POOLING
True, DIM_MUL_IN_ATT True, NORM layernorm, no layer scale, dropout 0 -- anything else raises at construction."""
import math
import os
from functools import partial

import torch
from torch import nn

from ... import mvit_functional as MF
from ...tc_functional import tc_linear, tc_mlp
from .build import MODEL_REGISTRY
from .vit import VisionTransformer as _TimeSformerModel
from .vit import trunc_normal_

LN_EPS = 1e-6           # partial(nn.LayerNorm, eps=1e-6), slowfast_mvit/mvit.py:68-69


def round_width(width, multiplier, min_width=1, divisor=1):
    """slowfast_mvit/utils.py:7-20."""
    if not multiplier:
        return width
    width *= multiplier
    min_width = min_width or divisor
    out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


def mvit_geometry(mvit, num_frames, crop):
    """Per-block widths, heads, pooling kernels / strides and token grids from the MVIT.* keys (slowfast_mvit/mvit.py:77-228)."""
    need = dict(MODE="conv", POOL_FIRST=False, SEPARATE_QKV=False, CLS_EMBED_ON=True, USE_ABS_POS=False, REL_POS_SPATIAL=True,
                REL_POS_TEMPORAL=True, RESIDUAL_POOLING=True, DIM_MUL_IN_ATT=True, NORM="layernorm", PATCH_2D=False,
                NORM_STEM=False, USE_MEAN_POOLING=False, USE_FIXED_SINCOS_POS=False)
    for k, v in need.items():
        if k in mvit and mvit[k] != v:
            raise NotImplementedError(f"the sm_100a MViT path is built for MVIT.{k} = {v} (got {mvit[k]})")
    if mvit.get("LAYER_SCALE_INIT_VALUE", 0.0) or mvit.get("DROPOUT_RATE", 0.0):
        raise NotImplementedError("MVIT.LAYER_SCALE_INIT_VALUE / DROPOUT_RATE other than 0 are not built")
    if not mvit.get("POOL_KVQ_KERNEL") or not mvit.get("POOL_KV_STRIDE_ADAPTIVE"):
        raise NotImplementedError("MVIT.POOL_KVQ_KERNEL and MVIT.POOL_KV_STRIDE_ADAPTIVE are required (as in the shipped YAMLs)")
    depth = mvit["DEPTH"]
    pstride = list(mvit["PATCH_STRIDE"])
    grid = [num_frames // pstride[0], crop // pstride[1], crop // pstride[2]]
    dim_mul, head_mul = [1.0] * (depth + 1), [1.0] * (depth + 1)
    for i, m in mvit["DIM_MUL"]:
        dim_mul[i] = m
    for i, m in mvit["HEAD_MUL"]:
        head_mul[i] = m
    kernel = list(mvit["POOL_KVQ_KERNEL"])
    stride_q = [[] for _ in range(depth)]
    for row in mvit["POOL_Q_STRIDE"]:
        stride_q[row[0]] = list(row[1:])
    skv, stride_kv = list(mvit["POOL_KV_STRIDE_ADAPTIVE"]), []          # mvit.py:153-163
    for i in range(depth):
        if stride_q[i]:
            skv = [max(skv[d] // stride_q[i][d], 1) for d in range(3)]
        stride_kv.append(list(skv))
    blocks, dim, heads, size = [], mvit["EMBED_DIM"], mvit["NUM_HEADS"], list(grid)
    for i in range(depth):
        heads = round_width(heads, head_mul[i])
        dim_out = round_width(dim, dim_mul[i], divisor=round_width(heads, head_mul[i]))
        sq, skv_i = stride_q[i], stride_kv[i]
        q_sz = size[1] // sq[1] if sq else size[1]
        kv_sz = size[1] // skv_i[1] if skv_i else size[1]
        blocks.append(dict(dim=dim, dim_out=dim_out, heads=heads, kernel_q=kernel if sq else [], stride_q=sq,
                           kernel_kv=kernel if skv_i else [], stride_kv=skv_i, grid=list(size),
                           rel_sp=2 * max(q_sz, kv_sz) - 1, rel_t=2 * size[0] - 1))
        if sq:
            size = [s // st for s, st in zip(size, sq)]
        dim = dim_out
    return dict(patch=(list(mvit["PATCH_KERNEL"]), pstride, list(mvit["PATCH_PADDING"])), grid=grid,
                embed_dim=mvit["EMBED_DIM"], out_dim=dim, out_grid=size, blocks=blocks, mlp_ratio=mvit["MLP_RATIO"])


def _pools(kernel, stride):
    """attention.py:239-252: a pooling conv exists unless kernel and stride are all ones."""
    return bool(kernel) and not (math.prod(kernel) == 1 and math.prod(stride) == 1)


class Mlp(nn.Module):
    """slowfast_mvit/common.py:7-34 (container)."""

    def __init__(self, dim, hidden, out):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, out)


class MultiScaleAttention(nn.Module):
    """slowfast_mvit/attention.py:162-280 (container)."""

    def __init__(self, blk):
        super().__init__()
        dim, dim_out, heads = blk["dim"], blk["dim_out"], blk["heads"]
        hd = dim_out // heads
        self.num_heads, self.dim_out, self.scale = heads, dim_out, hd ** -0.5
        self.qkv = nn.Linear(dim, dim_out * 3, bias=True)
        self.proj = nn.Linear(dim_out, dim_out)
        for nm, kern, st in (("q", blk["kernel_q"], blk["stride_q"]), ("k", blk["kernel_kv"], blk["stride_kv"]),
                             ("v", blk["kernel_kv"], blk["stride_kv"])):
            if _pools(kern, st):
                setattr(self, "pool_" + nm, nn.Conv3d(hd, hd, kern, stride=st, padding=[k // 2 for k in kern], groups=hd,
                                                      bias=False))
                setattr(self, "norm_" + nm, nn.LayerNorm(hd, eps=LN_EPS))
            else:
                setattr(self, "pool_" + nm, None)
        self.rel_pos_h = nn.Parameter(torch.zeros(blk["rel_sp"], hd))
        self.rel_pos_w = nn.Parameter(torch.zeros(blk["rel_sp"], hd))
        self.rel_pos_t = nn.Parameter(torch.zeros(blk["rel_t"], hd))
        for t in (self.rel_pos_h, self.rel_pos_w, self.rel_pos_t):               # attention.py:268-276
            trunc_normal_(t, std=0.02)


class MultiScaleBlock(nn.Module):
    """slowfast_mvit/attention.py:445-543 (container)."""

    def __init__(self, blk, mlp_ratio, drop_path):
        super().__init__()
        self.dim, self.dim_out, self.drop_prob = blk["dim"], blk["dim_out"], drop_path
        self.norm1 = nn.LayerNorm(blk["dim"], eps=LN_EPS)
        self.attn = MultiScaleAttention(blk)
        self.norm2 = nn.LayerNorm(blk["dim_out"], eps=LN_EPS)
        self.mlp = Mlp(blk["dim_out"], int(blk["dim_out"] * mlp_ratio), blk["dim_out"])
        if blk["dim"] != blk["dim_out"]:
            self.proj = nn.Linear(blk["dim"], blk["dim_out"])


class PatchEmbed(nn.Module):
    """slowfast_mvit/stem_helper.py:290-322 (container)."""

    def __init__(self, dim_in, dim_out, kernel, stride, padding):
        super().__init__()
        self.proj = nn.Conv3d(dim_in, dim_out, kernel_size=kernel, stride=stride, padding=padding)


class MViT_encoder(nn.Module):
    """slowfast_mvit/mvit.py:26-406: clips [B, 3, T, H, W] -> cls feature [B, out_dim]."""
    _require_cuda = True        # CPU tests of the host logic (tests/shadow_ops.py) switch this off

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        crop = cfg.DATA.TRAIN_CROP_SIZE
        self.geo = geo = mvit_geometry(cfg.MVIT, cfg.DATA.NUM_FRAMES, crop)
        self.patch_dims = list(geo["grid"])
        kern, stride, pad = geo["patch"]
        self.patch_embed = PatchEmbed(cfg.DATA.INPUT_CHANNEL_NUM[0], geo["embed_dim"], kern, stride, pad)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, geo["embed_dim"]))
        depth = len(geo["blocks"])
        dpr = [x.item() for x in torch.linspace(0, cfg.MVIT.DROPPATH_RATE, depth)]          # mvit.py:103-105
        self.blocks = nn.ModuleList([MultiScaleBlock(b, geo["mlp_ratio"], dpr[i]) for i, b in enumerate(geo["blocks"])])
        self.norm = nn.LayerNorm(geo["out_dim"], eps=LN_EPS)
        trunc_normal_(self.cls_token, std=0.02)
        self.apply(self._init_weights)
        self.fixed_drop_scales = None        # tests replay DropPath draws: list over blocks of [B] factors (or None)

    def _init_weights(self, m):
        """mvit.py:285-292."""
        if isinstance(m, (nn.Linear, nn.Conv2d, nn.Conv3d)):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.constant_(m.bias, 0.02)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0.02)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        """mvit.py:294-315."""
        if not self.cfg.MVIT.ZERO_DECAY_POS_CLS:
            return []
        return ["rel_pos_h", "rel_pos_w", "rel_pos_hw", "rel_pos_t", "cls_token"]

    # ---------------------------------------------------------------------------------------------- the schedule
    def precision(self):
        return os.environ.get("PVRL_PRECISION") or (self.cfg.B200.PRECISION if "B200" in self.cfg else "bf16")

    def _drop(self, x, i, B):
        """DropPath (common.py:46-59): one Bernoulli(keep) draw per clip, x / keep * mask."""
        if self.fixed_drop_scales is not None:
            s = self.fixed_drop_scales[i]
        elif self.training and self.blocks[i].drop_prob > 0.0:
            keep = 1.0 - self.blocks[i].drop_prob
            s = torch.floor(keep + torch.rand(B, device=x.device)) / keep
        else:
            s = None
        return x if s is None else x * s.to(x.dtype).reshape(B, *([1] * (x.dim() - 1)))

    def _attention(self, att, blk, xn, grid, prec, act):
        """MultiScaleAttention.forward (attention.py:282-411), POOL_FIRST False."""
        heads, C = blk["heads"], blk["dim_out"] // blk["heads"]
        qkv = tc_linear(xn, att.qkv.weight, att.qkv.bias, precision=prec)                       # [B, N, 3 * heads * C]
        wq = None if att.pool_q is None else att.pool_q.weight
        wk = None if att.pool_k is None else att.pool_k.weight
        wv = None if att.pool_v is None else att.pool_v.weight
        kernel = blk["kernel_q"] or blk["kernel_kv"] or [1, 1, 1]
        q, k, v = MF.pool_qkv(qkv, wq, wk, wv, heads, C, grid, kernel, blk["stride_q"], blk["stride_kv"])
        q_grid = MF.ops.pool_out_grid(grid, kernel, blk["stride_q"], [s // 2 for s in kernel]) if wq is not None else list(grid)
        k_grid = MF.ops.pool_out_grid(grid, kernel, blk["stride_kv"], [s // 2 for s in kernel]) if wk is not None else list(grid)
        if wq is not None:
            q = MF.layer_norm(q, att.norm_q.weight, att.norm_q.bias, LN_EPS, act)
        if wk is not None:
            k = MF.layer_norm(k, att.norm_k.weight, att.norm_k.bias, LN_EPS, act)
            v = MF.layer_norm(v, att.norm_v.weight, att.norm_v.bias, LN_EPS, act)
        bq = MF.rel_pos_projections(q, q_grid, k_grid, att.rel_pos_h, att.rel_pos_w, att.rel_pos_t)
        o = MF.pooled_attention(q, k, v, bq, k_grid, att.scale, residual_pooling=True)          # [B, Nq, heads * C]
        return tc_linear(o, att.proj.weight, att.proj.bias, precision=prec), q_grid

    def _block(self, i, x, grid, prec, act):
        """MultiScaleBlock.forward (attention.py:545-567), DIM_MUL_IN_ATT: the skip is the widened *normalised* input."""
        m, blk = self.blocks[i], self.geo["blocks"][i]
        B = x.shape[0]
        xn = MF.layer_norm(x, m.norm1.weight, m.norm1.bias, LN_EPS, act)
        a, new_grid = self._attention(m.attn, blk, xn, grid, prec, act)
        skip = tc_linear(xn, m.proj.weight, m.proj.bias, precision=prec) if blk["dim"] != blk["dim_out"] else x
        sq = blk["stride_q"]
        if sq and math.prod(sq) > 1:
            skip = MF.max_pool_skip(skip, grid, sq)
        x = skip.float() + self._drop(a.float(), i, B)
        xn = MF.layer_norm(x, m.norm2.weight, m.norm2.bias, LN_EPS, act)
        h = tc_mlp(xn, m.mlp.fc1.weight, m.mlp.fc1.bias, m.mlp.fc2.weight, m.mlp.fc2.bias, precision=prec)
        return x + self._drop(h.float(), i, B), new_grid

    def forward(self, x, taps=None):
        """MViT_encoder.forward (mvit.py:338-406)."""
        if self._require_cuda and not x.is_cuda:
            raise RuntimeError("the sm_100a path needs CUDA tensors (no CPU fallback): move the model and the clips to a GPU")
        prec = self.precision()
        act = torch.float32 if prec == "bf16x3" else torch.bfloat16
        kern, stride, pad = self.geo["patch"]
        B = x.shape[0]
        rows, grid = MF.conv3d_stem_rows(x, kern, stride, pad, act)
        w = self.patch_embed.proj.weight
        w2 = w.reshape(w.shape[0], -1)
        w2 = torch.nn.functional.pad(w2, (0, rows.shape[1] - w2.shape[1]))                     # 441 -> 448 columns
        tok = tc_linear(rows, w2, self.patch_embed.proj.bias, precision=prec)                  # [B * T'H'W', D0]
        if list(grid) != self.patch_dims:
            raise ValueError(f"clip geometry {list(grid)} differs from the configured {self.patch_dims} (no absolute position "
                             "embedding to resize: build the model for this clip size)")
        x = torch.cat((self.cls_token.expand(B, -1, -1).float(), tok.float().reshape(B, -1, w.shape[0])), dim=1)
        for i in range(len(self.blocks)):
            x, grid = self._block(i, x, grid, prec, act)
            if taps is not None:
                taps.append(x)
        # the final LayerNorm is row-wise and only the cls row is read (mvit.py:399-401): normalise those B rows
        return MF.layer_norm(x[:, 0].contiguous(), self.norm.weight, self.norm.bias, LN_EPS, torch.float32)


class VisionTransformer(_TimeSformerModel):
    """reference lib/models/mvit.py:45-232: the TimeSformer wrapper (matching head, order transformer, teacher, batch
    assembly -- inherited unchanged) around `video_encoder = MViT_encoder(cfg)`."""

    def __init__(self, num_classes=1000, label_emb="", mlp=0, text_model="", num_seg=0, cfg=None, **unused):
        nn.Module.__init__(self)
        if cfg.MODEL.MODEL_NAME != "MViT":
            raise ValueError("lib/models/mvit.py serves MODEL.MODEL_NAME: MViT")
        self.cfg = cfg
        self.num_classes = num_classes
        self.temp = cfg.DEV.TEMP
        self.order_pretrain = cfg.DEV.ORDER_PRETRAIN_ENABLED
        self.order_max_len = cfg.DEV.ORDER_PRETRAIN_MAX_LEN
        self.order_fix_recognition = cfg.DEV.ORDER_FIX_RECOGNITION
        self.order_tfm_layers = cfg.DEV.ORDER_TFM_LAYERS
        self.order_recog_batch = cfg.DEV.ORDER_RECOG_BATCH
        self.depth = cfg.MVIT.DEPTH
        self.video_encoder = MViT_encoder(cfg)
        embed_dim = self.video_encoder.norm.weight.shape[0]
        self.num_features = self.embed_dim = embed_dim
        self._build_heads(cfg, embed_dim, num_classes, label_emb, mlp, text_model, num_seg)
        self._engine = None
        self._label_dev = None
        self.fixed_rand_inds = None

    def forward_features(self, x):
        return self.video_encoder(x.float())


default_cfgs = {"mvit": {"url": "https://dl.fbaipublicfiles.com/mvit/mvitv2_models/MViTv2_S_in1k.pyth", "num_classes": 1000,
                         "input_size": (3, 224, 224), "first_conv": "patch_embed.proj", "classifier": "head"}}


@MODEL_REGISTRY.register()
class MViT(nn.Module):
    """Registry wrapper, reference lib/models/mvit.py:234-266 (MODEL.MODEL_NAME: MViT)."""

    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.pretrained = cfg.MODEL.PRETRAINED
        self.model = VisionTransformer(num_classes=cfg.MODEL.NUM_CLASSES, label_emb=cfg.TRAIN.LABEL_EMB, mlp=cfg.MODEL.MLP,
                                       text_model=cfg.MODEL.TEXT_MODEL, num_seg=cfg.MODEL.NUM_SEG, cfg=cfg)
        self.attention_type = cfg.TIMESFORMER.ATTENTION_TYPE
        self.model.default_cfg = default_cfgs["mvit"]
        self.num_patches = (cfg.DATA.TRAIN_CROP_SIZE // 16) ** 2
        if self.pretrained:
            path = cfg.TIMESFORMER.PRETRAINED_MODEL
            if not path:
                raise RuntimeError("MODEL.PRETRAINED True needs TIMESFORMER.PRETRAINED_MODEL (no network here to fetch the "
                                   "MViTv2-S checkpoint the reference downloads, lib/models/mvit.py:42,259-262)")
            load_pretrained(self.model, path)

    def forward(self, x):
        return self.model(x)


def load_pretrained(model, path):
    """MViTv2 checkpoint -> `model.video_encoder` (reference helpers.py:100-145 for the mvit entry).  Keys of the released
    image checkpoint MViTv2_S_in1k.pyth live under 'model_state' without the 'video_encoder.' prefix and are converted as
    the reference does: `pool_*` and `patch_embed.proj.weight` (2-D convolutions) are repeated over the temporal kernel
    extent (`unsqueeze(2).repeat`, helpers.py:131-133), `rel_pos_*` tables are linearly interpolated to the model's length
    (helpers.py:134-138); an already converted checkpoint ('video_encoder.'-prefixed keys) loads as is.  Tensors whose
    shape still does not match (the 1000-way ImageNet head) are reported, never dropped silently."""
    ck = torch.load(path, map_location="cpu")
    for key in ("model_state", "model", "state_dict"):
        if isinstance(ck, dict) and key in ck:
            ck = ck[key]
            break
    own = model.state_dict()
    sd, skipped = {}, []
    for k, v in ck.items():
        k = k[6:] if k.startswith("model.") else k
        cand = k if k in own else "video_encoder." + k
        if cand not in own:
            skipped.append((k, "no such parameter"))
            continue
        tgt = own[cand]
        if tgt.shape != v.shape:
            if ("pool_" in k or "patch_embed.proj.weight" in k) and v.dim() == 4 and tgt.dim() == 5 \
                    and tgt.shape[:2] == v.shape[:2] and tgt.shape[3:] == v.shape[2:]:
                v = v.unsqueeze(2).repeat(1, 1, tgt.shape[2], 1, 1)                        # helpers.py:131-133
            elif "rel_pos_" in k and v.dim() == 2 and tgt.dim() == 2 and tgt.shape[1] == v.shape[1]:
                v = torch.nn.functional.interpolate(v.t().unsqueeze(0).float(), size=tgt.shape[0], mode="linear")[0] \
                    .t().contiguous().to(v.dtype)                                          # helpers.py:134-138
            else:
                skipped.append((k, f"shape {tuple(v.shape)} != {tuple(tgt.shape)}"))
                continue
        sd[cand] = v
    if skipped:
        print("load_pretrained: skipped " + ", ".join(f"{k} ({why})" for k, why in skipped), flush=True)
    res = model.load_state_dict(sd, strict=False)
    model.pretrained_skipped = skipped
    return res
