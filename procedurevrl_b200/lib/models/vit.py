"""Drop-in mirror of the reference's TimeSformer + matching head (reference lib/models/vit.py), running on the
sm_100a kernels of libpvrl_sm100.so.

Same surface as the reference: registry entry `vit_base_patch16_224_develop(cfg)` whose
`.model` is a `VisionTransformer` with the reference's attribute names (`blocks`, `pos_drop`, `head`,
`order_tfm`, `text_model`, ...), identical `state_dict()` schema, identical `forward` signatures and return
values.  The `nn.Linear / nn.LayerNorm / nn.Conv2d` sub-modules below are *parameter containers only* (they
give the reference's names, shapes and default initialisers); their own forward is never used -- the math is
`EncoderEngine` (forward_features, vit.py:365-423) plus the head / similarity kernels (vit.py:300-322)."""
import os
from functools import partial

import torch
from torch import nn

from ... import functional as PF
from ...engine import EncoderEngine, encode
from .build import MODEL_REGISTRY
from .order_tfm import DiffusionTransformer as OrderTransformer


def trunc_normal_(t, std=0.02):
    """vit_utils.py:59-77 (truncated at +-2 std around 0)."""
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class Mlp(nn.Module):
    """vit.py:44-60 (container)."""

    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)
        self.drop = nn.Dropout(0.0)


class Attention(nn.Module):
    """vit.py:62-92 (container)."""

    def __init__(self, dim, num_heads, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.with_qkv = True
        self.qkv, self.proj = nn.Linear(dim, dim * 3, bias=qkv_bias), nn.Linear(dim, dim)
        self.proj_drop, self.attn_drop = nn.Dropout(0.0), nn.Dropout(0.0)


class DropPath(nn.Module):
    """vit_utils.py:157-165 (rate holder; the mask is drawn by VisionTransformer._drop_scales)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob


class Block(nn.Module):
    """vit.py:94-158 (container)."""

    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, drop_path, norm_layer, attention_type):
        super().__init__()
        self.attention_type = attention_type
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads, qkv_bias)
        if attention_type == "divided_space_time":
            self.temporal_norm1 = norm_layer(dim)
            self.temporal_attn = Attention(dim, num_heads, qkv_bias)
            self.temporal_fc = nn.Linear(dim, dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchEmbed(nn.Module):
    """vit.py:160-180 (container)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class PreExtractedTextTower(nn.Module):
    """Stand-in for the frozen CLIP ViT-B/16 text tower (vit.py:257-261) when the `clip` package is absent.
    The hot path feeds pre-extracted text embeddings (meta['clip_text_emb'], north star); asking this tower to
    encode token ids is an error rather than a silent approximation."""

    def encode_text(self, ids):
        raise RuntimeError("CLIP text tower is not available: pass pre-extracted embeddings as meta['clip_text_emb']")


class VisionTransformer(nn.Module):
    _require_cuda = True        # CPU tests of the host logic (tests/shadow_ops.py) switch this off

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4.0, qkv_bias=False, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.1,
                 norm_layer=nn.LayerNorm, num_frames=8, attention_type="divided_space_time", label_emb="", mlp=0,
                 text_model="", lp=False, num_seg=0, extra_tr="order", drope=0.0, cfg=None):
        super().__init__()
        if attention_type not in ("divided_space_time", "space_only", "joint_space_time"):
            raise ValueError(f"TIMESFORMER.ATTENTION_TYPE={attention_type}")
        self.cfg = cfg
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_heads, self.mlp_ratio, self.patch_size = num_heads, mlp_ratio, patch_size
        self.temp = cfg.DEV.TEMP
        self.order_pretrain = cfg.DEV.ORDER_PRETRAIN_ENABLED
        self.order_max_len = cfg.DEV.ORDER_PRETRAIN_MAX_LEN
        self.order_fix_recognition = cfg.DEV.ORDER_FIX_RECOGNITION
        self.order_tfm_layers = cfg.DEV.ORDER_TFM_LAYERS
        self.order_recog_batch = cfg.DEV.ORDER_RECOG_BATCH
        self.attention_type, self.depth, self.num_frames = attention_type, depth, num_frames

        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        if attention_type != "space_only":                                             # vit.py:213-215
            self.time_embed = nn.Parameter(torch.zeros(1, num_frames, embed_dim))
            self.time_drop = nn.Dropout(p=drop_rate)
        self.dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]        # vit.py:220
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, qkv_bias, self.dpr[i], norm_layer,
                                           attention_type) for i in range(depth)])
        self.norm = norm_layer(embed_dim)

        self._build_heads(cfg, embed_dim, num_classes, label_emb, mlp, text_model, num_seg)

        trunc_normal_(self.pos_embed, std=0.02)
        trunc_normal_(self.cls_token, std=0.02)
        # vit.py:273-281 zero-initialises temporal_fc in *every* block (the `i > 0` guard sees the ModuleList first)
        if attention_type == "divided_space_time":
            for blk in self.blocks:
                nn.init.constant_(blk.temporal_fc.weight, 0)
                nn.init.constant_(blk.temporal_fc.bias, 0)

        self._engine = None
        self._label_dev = None
        self.fixed_rand_inds = None           # tests replay the reference's randperm (vit.py:345)
        self.fixed_drop_scales = None         # tests replay DropPath draws

    def _build_heads(self, cfg, embed_dim, num_classes, label_emb, mlp, text_model, num_seg):
        """Matching head, order transformer, fine-tune heads and the text tower (vit.py:229-266; lib/models/mvit.py:72-108
        builds exactly the same set around the MViT encoder)."""
        self.mlp, self.label = mlp, label_emb
        if label_emb != "":                                                            # pre-training (vit.py:231-236)
            self.label_emb = torch.load(label_emb)
            self.head = nn.Linear(embed_dim, self.label_emb.shape[1])
            self.order_tfm = OrderTransformer(num_seg=self.order_max_len - 1, tfm_layers=self.order_tfm_layers,
                                              dropout=cfg.MODEL.DROP_E, hidden_size=self.head.weight.shape[0], cfg=cfg)
        else:                                                                          # fine-tuning (vit.py:237-253)
            if cfg.DEV.MATCH_LANG_EMB:
                self.label_emb = torch.load(cfg.DEV.TEST_LANG_EMB)
                self.head = nn.Linear(embed_dim, self.label_emb.shape[1])
                for p in self.head.parameters():
                    p.requires_grad = False
            else:
                self.label_emb = False
                self.test_lang_emb = torch.load(cfg.DEV.TEST_LANG_EMB)
                self.head = nn.Linear(embed_dim, self.test_lang_emb.shape[1])
                for p in self.head.parameters():
                    p.requires_grad = False
                if cfg.TRAIN.DATASET == "Epickitchens":
                    self.head_n = nn.Linear(self.test_lang_emb.shape[1], 300)
                    self.head_v = nn.Linear(self.test_lang_emb.shape[1], 97)
                else:
                    self.head_cls = nn.Linear(self.test_lang_emb.shape[1], num_classes)
            self.apply(self._init_weights)

        self.text = text_model
        if text_model == "clip_vit_b_16":                                              # vit.py:256-261
            try:
                import clip  # noqa: WPS433
                clip_model, _ = clip.load("ViT-B/16", jit=False)
                del clip_model.visual
                self.text_model = clip_model.float()
            except ImportError:
                self.text_model = PreExtractedTextTower()
            for p in self.text_model.parameters():
                p.requires_grad = False

        if num_seg > 0:                                                                # vit.py:264-266
            self.num_seg = num_seg
            self.order_tfm = OrderTransformer(num_seg=num_seg, tfm_layers=self.order_tfm_layers, dropout=cfg.MODEL.DROP_E,
                                              hidden_size=self.head.weight.shape[0], cfg=cfg)

    def _init_weights(self, m):
        """vit.py:442-449."""
        if isinstance(m, nn.Linear):
            trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {"pos_embed", "cls_token", "time_embed"}

    def get_classifier(self):
        return self.head

    # ------------------------------------------------------------------------------------------------ engine
    def engine(self):
        if self._engine is None:
            names = dict(self.named_parameters())
            if self._require_cuda and not names["cls_token"].is_cuda:
                raise RuntimeError("the sm_100a path needs the model on a CUDA device (no CPU fallback): call .cuda()")
            prec = os.environ.get("PVRL_PRECISION") or (self.cfg.B200.PRECISION if "B200" in self.cfg else "bf16")
            self._engine = EncoderEngine(names, self.depth, self.num_frames, self.embed_dim, self.num_heads,
                                         self.mlp_ratio, self.patch_size, eps=self.norm.eps,
                                         attention_type=self.attention_type, precision=prec, prefix="")
        return self._engine

    def _apply(self, fn, *a, **k):
        self._engine = None                   # parameters may move / change dtype
        return super()._apply(fn, *a, **k)

    def _drop_scales(self, Bc, T, HW, dev):
        """DropPath factors mask/keep per block (vit_utils.py:140-155): one row per sequence of the current view --
        (b h w) in the temporal branch, (b t) in the spatial branch, b for the MLP (vit.py:132,144,157)."""
        if self.fixed_drop_scales is not None:
            return self.fixed_drop_scales
        if not self.training or max(self.dpr) == 0.0:
            return None
        # one draw for all blocks (4 tiny kernels per step instead of 4 per block); block i's slice = its rows in view order
        plain = self.attention_type != "divided_space_time"
        if plain:       # plain blocks: one row per sequence of the view (vit.py:125-126)
            n = Bc * T if self.attention_type == "space_only" else Bc
            parts = (("attn", n), ("mlp", n))
        else:
            parts = (("temporal", Bc * HW), ("spatial", Bc * T), ("mlp", Bc))
        per_block = sum(n for _, n in parts)
        live = [i for i, r in enumerate(self.dpr) if r != 0.0]
        key = (per_block, len(live), str(dev))
        if getattr(self, "_keep_key", None) != key:
            self._keep_vec = torch.tensor([1.0 - self.dpr[i] for i in live], device=dev).repeat_interleave(per_block)
            self._keep_key = key
        s = torch.floor(self._keep_vec + torch.rand(per_block * len(live), device=dev)) / self._keep_vec
        out, off = [None] * len(self.dpr), 0
        for i in live:
            d = {}
            for name, n in parts:
                d[name] = s[off:off + n]
                off += n
            out[i] = d
        return out

    def forward_features(self, x):
        """vit.py:365-423: frames [Bc, 3, T, H, W] -> cls feature [Bc, D]."""
        Bc, _, T, H, W = x.shape
        HW = (H // self.patch_size) * (W // self.patch_size)
        # uint8 frames go to the kernels as they are (normalisation with DATA.MEAN / DATA.STD is fused into the im2col)
        eng = self.engine()
        if "DATA" in self.cfg:
            eng.pixel_mean, eng.pixel_std = tuple(self.cfg.DATA.MEAN), tuple(self.cfg.DATA.STD)
        return encode(eng, x if x.dtype == torch.uint8 else x.float(), self._drop_scales(Bc, T, HW, x.device))

    def check_device_norm(self, label_emb, device, norm=False):
        """vit.py:435-440 with its GPU semantics: rows are L2-normalised when the bank first reaches the device."""
        if self._label_dev != device:
            label_emb = label_emb.to(device).float()
            if norm:
                label_emb = label_emb / label_emb.norm(dim=1, keepdim=True)
            self._label_dev = device
        return label_emb.contiguous()

    # ------------------------------------------------------------------------------------------------ forward
    def forward(self, x):
        """vit.py:283-358."""
        text = None
        if len(self.text) > 0 and self.training:
            x, text = x
        batch_size = x.shape[0]
        if self.order_pretrain:                                                        # vit.py:290-291
            x = x.reshape(batch_size * self.order_max_len, *x.shape[2:])
        elif getattr(self, "num_seg", 0) > 0:                                          # vit.py:292-293
            b, c, mt, h, w = x.shape
            x = x.reshape(b, c, self.num_seg, mt // self.num_seg, h, w).permute(0, 2, 1, 3, 4, 5)
            x = x.reshape(b * self.num_seg, c, mt // self.num_seg, h, w)
        x = self.forward_features(x.contiguous())
        num_seg = getattr(self, "num_seg", 0)

        if self.cfg.DEV.MATCH_LANG_EMB:                                                # vit.py:299-307
            self.label_emb = self.check_device_norm(self.label_emb, x.device, norm=True)
            x = PF.l2_normalize(PF.linear_small(x, self.head.weight, self.head.bias))
            video_emb = x
            if num_seg > 0:
                x = PF.l2_normalize(self.order_tfm(video_emb))
            x = PF.similarity_logits(x, self.label_emb, self.temp)
        else:                                                                          # vit.py:308-322
            x = PF.linear_small(x, self.head.weight, self.head.bias)
            if num_seg > 0:
                video_emb = PF.l2_normalize(x)
                x = self.order_tfm(video_emb)
                x = PF.linear_small(x, self.head_cls.weight, self.head_cls.bias)
            else:
                x = PF.l2_normalize(x)
                if hasattr(self, "head_n"):
                    v = PF.linear_small(x, self.head_v.weight, self.head_v.bias) / self.temp
                    n = PF.linear_small(x, self.head_n.weight, self.head_n.bias) / self.temp
                    return (v, n)
                x = PF.linear_small(x, self.head_cls.weight, self.head_cls.bias) / self.temp

        if isinstance(self.label_emb, torch.Tensor) and len(self.text) > 0 and self.training:   # vit.py:325-352
            teacher_x = self.get_pseudo_labels(x.device, text)
            pred_video_emb, mask_inds, mse_loss, intermediate = self.order_tfm(video_emb, is_pretrain=True)
            masked_teacher_x = self.get_mask_samples(teacher_x, mask_inds)
            inter_pred = PF.similarity_logits(PF.l2_normalize(intermediate), self.label_emb, self.temp)
            inter_teacher = masked_teacher_x.unsqueeze(0).expand(self.order_tfm.level_batch, -1, -1) \
                .reshape(-1, masked_teacher_x.size(-1))
            if self.fixed_rand_inds is not None:
                rand_inds = self.fixed_rand_inds.to(x.device)[:batch_size * self.order_recog_batch]
            else:
                # uniform random permutation (vit.py:345) drawn as argsort of uniforms: sync-free and capturable
                rand_inds = torch.rand(x.shape[0], device=x.device).argsort()[:batch_size * self.order_recog_batch]
            x = torch.cat((x[rand_inds], inter_pred), dim=0)
            teacher_x = torch.cat((teacher_x[rand_inds], inter_teacher), dim=0)
            return x, teacher_x, mse_loss

        if not self.training:                                                          # vit.py:355-356
            x = PF.softmax_rows(x)
        return x

    def get_mask_samples(self, all_samples, mask_inds):
        """vit.py:360-363."""
        s = all_samples.reshape(-1, self.order_max_len, all_samples.shape[-1])
        return s[torch.arange(s.shape[0], device=s.device), mask_inds, :]

    @torch.no_grad()
    def get_pseudo_labels(self, device, text):
        """vit.py:425-433: teacher logits from (text embedding + CLIP visual feature) / 2.  The text embedding is
        taken pre-extracted from meta['clip_text_emb'] when present, else produced by the frozen text tower."""
        if "clip_text_emb" in text:
            text_emb = text["clip_text_emb"].to(device).float()
        else:
            text_emb = self.text_model.encode_text(text["clip_text_ids"].to(device)).float()
        text_emb = text_emb.reshape(-1, text_emb.shape[-1])
        vis = text["clip_vis_feat"].to(device).float().reshape(-1, text_emb.shape[-1])
        e = PF.l2_normalize((text_emb + vis) / 2)
        return PF.similarity_logits(e, self.label_emb, self.temp)


default_cfgs = {"vit_base_patch16_224": {"num_classes": 1000, "input_size": (3, 224, 224), "first_conv": "patch_embed.proj",
                                         "classifier": "head", "mean": (0.5, 0.5, 0.5), "std": (0.5, 0.5, 0.5)}}


@MODEL_REGISTRY.register()
class vit_base_patch16_224_develop(nn.Module):
    """Registry wrapper, vit.py:473-506 (MODEL.MODEL_NAME: vit_base_patch16_224_develop)."""

    def __init__(self, cfg, **kwargs):
        super().__init__()
        self.pretrained = cfg.MODEL.PRETRAINED
        self.model = VisionTransformer(
            img_size=cfg.DATA.TRAIN_CROP_SIZE, num_classes=cfg.MODEL.NUM_CLASSES, patch_size=16, embed_dim=768,
            depth=cfg.TIMESFORMER.DEPTH, num_heads=12, mlp_ratio=4, qkv_bias=True,
            norm_layer=partial(nn.LayerNorm, eps=1e-6), drop_rate=0.0, attn_drop_rate=0.0,
            drop_path_rate=cfg.MODEL.DROP_PATH, num_frames=cfg.DATA.NUM_FRAMES,
            attention_type=cfg.TIMESFORMER.ATTENTION_TYPE, label_emb=cfg.TRAIN.LABEL_EMB, mlp=cfg.MODEL.MLP,
            text_model=cfg.MODEL.TEXT_MODEL, lp=cfg.MODEL.TEXT_LP, num_seg=cfg.MODEL.NUM_SEG,
            extra_tr=cfg.MODEL.EXTRA_TR, drope=cfg.MODEL.DROP_E, cfg=cfg, **kwargs)
        self.attention_type = cfg.TIMESFORMER.ATTENTION_TYPE
        self.model.default_cfg = default_cfgs["vit_base_patch16_224"]
        self.num_patches = (cfg.DATA.TRAIN_CROP_SIZE // 16) ** 2
        if self.pretrained:
            path = cfg.TIMESFORMER.PRETRAINED_MODEL
            if not path:
                raise RuntimeError("MODEL.PRETRAINED True needs TIMESFORMER.PRETRAINED_MODEL (no network here to fetch "
                                   "the timm ViT-B/16 checkpoint the reference downloads, helpers.py:108-115)")
            load_pretrained(self.model, path, cfg.DATA.NUM_FRAMES)

    def forward(self, x):
        return self.model(x)


def load_pretrained(model, path, num_frames):
    """Checkpoint -> TimeSformer initialisation with the key conventions of reference helpers.py:24-52,196-243:
    accepts {'model_state': ...} / {'model': ...}, strips a leading 'model.', nearest-resizes `pos_embed`
    (helpers.py:199-211) and `time_embed` (:213-218) when their lengths differ, copies spatial attention / norm1 into
    the temporal branch when the checkpoint has none (:220-237), drops a classifier of another width (:185-192) and
    loads non-strictly.  Anything else whose shape does not match is reported, never dropped silently."""
    ck = torch.load(path, map_location="cpu")
    for key in ("model_state", "model", "state_dict"):
        if isinstance(ck, dict) and key in ck:
            ck = ck[key]
            break
    sd = {(k[6:] if k.startswith("model.") else k): v for k, v in ck.items()}
    own = model.state_dict()
    interp = torch.nn.functional.interpolate
    if "pos_embed" in sd and "pos_embed" in own and sd["pos_embed"].shape[1] != own["pos_embed"].shape[1]:
        pe = sd["pos_embed"]
        grid = interp(pe[:, 1:].transpose(1, 2), size=own["pos_embed"].shape[1] - 1, mode="nearest").transpose(1, 2)
        sd["pos_embed"] = torch.cat((pe[:, :1], grid), 1)
    if "time_embed" in sd and sd["time_embed"].shape[1] != num_frames:
        sd["time_embed"] = interp(sd["time_embed"].transpose(1, 2), size=num_frames, mode="nearest").transpose(1, 2)
    if getattr(model, "attention_type", "divided_space_time") == "divided_space_time":
        for k in list(sd):
            if "blocks" in k and ".attn." in k and k.replace(".attn.", ".temporal_attn.") not in sd:
                sd[k.replace(".attn.", ".temporal_attn.")] = sd[k]
            if "blocks" in k and ".norm1." in k and k.replace(".norm1.", ".temporal_norm1.") not in sd:
                sd[k.replace(".norm1.", ".temporal_norm1.")] = sd[k]
    skipped = [(k, f"shape {tuple(v.shape)} != {tuple(own[k].shape)}") for k, v in sd.items()
               if k in own and own[k].shape != v.shape]
    if skipped:
        print("load_pretrained: skipped " + ", ".join(f"{k} ({why})" for k, why in skipped), flush=True)
    sd = {k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}
    model.pretrained_skipped = skipped
    return model.load_state_dict(sd, strict=False)
