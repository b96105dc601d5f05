"""Clip-level order / diffusion transformer (reference lib/models/tfm_model.py:70-289), SURVEY.md 8f-1.

Kept in PyTorch for now (4 layers x 9 tokens x 512: < 0.1 % of the step's FLOPs) but re-expressed without the
reference's host synchronisations: mask positions, pad starts and noise are drawn on the device and applied
with masks instead of `.item()` loops (tfm_model.py:279-287) and `.cpu()` lookups (diffusion_model.py:346),
so the whole pretrain step stays asynchronous.  Parameter names match the reference state_dict
(`order_tfm.pad_embedding.weight`, `...temporalModelling.resblocks.i.attn.in_proj_weight`, `...time_mlp.1.weight`)."""
import math
from collections import OrderedDict

import torch
from torch import nn


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    """tfm_model.py:32-53."""

    def __init__(self, d_model, n_head, dropout=0.0):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head, dropout=dropout)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)

    def forward(self, x, pad_mask=None):
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False, key_padding_mask=pad_mask)[0]
        return x + self.mlp(self.ln_2(x))


class TemporalModelling(nn.Module):
    def __init__(self, width, layers, heads, dropout=0.0):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, dropout) for _ in range(layers)])

    def forward(self, x, pad_mask=None):
        for blk in self.resblocks:
            x = blk(x, pad_mask=pad_mask)
        return x


class SinusoidalPositionEmbeddings(nn.Module):
    """diffusion_model.py:34-46."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, t):
        half = self.dim // 2
        e = math.log(10000) / (half - 1)
        e = torch.exp(torch.arange(half, device=t.device) * -e)
        e = t[:, None].float() * e[None, :]
        return torch.cat((e.sin(), e.cos()), dim=-1)


class DiffusionTransformer(nn.Module):
    def __init__(self, num_seg=8, tfm_layers=4, tfm_heads=8, hidden_size=512, dropout=0.0, cfg=None):
        super().__init__()
        self.cfg = cfg
        self.hidden_size, self.num_seg, self.tfm_layers, self.tfm_heads = hidden_size, num_seg, tfm_layers, tfm_heads
        self.max_len = cfg.DEV.ORDER_PRETRAIN_MAX_LEN
        self.pad_embedding = nn.Embedding(1, hidden_size)
        self.type_embedding = nn.Embedding(2, hidden_size)
        self.temporalEmbedding = nn.Embedding(self.max_len, hidden_size)
        self.temporalModelling = TemporalModelling(hidden_size, tfm_layers, tfm_heads, 0.0)
        self.time_mlp = nn.Sequential(SinusoidalPositionEmbeddings(hidden_size // 4),
                                      nn.Linear(hidden_size // 4, hidden_size), nn.GELU(),
                                      nn.Linear(hidden_size, hidden_size))
        self.initialize_parameters()
        self.total_levels = self.level_batch = tfm_layers
        # linear beta schedule (diffusion_model.py:328-331) and q(x_t | x_0) coefficients (tfm_model.py:106-127)
        betas = torch.linspace(0.0001, 0.02, self.total_levels)
        ac = torch.cumprod(1.0 - betas, dim=0)
        self.register_buffer("sqrt_alphas_cumprod", torch.sqrt(ac), persistent=False)
        self.register_buffer("sqrt_one_minus_alphas_cumprod", torch.sqrt(1.0 - ac), persistent=False)
        self.fixed_draws = None      # tests inject (mask_inds, pad_start, noise[levels,B,C]) to replay the reference's draws

    def initialize_parameters(self):
        """tfm_model.py:251-263."""
        nn.init.normal_(self.pad_embedding.weight, std=0.01)
        nn.init.normal_(self.temporalEmbedding.weight, std=0.01)
        w, l = self.temporalModelling.width, self.temporalModelling.layers
        proj_std, attn_std, fc_std = (w ** -0.5) * ((2 * l) ** -0.5), w ** -0.5, (2 * w) ** -0.5
        for blk in self.temporalModelling.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(blk.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(blk.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(blk.mlp.c_proj.weight, std=proj_std)

    # one denoising level: tfm_model.py:187-196 / :229-237
    def _level(self, feats, is_mask, t_index, pad_mask):
        S, B, C = feats.shape
        dev = feats.device
        type_emb = torch.where(is_mask.unsqueeze(-1), self.type_embedding.weight[1], self.type_embedding.weight[0])
        pos = self.temporalEmbedding.weight[:S].unsqueeze(1)
        t = torch.full((B,), t_index, device=dev, dtype=torch.long)
        h = feats + type_emb + pos + self.time_mlp(t).unsqueeze(0)
        h = self.temporalModelling(h, pad_mask=pad_mask)
        return (h * is_mask.unsqueeze(-1)).sum(0)          # the token at the mask position of every sample

    def forward(self, x, is_pretrain=False):
        if self.training and is_pretrain:
            return self._pretrain(x)
        return self._forecast(x)

    def _pretrain(self, x):
        """tfm_model.py:137-156,165-204: mask one clip per video, pad a random tail, denoise over the levels."""
        S, C = self.max_len, x.shape[1]
        feats = x.reshape(-1, S, C).transpose(0, 1)                       # '(b t) c -> t b c'
        B, dev = feats.shape[1], x.device
        if self.fixed_draws is not None:
            mask_inds, pad_start, noise = (t.to(dev) for t in self.fixed_draws)
        else:
            mask_inds = torch.randint(0, S, (B,), device=dev)
            # pad_start ~ U{mask+1 .. S-1}, or S (no padding) when the mask is the last token (tfm_model.py:279-284)
            span = (S - 1 - mask_inds).clamp(min=1)
            pad_start = mask_inds + 1 + (torch.rand(B, device=dev) * span).long().clamp(max=S)
            pad_start = torch.where(mask_inds + 1 == S, torch.full_like(mask_inds, S), pad_start.clamp(max=S - 1))
            noise = torch.randn(self.tfm_layers, B, C, device=dev)
        pos = torch.arange(S, device=dev).unsqueeze(1)
        is_mask = pos == mask_inds.unsqueeze(0)                           # [S, B]
        pad = pos >= pad_start.unsqueeze(0)                               # [S, B]
        x0 = (feats * is_mask.unsqueeze(-1)).sum(0)                       # clip embeddings that get masked out
        feats = torch.where(pad.unsqueeze(-1), self.pad_embedding.weight[0], feats)
        pad_mask = pad.t()
        outs, den = [], None
        for lvl in range(self.tfm_layers):
            t_index = self.total_levels - 1 - lvl
            src = (x0 if lvl == 0 else den).detach()
            noisy = self.sqrt_alphas_cumprod[t_index] * src + self.sqrt_one_minus_alphas_cumprod[t_index] * noise[lvl]
            lvl_feats = torch.where(is_mask.unsqueeze(-1), noisy.unsqueeze(0), feats)
            den = self._level(lvl_feats, is_mask, t_index, pad_mask)
            outs.append(den)
        x0_target = x0.unsqueeze(0).expand(self.total_levels, -1, -1).reshape(-1, C)
        inter = torch.cat(outs)
        return den, mask_inds, [x0_target, inter], inter

    def _forecast(self, x):
        """tfm_model.py:206-249: append an (all-zero noise) token and denoise it over the levels."""
        S, C = self.max_len, x.shape[1]
        feats0 = x.reshape(-1, self.num_seg, C).transpose(0, 1)
        B, dev = feats0.shape[1], x.device
        orig = torch.cat((feats0, torch.zeros(S - self.num_seg, B, C, device=dev, dtype=x.dtype)), dim=0)
        is_mask = (torch.arange(S, device=dev) == S - 1).unsqueeze(1).expand(S, B)
        feats, den = orig, None
        for lvl in range(self.tfm_layers):
            t_index = self.total_levels - 1 - lvl
            if lvl != 0:
                noisy = self.sqrt_alphas_cumprod[t_index] * den.detach()
                feats = torch.where(is_mask.unsqueeze(-1), noisy.unsqueeze(0), feats)
            den = self._level(feats, is_mask, t_index, None)
            feats = torch.where(is_mask.unsqueeze(-1), den.unsqueeze(0), orig)
        return den
