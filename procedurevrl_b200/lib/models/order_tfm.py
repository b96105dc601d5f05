"""Clip-level order / diffusion transformer (reference lib/models/tfm_model.py:70-289), SURVEY.md 8f-1.

The pre-training path (`_pretrain`) runs on the `pvrl_ot_*` kernels through procedurevrl_b200.order_engine (5 launches
per block forward, 11 backward, instead of ~1500 eager kernels per step) and has no host synchronisation: mask
positions, pad starts and noise are drawn on the device instead of the reference's `.item()` loops
(tfm_model.py:279-287) and `.cpu()` lookups (diffusion_model.py:346).  The forecasting path (`_forecast`, fine-tuning /
evaluation only) is still expressed with torch modules.  Parameter names match the reference state_dict
(`order_tfm.pad_embedding.weight`, `...temporalModelling.resblocks.i.attn.in_proj_weight`, `...time_mlp.1.weight`)."""
import math
from collections import OrderedDict

import torch
from torch import nn

from ...order_engine import order_levels


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
        self._sqrt_ac, self._sqrt_1mac = torch.sqrt(ac).tolist(), torch.sqrt(1.0 - ac).tolist()   # host copies: no sync
        self.grad_into_params = False   # trainer.PretrainStep: block gradients accumulate straight into the flat .grad views
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
        B, dev = x.shape[0] // S, x.device                                # rows of x are '(b t)': token (b, t) = row b*S + t
        x = x.float().contiguous()
        if self.fixed_draws is not None:
            mask_inds, pad_start, noise = (t.to(dev) for t in self.fixed_draws)
        else:
            mask_inds = torch.randint(0, S, (B,), device=dev)
            # pad_start ~ U{mask+1 .. S-1}, or S (no padding) when the mask is the last token (tfm_model.py:279-284)
            span = (S - 1 - mask_inds).clamp(min=1)
            pad_start = mask_inds + 1 + (torch.rand(B, device=dev) * span).long().clamp(max=S)
            pad_start = torch.where(mask_inds + 1 == S, torch.full_like(mask_inds, S), pad_start.clamp(max=S - 1))
            noise = torch.randn(self.tfm_layers, B, C, device=dev)
        mask_inds, pad_start = mask_inds.long().contiguous(), pad_start.long().contiguous()
        L = self.tfm_layers
        rows = torch.arange(B, device=dev) * S + mask_inds
        x0 = x.index_select(0, rows)                                      # clip embeddings that get masked out
        # diffusion-time embeddings of all levels in one batched call (level lvl runs at t = L - 1 - lvl)
        tvecs = self.time_mlp(torch.arange(L - 1, -1, -1, device=dev))
        coef = [(float(self._sqrt_ac[L - 1 - lvl]), float(self._sqrt_1mac[L - 1 - lvl])) for lvl in range(L)]
        inter = order_levels(self.temporalModelling.resblocks, x, tvecs, self.type_embedding.weight,
                             self.temporalEmbedding.weight[:S], self.pad_embedding.weight, x0.detach(),
                             noise.float().contiguous(), mask_inds, pad_start, B, S, self.tfm_heads, coef,
                             eps=self.temporalModelling.resblocks[0].ln_1.eps, grad_into_params=self.grad_into_params)
        den = inter[(L - 1) * B:]
        x0_target = x0.unsqueeze(0).expand(self.total_levels, -1, -1).reshape(-1, C)
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
