"""Explicit forward / backward schedule of the clip-order transformer's pre-training levels on the `pvrl_ot_*`
kernels (include/pvrl.h) -- the B200-native replacement for the ~1500 eager kernels per step of
DiffusionTransformer.diffusion_signal_training (reference lib/models/tfm_model.py:165-204) and its autograd.

Per denoising level: one level-input kernel, then per ResidualAttentionBlock (tfm_model.py:32-53) five launches
(LN+in_proj, attention, out_proj+residual, LN+c_fc (+QuickGELU of its output), c_proj+residual).  Levels are independent in the
backward because each level's noisy input is built from the *detached* output of the previous level
(tfm_model.py:183-186) and they share the block weights, so the backward stacks the levels along the row axis and runs
once: eleven launches per block over L*B*S rows.  torch supplies memory and the autograd hook-up only."""
import torch

from . import ops

BLOCK_PARAMS = ("ln_1.weight", "ln_1.bias", "attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight",
                "attn.out_proj.bias", "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias",
                "mlp.c_proj.weight", "mlp.c_proj.bias")


class OrderLevels(torch.autograd.Function):
    """inter[L*B, C] = denoised mask-position tokens of every level.
    Differentiable inputs: video [B*S, C], tvecs [L, C] (time_mlp outputs), type_w [2, C], pos_w [S, C], pad_w [1, C]
    and the 12 parameters of every block; x0 / noise / mask_inds / pad_start / coefficients are constants."""

    @staticmethod
    def forward(ctx, cfg, video, tvecs, type_w, pos_w, pad_w, x0, noise, mask_inds, pad_start, *params):
        B, S, H, eps, coef = cfg["B"], cfg["S"], cfg["heads"], cfg["eps"], cfg["coef"]
        L, C = tvecs.shape
        M = B * S
        nblk = len(params) // len(BLOCK_PARAMS)
        dev = video.device
        f32 = dict(device=dev, dtype=torch.float32)
        video = video.contiguous()
        tvecs, type_w, pos_w, pad_w = (t.detach().contiguous() for t in (tvecs, type_w, pos_w, pad_w))
        P = [p.detach() for p in params]
        mask_rows = torch.arange(B, device=dev) * S + mask_inds
        # activations saved for the backward live in [L*M, ...] buffers, level lvl in rows [lvl*M, (lvl+1)*M): the levels
        # are independent in the backward (detached inputs) and share the block weights, so the backward runs ONCE over
        # all L*M rows -- a quarter of the launches, and the weight gradients sum over the levels inside the dW kernels
        Hd = P[8].shape[0]
        saved = [dict(xhat1=torch.empty(L * M, C, **f32), rstd1=torch.empty(L * M, **f32),
                      qkv=torch.empty(L * M, 3 * C, **f32), probs=torch.empty(L * B, H, S, S, **f32),
                      o=torch.empty(L * M, C, **f32), xhat2=torch.empty(L * M, C, **f32),
                      rstd2=torch.empty(L * M, **f32), u=torch.empty(L * M, Hd, **f32),
                      act=torch.empty(L * M, Hd, **f32)) for _ in range(nblk)]
        outs = []
        src = x0.contiguous()
        for lvl in range(L):
            ca, cb = coef[lvl]
            r0, r1 = lvl * M, (lvl + 1) * M
            h = torch.empty(M, C, **f32)
            ops.ot_embed_fwd(video, src, noise[lvl].contiguous(), ca, cb, mask_inds, pad_start, type_w, pos_w, pad_w,
                             tvecs[lvl], h, B, S)
            for i in range(nblk):
                (ln1w, ln1b, win, bin_, wout, bout, ln2w, ln2b, wfc, bfc, wproj, bproj) = P[12 * i:12 * i + 12]
                sv = saved[i]
                qkv, o, u, act = sv["qkv"][r0:r1], sv["o"][r0:r1], sv["u"][r0:r1], sv["act"][r0:r1]
                ops.ot_linear_fwd(h, win, bin_, qkv, ops.OT_X_LN, ln1w, ln1b, eps, sv["xhat1"][r0:r1], sv["rstd1"][r0:r1])
                ops.ot_attn_fwd(qkv, pad_start, sv["probs"][lvl * B:(lvl + 1) * B], o, B, S, H)
                h_mid = torch.empty(M, C, **f32)
                ops.ot_linear_fwd(o, wout, bout, h_mid, resid=h)
                ops.ot_linear_fwd(h_mid, wfc, bfc, u, ops.OT_X_LN, ln2w, ln2b, eps, sv["xhat2"][r0:r1], sv["rstd2"][r0:r1],
                                  act_out=act)                 # u for QuickGELU' in the backward, act = QuickGELU(u)
                h = torch.empty(M, C, **f32)
                ops.ot_linear_fwd(act, wproj, bproj, h, resid=h_mid)
            den = h.index_select(0, mask_rows)          # tokens at the mask positions (tfm_model.py:194)
            outs.append(den)
            src = den
        ctx.cfg, ctx.saved, ctx.P = cfg, saved, P
        ctx.consts = (mask_inds, pad_start, mask_rows)
        ctx.shapes = (M, C, L, nblk, [p.shape for p in params])
        return torch.cat(outs)

    @staticmethod
    def backward(ctx, d_inter):
        cfg, saved, P = ctx.cfg, ctx.saved, ctx.P
        B, S, H = cfg["B"], cfg["S"], cfg["heads"]
        mask_inds, pad_start, mask_rows = ctx.consts
        M, C, L, nblk, pshapes = ctx.shapes
        LM = L * M
        dev = d_inter.device
        f32 = dict(device=dev, dtype=torch.float32)
        d_inter = d_inter.contiguous().float()
        # one zero-filled buffer for everything the kernels accumulate into
        sink = cfg.get("sink")                           # the parameters' own .grad buffers: the kernels accumulate in place
        sizes = [M * C, 2 * C, S * C, C, LM * C] + ([] if sink is not None else [int(torch.Size(s).numel()) for s in pshapes])
        flat = torch.zeros(sum(sizes), **f32)
        views, off = [], 0
        for sz in sizes:
            views.append(flat[off:off + sz])
            off += sz
        dvideo, dtype, dpos, dpad = views[0].view(M, C), views[1].view(2, C), views[2].view(S, C), views[3].view(1, C)
        dh = views[4].view(LM, C)                        # gradient w.r.t. the block outputs of all levels, rows as saved
        G = sink if sink is not None else [v.view(s) for v, s in zip(views[5:], pshapes)]
        rows_all = (mask_rows.unsqueeze(0) + torch.arange(L, device=dev).unsqueeze(1) * M).reshape(-1)
        dh.index_copy_(0, rows_all, d_inter)
        for i in reversed(range(nblk)):
            (ln1w, ln1b, win, bin_, wout, bout, ln2w, ln2b, wfc, bfc, wproj, bproj) = P[12 * i:12 * i + 12]
            (g_ln1w, g_ln1b, g_win, g_bin, g_wout, g_bout, g_ln2w, g_ln2b, g_wfc, g_bfc, g_wproj, g_bproj) = \
                G[12 * i:12 * i + 12]
            sv = saved[i]
            u = sv["u"]
            # h_out = h_mid + c_proj(QuickGELU(u)),  u = c_fc(LN2(h_mid))
            ops.ot_linear_dw(dh, sv["act"], g_wproj, g_bproj)
            dU = torch.empty_like(u)
            ops.ot_linear_dx(dh, wproj, dU, pre=u)
            ops.ot_linear_dw(dU, sv["xhat2"], g_wfc, g_bfc, ops.OT_X_LN, ln2w, ln2b)
            dA = torch.empty(LM, C, **f32)
            ops.ot_linear_dx(dU, wfc, dA)
            ops.ot_ln_bwd(dA, sv["xhat2"], sv["rstd2"], ln2w, dh, g_ln2w, g_ln2b)
            # h_mid = h_in + out_proj(attn(in_proj(LN1(h_in))))
            ops.ot_linear_dw(dh, sv["o"], g_wout, g_bout)
            dO = torch.empty(LM, C, **f32)
            ops.ot_linear_dx(dh, wout, dO)
            dqkv = torch.empty(LM, 3 * C, **f32)
            ops.ot_attn_bwd(sv["qkv"], sv["probs"], dO, dqkv, L * B, S, H)
            ops.ot_linear_dw(dqkv, sv["xhat1"], g_win, g_bin, ops.OT_X_LN, ln1w, ln1b)
            dA = torch.empty(LM, C, **f32)
            ops.ot_linear_dx(dqkv, win, dA)
            ops.ot_ln_bwd(dA, sv["xhat1"], sv["rstd1"], ln1w, dh, g_ln1w, g_ln1b)
            saved[i] = None
        dtvecs = torch.empty(L, C, **f32)
        for lvl in range(L):
            ops.ot_embed_bwd(dh[lvl * M:(lvl + 1) * M], mask_inds, pad_start, dvideo, dtype, dpos, dpad, dtvecs[lvl], B, S)
        ctx.saved = None
        if sink is not None:
            return (None, dvideo, dtvecs, dtype, dpos, dpad, None, None, None, None) + (None,) * len(pshapes)
        return (None, dvideo, dtvecs, dtype, dpos, dpad, None, None, None, None) + tuple(G)


def order_levels(blocks, video, tvecs, type_w, pos_w, pad_w, x0, noise, mask_inds, pad_start, B, S, heads, coef,
                 eps=1e-5, grad_into_params=False):
    """blocks: the ResidualAttentionBlock modules (parameter containers); returns inter [L*B, C].
    grad_into_params: the backward accumulates the block parameters' gradients straight into their existing `.grad`
    buffers (trainer.PretrainStep: views of the flat gradient) instead of returning 48 tensors for autograd to add."""
    params = []
    for blk in blocks:
        sd = dict(blk.named_parameters())
        params += [sd[n] for n in BLOCK_PARAMS]
    cfg = dict(B=B, S=S, heads=heads, eps=eps, coef=coef)
    if grad_into_params and torch.is_grad_enabled():
        cfg["sink"] = [p.grad for p in params]
        assert all(g is not None and g.is_contiguous() and g.dtype == torch.float32 for g in cfg["sink"])
    return OrderLevels.apply(cfg, video, tvecs, type_w, pos_w, pad_w, x0, noise, mask_inds, pad_start, *params)
