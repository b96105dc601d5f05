"""TEST INFRASTRUCTURE: plain-torch restatement of every function of `procedurevrl_b200.ops` with the same
signatures and the same rounding points (bf16 operands / outputs where the CUDA kernels round).

Two uses: (1) CPU tests monkeypatch `procedurevrl_b200.ops` with these to check the *host logic* -- the
engine's forward/backward schedule, row maps, gradient routing -- against the oracle without a GPU;
(2) they document the contract each CUDA kernel is unit-tested against in tests/test_ops_gpu.py.
Never imported by the product package."""
import math

import torch
import torch.nn.functional as F

BF16, F32 = 0, 1
MAP_IDENT, MAP_SKIPCLS, MAP_SPATIAL, MAP_PATCH, MAP_CLS = 0, 1, 2, 3, 4
EPI_STORE, EPI_GELU, EPI_DGELU, EPI_RESID, EPI_ATOMIC = 0, 1, 2, 3, 4

_launches = [0]


def launch_count():
    return _launches[0]


def lib():
    return None


def _rows(map, M, T, HW, dev):
    """residual-stream row of logical row m; -(bt+1) for the cls rows of MAP_SPATIAL."""
    m = torch.arange(M, device=dev)
    L, S = T * HW, 1 + T * HW
    if map == MAP_IDENT:
        return m
    if map == MAP_SKIPCLS:
        return m + m // L + 1
    if map == MAP_CLS:
        return m * S
    if map == MAP_PATCH:
        bt, n = m // HW, m % HW
        return (bt // T) * S + 1 + n * T + bt % T
    bt, n = m // (HW + 1), m % (HW + 1)
    r = (bt // T) * S + 1 + (n - 1) * T + bt % T
    return torch.where(n == 0, -(bt + 1), r)


def gemm(A, B, out, *, M, N, K, trans=0, epilogue=EPI_STORE, out2=None, bias=None, rowscale=None, rs_div=0,
         map=MAP_IDENT, aux=None, resid=None, add_pos=None, add_time=None, T=1, HW=1, k_splits=0,
         lda=None, ldb=None, ldo=None, colsum=None):
    _launches[0] += 1
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    if trans == 0:
        acc = A[:M, :K].float() @ B[:N, :K].float().t()
    else:
        acc = A[:K, :M].float().t() @ B[:K, :N].float()
    dev = acc.device
    if epilogue == EPI_ATOMIC:
        out.view(-1, out.shape[-1])[:M, :N] += acc
        return out
    if bias is not None:
        acc = acc + bias
    if epilogue == EPI_GELU:      # out = gelu'(z), out2 = gelu(z)
        out[:M] = (0.5 * (1 + torch.erf(acc / math.sqrt(2))) + acc * torch.exp(-0.5 * acc * acc) / math.sqrt(2 * math.pi)).to(out.dtype)
        out2[:M] = F.gelu(acc).to(out2.dtype)
        return out
    if epilogue == EPI_DGELU:
        acc = acc * aux[:M].float()
    if rowscale is not None:
        acc = acc * rowscale[torch.arange(M, device=dev) // rs_div].unsqueeze(1)
    rows = _rows(map, M, T, HW, dev)
    o2 = out.view(-1, out.shape[-1])
    if epilogue == EPI_RESID:
        cls = rows < 0
        if cls.any():
            out2[(-rows[cls] - 1)] = acc[cls]
        tok = ~cls
        val = acc[tok]
        if resid is not None:
            val = val + resid.view(-1, resid.shape[-1])[rows[tok]]
        if add_pos is not None:
            m = torch.arange(M, device=dev)
            bt, n = m // HW, m % HW
            val = val + add_pos[1 + n] + add_time[bt % T]
        o2[rows[tok]] = val
        return out
    if colsum is not None:
        colsum += acc.sum(0)
    o2[rows] = acc.to(out.dtype)
    return out


def patchify(frames, out, patch=16):
    _launches[0] += 1
    Bc, C, T, H, W = frames.shape
    nh, nw = H // patch, W // patch
    x = frames.permute(0, 2, 1, 3, 4).reshape(Bc * T, C, nh, patch, nw, patch).permute(0, 2, 4, 1, 3, 5)
    out.copy_(x.reshape(Bc * T * nh * nw, C * patch * patch).to(out.dtype))
    return out


def patchify_u8(frames, out, patch, mean, std):
    m = torch.tensor(mean, dtype=torch.float32, device=frames.device).view(1, 3, 1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32, device=frames.device).view(1, 3, 1, 1, 1)
    return patchify((frames.float() / 255.0 - m) / s, out, patch)


def cls_init(x, cls_token, pos_embed):
    _launches[0] += 1
    x[:, 0] = cls_token.reshape(-1) + pos_embed.reshape(-1, x.shape[-1])[0]


def _ln_src(x, x_cls, M, D, map, T, HW):
    rows = _rows(map, M, T, HW, x.device)
    xf = x.reshape(-1, D)
    if map != MAP_SPATIAL:
        return xf[rows], rows
    S = 1 + T * HW
    cls = rows < 0
    src = xf[rows.clamp(min=0)].clone()
    b = (-rows[cls] - 1) // T
    src[cls] = x_cls.reshape(-1, D)[b * S]
    return src, rows


def layernorm_fwd(x, w, b, y, stats, M, D, eps, map=MAP_IDENT, x_cls=None, T=1, HW=1):
    _launches[0] += 1
    src, _ = _ln_src(x, x_cls, M, D, map, T, HW)
    mu = src.mean(-1, keepdim=True)
    var = ((src - mu) ** 2).mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + eps)
    y[:M] = ((src - mu) * rstd * w + b).to(y.dtype)
    if stats is not None:
        stats[:M, 0], stats[:M, 1] = mu.squeeze(-1), rstd.squeeze(-1)
    return y


def layernorm_bwd(dy, x, w, stats, dx, dw, db, M, D, map=MAP_IDENT, x_cls=None, T=1, HW=1, emit=None):
    _layernorm_bwd(dy, x, w, stats, dx, dw, db, M, D, map, x_cls, T, HW)
    if emit is not None:        # pvrl_layernorm_bwd_emit == the gather_cast that would follow, for the supported map pairs
        out, emap, rowscale, rs_div, colsum = emit
        assert (map, emap) in ((MAP_IDENT, MAP_SPATIAL), (MAP_SPATIAL, MAP_SKIPCLS), (MAP_SKIPCLS, MAP_IDENT),
                               (MAP_SKIPCLS, MAP_PATCH)) and out.dtype == dy.dtype
        L, S = T * HW, 1 + T * HW
        Bc = M // {MAP_IDENT: S, MAP_SPATIAL: T * (HW + 1), MAP_SKIPCLS: L}[map]
        M2 = Bc * {MAP_SPATIAL: T * (HW + 1), MAP_SKIPCLS: L, MAP_IDENT: S, MAP_PATCH: L}[emap]
        _launches[0] -= 1
        gather_cast(dx, out, M2, D, emap, rowscale=rowscale, rs_div=rs_div, T=T, HW=HW, colsum=colsum)


def _layernorm_bwd(dy, x, w, stats, dx, dw, db, M, D, map=MAP_IDENT, x_cls=None, T=1, HW=1):
    _launches[0] += 1
    src, rows = _ln_src(x, x_cls, M, D, map, T, HW)
    d = dy[:M].float()
    xh = (src - stats[:M, :1]) * stats[:M, 1:2]
    g = d * w
    dxr = stats[:M, 1:2] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    if dw is not None:
        dw += (d * xh).sum(0)
    if db is not None:
        db += d.sum(0)
    dxf = dx.view(-1, D)
    if map == MAP_SPATIAL:
        S = 1 + T * HW
        cls = rows < 0
        dxf.index_add_(0, ((-rows[cls] - 1) // T) * S, dxr[cls])
        dxf.index_add_(0, rows[~cls], dxr[~cls])
    else:
        dxf.index_add_(0, rows, dxr)


def gather_cast(src, out, M, D, map=MAP_IDENT, rowscale=None, rs_div=0, T=1, HW=1, colsum=None):
    _launches[0] += 1
    rows = _rows(map, M, T, HW, src.device)
    sf = src.reshape(-1, D)
    f = torch.ones(M, device=src.device)
    if rowscale is not None:
        f = rowscale[torch.arange(M, device=src.device) // rs_div].clone()
    if map == MAP_SPATIAL:
        S = 1 + T * HW
        cls = rows < 0
        f = torch.where(cls, f / T, f)
        rows = torch.where(cls, ((-rows - 1) // T) * S, rows)
    val = sf[rows] * f.unsqueeze(1)
    out[:M] = val.to(out.dtype)
    if colsum is not None:
        colsum += val.sum(0)
    return out


def cls_merge(x0, side, x2, Bc, T, S, D):
    _launches[0] += 1
    x2[:, 0] = x0[:, 0] + side.view(Bc, T, D).sum(1) / T


def colsum(a, out, M, N):
    _launches[0] += 1
    out += a[:M, :N].float().sum(0)


def cast_weight_table(triples):
    return triples


def cast_weight_multi(table):
    for w, o, oT in table:
        cast_weight(w, o, oT)


def cast_weight(w, w_out, wT_out):
    _launches[0] += 1
    if w_out is not None:
        w_out.copy_(w.to(w_out.dtype))
    if wT_out is not None:
        wT_out.copy_(w.t().to(wT_out.dtype))


def split3(a, out, M, K, pattern, along):
    _launches[0] += 1
    a = a.reshape(M, K)
    hi = a.bfloat16()
    lo = (a - hi.float()).bfloat16()
    parts = (hi, hi, lo) if pattern == 0 else (hi, lo, hi)
    out.copy_(torch.cat(parts, dim=1 if along == 1 else 0))
    return out


def embed_bwd(dx, dcls, dpos, dtime, Bc, D, T, HW):
    _launches[0] += 1
    tok = dx[:, 1:].reshape(Bc, HW, T, D)
    c = dx[:, 0].sum(0)
    if dcls is not None:
        dcls += c
    if dpos is not None:
        dpos[0] += c
        dpos[1:] += tok.sum((0, 2))
    if dtime is not None:
        dtime += tok.sum((0, 1))


def _split_heads(qkv, n_seq, seq, H):
    C = H * 64
    q, k, v = qkv.float().split(C, dim=1)
    return [t.reshape(n_seq, seq, H, 64).transpose(1, 2) for t in (q, k, v)]


def attn_fwd(qkv, out, lse, n_seq, seq, H, scale):
    _launches[0] += 1
    q, k, v = _split_heads(qkv, n_seq, seq, H)
    s = (q @ k.transpose(-1, -2)) * scale
    out.copy_((s.softmax(-1) @ v).transpose(1, 2).reshape(n_seq * seq, H * 64).to(out.dtype))
    if lse is not None:
        lse.copy_(torch.logsumexp(s, -1))
    return out


def attn_bwd(qkv, out, dout, lse, dqkv, n_seq, seq, H, scale):
    _launches[0] += 1
    q, k, v = _split_heads(qkv, n_seq, seq, H)
    do = dout.float().reshape(n_seq, seq, H, 64).transpose(1, 2)
    o = out.float().reshape(n_seq, seq, H, 64).transpose(1, 2)
    p = torch.exp((q @ k.transpose(-1, -2)) * scale - lse.unsqueeze(-1))
    dv = p.transpose(-1, -2) @ do
    dp = do @ v.transpose(-1, -2)
    ds = p * (dp - (do * o).sum(-1, keepdim=True)) * scale
    dq, dk = ds @ k, ds.transpose(-1, -2) @ q
    cat = torch.cat([t.transpose(1, 2).reshape(n_seq * seq, H * 64) for t in (dq, dk, dv)], dim=1)
    dqkv.copy_(cat.to(dqkv.dtype))
    return dqkv


attn_tc_fwd, attn_tc_bwd = attn_fwd, attn_bwd


def linear_small_fwd(x, w, b, y):
    _launches[0] += 1
    y.copy_(x @ w.t() + (b if b is not None else 0))
    return y


def linear_small_bwd(x, w, dy, dx, dw, db):
    _launches[0] += 1
    if dw is not None:
        dw += dy.t() @ x
        if db is not None:
            db += dy.sum(0)
    if dx is not None:
        dx.copy_(dy @ w)


def l2norm_fwd(x, y, norms):
    _launches[0] += 1
    n = x.norm(dim=1)
    y.copy_(x / n.unsqueeze(1))
    if norms is not None:
        norms.copy_(n)
    return y


def l2norm_bwd(y, norms, dy, dx):
    _launches[0] += 1
    dx.copy_((dy - y * (y * dy).sum(1, keepdim=True)) / norms.unsqueeze(1))
    return dx


def sim_logits_fwd(emb, label, logits, inv_temp):
    _launches[0] += 1
    logits.copy_(emb @ label.t() * inv_temp)
    return logits


def sim_logits_bwd(dlogits, label, demb, inv_temp):
    _launches[0] += 1
    demb += dlogits @ label * inv_temp
    return demb


def kl_topk_loss(pred, teacher_logits, row_loss, dpred, teacher_out, topk, gscale=1.0):
    _launches[0] += 1
    M = pred.shape[0]
    t = F.softmax(teacher_logits, 1)
    if topk != 0:
        tv = t.topk(k=topk, dim=1)[0]
        t = (t.unsqueeze(1) * (t.unsqueeze(1) == tv.unsqueeze(2)).float()).sum(1)
        t = t / t.sum(1, keepdim=True)
    logp = F.log_softmax(pred, 1)
    if row_loss is not None:
        row_loss.copy_(torch.where(t > 0, t * (torch.log(t.clamp_min(1e-45)) - logp), torch.zeros_like(t)).sum(1) / M)
    if dpred is not None:
        dpred.copy_((logp.exp() - t) * (gscale / M))
    if teacher_out is not None:
        teacher_out.copy_(t)


def softmax_rows(x, y):
    _launches[0] += 1
    y.copy_(x.softmax(1))
    return y


# ---------------------------------------------------------------------------------------------- order transformer
def _qgelu(u):
    return u * torch.sigmoid(1.702 * u)


def ot_linear_fwd(x, W, bias, y, x_mode=0, ln_w=None, ln_b=None, eps=1e-5, xhat=None, rstd=None, resid=None, act_out=None):
    _launches[0] += 1
    a = x
    if x_mode == 1:
        mean, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
        r = torch.rsqrt(var + eps)
        xh = (x - mean) * r
        a = xh * ln_w + ln_b
        if xhat is not None:
            xhat.copy_(xh)
        if rstd is not None:
            rstd.copy_(r.squeeze(1))
    elif x_mode == 2:
        a = _qgelu(x)
    out = a @ W.t()
    if bias is not None:
        out = out + bias
    if resid is not None:
        out = out + resid
    y.copy_(out)
    if act_out is not None:
        act_out.copy_(_qgelu(out))
    return y


def ot_linear_dx(dY, W, dA, pre=None):
    _launches[0] += 1
    g = dY @ W
    if pre is not None:
        s = torch.sigmoid(1.702 * pre)
        g = g * (s * (1 + 1.702 * pre * (1 - s)))
    dA.copy_(g)
    return dA


def ot_linear_dw(dY, A, dW, db, a_mode=0, ln_w=None, ln_b=None):
    _launches[0] += 1
    a = A * ln_w + ln_b if a_mode == 1 else (_qgelu(A) if a_mode == 2 else A)
    dW += dY.t() @ a
    if db is not None:
        db += dY.sum(0)


def ot_ln_bwd(dA, xhat, rstd, w, dh, dw, db):
    _launches[0] += 1
    g = dA * w
    dh += rstd.unsqueeze(1) * (g - g.mean(1, keepdim=True) - xhat * (g * xhat).mean(1, keepdim=True))
    dw += (dA * xhat).sum(0)
    db += dA.sum(0)


def _ot_split(qkv, B, S, H):
    C = H * 64
    q, k, v = (qkv[:, i * C:(i + 1) * C].reshape(B, S, H, 64).permute(0, 2, 1, 3) for i in range(3))
    return q, k, v


def ot_attn_fwd(qkv, pad_start, probs, o, B, S, H):
    _launches[0] += 1
    q, k, v = _ot_split(qkv, B, S, H)
    s = (q * 0.125) @ k.transpose(-1, -2)
    if pad_start is not None:
        pad = torch.arange(S, device=qkv.device).view(1, 1, 1, S) >= pad_start.view(B, 1, 1, 1)
        s = s.masked_fill(pad, float("-inf"))
    p = s.softmax(-1)
    probs.copy_(p)
    o.copy_((p @ v).permute(0, 2, 1, 3).reshape(B * S, H * 64))
    return o


def ot_attn_bwd(qkv, probs, dO, dqkv, B, S, H):
    _launches[0] += 1
    q, k, v = _ot_split(qkv, B, S, H)
    do = dO.reshape(B, S, H, 64).permute(0, 2, 1, 3)
    dp = do @ v.transpose(-1, -2)
    ds = probs * (dp - (probs * dp).sum(-1, keepdim=True))
    dq, dk, dv = ds @ k * 0.125, ds.transpose(-1, -2) @ q * 0.125, probs.transpose(-1, -2) @ do
    dqkv.copy_(torch.cat([t.permute(0, 2, 1, 3).reshape(B * S, H * 64) for t in (dq, dk, dv)], dim=1))
    return dqkv


def ot_embed_fwd(video, src, noise, ca, cb, mask_inds, pad_start, type_w, pos_w, pad_w, tvec, h, B, S):
    _launches[0] += 1
    C = video.shape[1]
    s_idx = torch.arange(S, device=video.device).view(1, S)
    is_mask = (s_idx == mask_inds.view(B, 1)).unsqueeze(-1)
    is_pad = (s_idx >= pad_start.view(B, 1)).unsqueeze(-1)
    x = torch.where(is_pad, pad_w.view(1, 1, C), video.view(B, S, C))
    x = torch.where(is_mask, (ca * src + cb * noise).view(B, 1, C), x)
    x = x + torch.where(is_mask, type_w[1].view(1, 1, C), type_w[0].view(1, 1, C)) + pos_w.view(1, S, C) + tvec.view(1, 1, C)
    h.copy_(x.reshape(B * S, C))
    return h


def ot_embed_bwd(dh, mask_inds, pad_start, dvideo, dtype, dpos, dpad, dtvec, B, S):
    _launches[0] += 1
    C = dh.shape[1]
    g = dh.view(B, S, C)
    s_idx = torch.arange(S, device=dh.device).view(1, S)
    is_mask = (s_idx == mask_inds.view(B, 1)).unsqueeze(-1)
    is_pad = (s_idx >= pad_start.view(B, 1)).unsqueeze(-1) & ~is_mask
    dtvec.copy_(g.sum((0, 1)))
    dpos += g.sum(0)
    dtype[1] += (g * is_mask).sum((0, 1))
    dtype[0] += (g * ~is_mask).sum((0, 1))
    dpad.view(-1).add_((g * is_pad).sum((0, 1)))
    dvideo += (g * (~is_mask & ~is_pad)).reshape(B * S, C)


# ---------------------------------------------------------------------------------------------- MViTv2 encoder ops
def pool_out_grid(grid, kernel, stride, pad):
    return [(g + 2 * p - k) // s + 1 for g, k, s, p in zip(grid, kernel, stride, pad)]


def ln_any_fwd(x, w, b, y, stats, M, D, eps):
    _launches[0] += 1
    xf = x.reshape(M, D).float()
    mean, var = xf.mean(1), xf.var(1, unbiased=False)
    rstd = 1.0 / torch.sqrt(var + eps)
    y.reshape(M, D).copy_((xf - mean[:, None]) * rstd[:, None] * w + b)
    stats[:, 0], stats[:, 1] = mean, rstd
    return y


def ln_any_bwd(dy, x, w, stats, dx, dw, db, M, D):
    _launches[0] += 1
    d, xf = dy.reshape(M, D).float(), x.reshape(M, D).float()
    mean, rstd = stats[:, :1], stats[:, 1:]
    xh = (xf - mean) * rstd
    g = d * w
    dx.reshape(M, D).copy_(rstd * (g - g.mean(1, keepdim=True) - xh * (g * xh).mean(1, keepdim=True)))
    if dw is not None:
        dw += (d * xh).sum(0)
    if db is not None:
        db += d.sum(0)


def _pool_conv(t, w, grid, kernel, stride, pad):
    """t [B, heads, 1 + L, C] fp32 -> depth-wise Conv3d over the grid of every head, cls row bypasses."""
    if w is None:
        return t
    B, Hh, _, C = t.shape
    T, H, W = grid
    vol = t[:, :, 1:].reshape(B * Hh, T, H, W, C).permute(0, 4, 1, 2, 3)
    vol = F.conv3d(vol, w.reshape(C, 1, *kernel), None, stride=list(stride), padding=list(pad), groups=C)
    return torch.cat((t[:, :, :1], vol.reshape(B, Hh, C, -1).transpose(2, 3)), dim=2)


def pool3d_fwd(src, col0, w, out, heads, C, grid, kernel, stride, pad):
    _launches[0] += 1
    B, N, _ = src.shape
    t = src[:, :, col0:col0 + heads * C].float().reshape(B, N, heads, C).permute(0, 2, 1, 3)
    out.copy_(_pool_conv(t, w, grid, kernel, stride, pad))
    return out


def pool3d_bwd(dout, src, col0, w, din, dw, heads, C, grid, kernel, stride, pad):
    _launches[0] += 1
    B, N, _ = din.shape
    if w is None:
        din[:, :, col0:col0 + heads * C] = dout.permute(0, 2, 1, 3).reshape(B, N, heads * C)
        return
    with torch.enable_grad():
        t = src[:, :, col0:col0 + heads * C].float().reshape(B, N, heads, C).permute(0, 2, 1, 3).detach().requires_grad_(True)
        wv = w.detach().clone().requires_grad_(True)
        gt, gw = torch.autograd.grad(_pool_conv(t, wv, grid, kernel, stride, pad), (t, wv), dout.float())
    din[:, :, col0:col0 + heads * C] = gt.permute(0, 2, 1, 3).reshape(B, N, heads * C)
    if dw is not None:
        dw += gw.reshape(dw.shape)


def maxpool3d_fwd(x, y, arg, grid, kernel, stride, pad):
    _launches[0] += 1
    B, _, D = x.shape
    T, H, W = grid
    vol = x[:, 1:].float().reshape(B, T, H, W, D).permute(0, 4, 1, 2, 3)
    o, idx = F.max_pool3d(vol, list(kernel), list(stride), list(pad), return_indices=True)
    y[:, 0] = x[:, 0]
    y[:, 1:] = o.reshape(B, D, -1).transpose(1, 2)
    arg.copy_(idx.reshape(B, D, -1).transpose(1, 2))
    return y


def maxpool3d_bwd(dy, arg, dx, grid, kernel, stride, pad):
    _launches[0] += 1
    dx[:, 0] = dy[:, 0].float()
    dx[:, 1:].scatter_add_(1, arg.long(), dy[:, 1:].float())


def im2col3d(frames, out, kernel, stride, pad):
    _launches[0] += 1
    B, Cin = frames.shape[:2]
    xp = F.pad(frames, (pad[2], pad[2], pad[1], pad[1], pad[0], pad[0]))
    cols = xp.unfold(2, kernel[0], stride[0]).unfold(3, kernel[1], stride[1]).unfold(4, kernel[2], stride[2])
    cols = cols.permute(0, 2, 3, 4, 1, 5, 6, 7).reshape(out.shape[0], Cin * kernel[0] * kernel[1] * kernel[2])
    out.zero_()
    out[:, :cols.shape[1]] = cols
    return out


def _pooled_attn(q, k, v, bq, kgrid, scale, resid):
    B, Hh, Nq, C = q.shape
    Kt, Kh, Kw = kgrid
    s = (q * scale) @ k.transpose(-1, -2)
    bt, bh, bw = bq[..., :Kt], bq[..., Kt:Kt + Kh], bq[..., Kt + Kh:]
    bias = (bt[..., :, None, None] + bh[..., None, :, None] + bw[..., None, None, :]).reshape(B, Hh, Nq - 1, -1)
    s = torch.cat((s[:, :, :1], torch.cat((s[:, :, 1:, :1], s[:, :, 1:, 1:] + bias), dim=3)), dim=2)
    o = s.softmax(-1) @ v
    if resid:
        o = torch.cat((o[:, :, :1], o[:, :, 1:] + q[:, :, 1:]), dim=2)
    return o.transpose(1, 2).reshape(B, Nq, Hh * C), torch.logsumexp(s, dim=-1)


def pooled_attn_fwd(q, k, v, bq, out, lse, kgrid, scale, resid):
    _launches[0] += 1
    o, l = _pooled_attn(q.float(), k.float(), v.float(), bq, kgrid, scale, resid)
    out.copy_(o)
    lse.copy_(l)
    return out


def pooled_attn_bwd(q, k, v, bq, dout, lse, dq, dk, dv, dbq, delta, kgrid, scale, resid):
    _launches[0] += 2
    with torch.enable_grad():
        ins = [t.detach().float().clone().requires_grad_(True) for t in (q, k, v, bq)]
        o, _ = _pooled_attn(*ins, kgrid, scale, resid)
        gq, gk, gv, gb = torch.autograd.grad(o, ins, dout.float())
    dq.copy_(gq)
    dk += gk
    dv += gv
    dbq.copy_(gb)
    B, Hh, Nq, C = q.shape
    o_attn, _ = _pooled_attn(q.float(), k.float(), v.float(), bq, kgrid, scale, False)      # delta = dO . (P V)
    delta.copy_((dout.float() * o_attn).reshape(B, Nq, Hh, C).sum(-1).transpose(1, 2))


ALL = [n for n, v in list(globals().items()) if callable(v) and not n.startswith("_") and n not in ("F", "math", "torch")]
