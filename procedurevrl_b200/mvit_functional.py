"""autograd.Function wrappers of the MViTv2 encoder kernels (include/pvrl.h "MViTv2 video encoder", csrc/mvit.cu) -- the
ops around the encoder's Linear layers (those are tc_functional.tc_linear / tc_mlp on the tcgen05 GEMM).

  layer_norm(x, w, b, eps, out_dtype)            nn.LayerNorm over widths 96 .. 768          attention.py:239-280,530,556
  pool_qkv(qkv, wq, wk, wv, ...)                 attention_pool of Q / K / V read in place from the qkv GEMM output,
                                                 written head-major [B, heads, 1 + L', 96]      attention.py:14-48,330-352
  rel_pos_projections(q, ...)                    q . R_t[dt], q . R_h[dh], q . R_w[dw]           attention.py:51-159
  pooled_attention(q, k, v, bq, ...)             softmax(q k^T scale + bias) v + q              attention.py:360-411
  max_pool_skip(x, ...)                          MaxPool3d of the skip path                     attention.py:521-543
  conv3d_stem_rows(frames, ...)                  im2col rows of the Conv3d patch stem           stem_helper.py:290-322

The Functions own save-for-backward and buffer allocation; all arithmetic happens behind the C ABI, except
`rel_pos_projections` (three einsums of [tokens, 96] x [<= 16 table rows, 96], 5 % of the attention FLOPs, left to torch
this round: fusing them into the attention kernel as one extra MMA per query tile is DESIGN.md section 9's plan)."""
import math

import torch

from . import ops


def _act(t, dtype):
    return t if t.dtype == dtype else t.to(dtype)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, eps, out_dtype):
        D = x.shape[-1]
        x2 = x.reshape(-1, D).contiguous()
        M = x2.shape[0]
        y = torch.empty(M, D, device=x.device, dtype=out_dtype)
        stats = torch.empty(M, 2, device=x.device, dtype=torch.float32)
        wf, bf = w.detach().float().contiguous(), b.detach().float().contiguous()
        ops.ln_any_fwd(x2, wf, bf, y, stats, M, D, eps)
        ctx.save_for_backward(x2, wf, stats)
        ctx.shape = x.shape
        return y.reshape(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, wf, stats = ctx.saved_tensors
        M, D = x2.shape
        dy2 = dy.reshape(M, D).contiguous()
        dx = torch.empty_like(x2)
        dw = torch.zeros(D, device=dy.device, dtype=torch.float32)
        db = torch.zeros(D, device=dy.device, dtype=torch.float32)
        ops.ln_any_bwd(dy2, x2, wf, stats, dx, dw, db, M, D)
        return dx.reshape(ctx.shape), dw, db, None, None


def layer_norm(x, weight, bias, eps, out_dtype):
    """LayerNorm over the last axis; x fp32 or bf16, output in `out_dtype`, statistics in fp32."""
    return _LayerNorm.apply(x, weight, bias, eps, out_dtype)


_NO_POOL = ((1, 1, 1), (1, 1, 1), (0, 0, 0))


def _pool_args(w, kernel, stride):
    """(weights [C, k] or None, kernel, stride, padding) of one attention_pool; no weights = re-layout only."""
    if w is None:
        return (None,) + _NO_POOL
    return w.detach().float().reshape(w.shape[0], -1).contiguous(), tuple(kernel), tuple(stride), tuple(k // 2 for k in kernel)


class _PoolQKV(torch.autograd.Function):
    """qkv [B, 1 + T*H*W, 3 * heads * C] (the qkv Linear's output, attention.py:330-335) -> q, k, v [B, heads, 1 + L', C]:
    the `reshape(B, N, 3, heads, C).permute(2, 0, 3, 1, 4)` never materialises, each pooling kernel reads its slice in place."""

    @staticmethod
    def forward(ctx, qkv, wq, wk, wv, heads, C, grid, kernel, stride_q, stride_kv):
        qkv = qkv.contiguous()
        B = qkv.shape[0]
        outs, args = [], []
        for which, (w, stride) in enumerate(((wq, stride_q), (wk, stride_kv), (wv, stride_kv))):
            w2, kern, st, pad = _pool_args(w, kernel, stride)
            og = ops.pool_out_grid(grid, kern, st, pad)
            out = torch.empty(B, heads, 1 + og[0] * og[1] * og[2], C, device=qkv.device, dtype=qkv.dtype)
            ops.pool3d_fwd(qkv, which * heads * C, w2, out, heads, C, grid, kern, st, pad)
            outs.append(out)
            args.append((w2, kern, st, pad))
        ctx.save_for_backward(qkv)
        ctx.args, ctx.heads, ctx.C, ctx.grid = args, heads, C, tuple(grid)
        ctx.wshapes = [None if w is None else w.shape for w in (wq, wk, wv)]
        return tuple(outs)

    @staticmethod
    def backward(ctx, dq, dk, dv):
        (qkv,) = ctx.saved_tensors
        dqkv = torch.empty_like(qkv)
        dws = []
        for which, (dout, (w2, kern, st, pad)) in enumerate(zip((dq, dk, dv), ctx.args)):
            dw = None if w2 is None else torch.zeros_like(w2)
            if dout is None:                        # an output nobody differentiated through: its slice of dqkv is zero
                og = ops.pool_out_grid(ctx.grid, kern, st, pad)
                dout = torch.zeros(qkv.shape[0], ctx.heads, 1 + og[0] * og[1] * og[2], ctx.C, device=qkv.device, dtype=qkv.dtype)
            ops.pool3d_bwd(_act(dout, qkv.dtype).contiguous(), qkv, which * ctx.heads * ctx.C, w2, dqkv, dw, ctx.heads, ctx.C,
                           ctx.grid, kern, st, pad)
            dws.append(None if dw is None else dw.reshape(ctx.wshapes[which]))
        return (dqkv, *dws, None, None, None, None, None, None)


def pool_qkv(qkv, wq, wk, wv, heads, C, grid, kernel, stride_q, stride_kv):
    """attention_pool (conv mode, without its LayerNorm) of Q / K / V.  wq / wk / wv: depth-wise Conv3d weights [C, 1, kt, kh, kw]
    or None where the block does not pool that tensor.  Returns head-major q, k, v."""
    return _PoolQKV.apply(qkv, wq, wk, wv, heads, C, tuple(grid), tuple(kernel), tuple(stride_q or (1, 1, 1)),
                          tuple(stride_kv or (1, 1, 1)))


_REL_INDEX = {}     # (n_q, n_k, device) -> int64 [n_q, n_k] row indices into the relative-position table


def _rel_index(n_q, n_k, device):
    """table row of (query position i, key position j): i * rq - j * rk + (n_k - 1) * rk with rq = max(n_k / n_q, 1),
    rk = max(n_q / n_k, 1) (attention.py:76-99,124-134).  Built once per geometry and device: no host-to-device copy (and no
    host synchronisation) inside the step."""
    key = (n_q, n_k, str(device))
    idx = _REL_INDEX.get(key)
    if idx is None:
        rq, rk = max(n_k / n_q, 1.0), max(n_q / n_k, 1.0)
        dist = torch.arange(n_q)[:, None] * rq - torch.arange(n_k)[None, :] * rk + (n_k - 1) * rk
        idx = _REL_INDEX[key] = dist.long().to(device)
    return idx


def rel_table(table, n_q, n_k):
    """R[i, j] = table[index(i, j)] (see _rel_index); the table is linearly resized first when its length is not
    2 * max(n_q, n_k) - 1 (attention.py:51-64)."""
    d = 2 * max(n_q, n_k) - 1
    if table.shape[0] != d:
        table = torch.nn.functional.interpolate(table.t().unsqueeze(0), size=d, mode="linear").squeeze(0).t()
    return table[_rel_index(n_q, n_k, table.device)]


def rel_pos_projections(q, q_grid, k_grid, rel_h, rel_w, rel_t):
    """bq [B, heads, Nq - 1, Kt + Kh + Kw] fp32: the (unscaled, pooled) non-cls queries against the rows of the three
    relative-position tables their grid position selects -- the decomposed bias of attention.py:101-159 before it is
    broadcast over the key grid (that broadcast happens inside the attention kernel)."""
    B, Hh, Nq, C = q.shape
    (qt, qh, qw), (kt, kh, kw) = q_grid, k_grid
    rq = q[:, :, 1:].float().reshape(B, Hh, qt, qh, qw, C)
    bt = torch.einsum("bnthwc,tkc->bnthwk", rq, rel_table(rel_t.float(), qt, kt))
    bh = torch.einsum("bnthwc,hkc->bnthwk", rq, rel_table(rel_h.float(), qh, kh))
    bw = torch.einsum("bnthwc,wkc->bnthwk", rq, rel_table(rel_w.float(), qw, kw))
    return torch.cat((bt, bh, bw), dim=-1).reshape(B, Hh, Nq - 1, kt + kh + kw).contiguous()


class _PooledAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, bq, k_grid, scale, residual_pooling):
        q, k, v, bq = q.contiguous(), k.contiguous(), v.contiguous(), bq.contiguous().float()
        B, Hh, Nq, C = q.shape
        out = torch.empty(B, Nq, Hh * C, device=q.device, dtype=q.dtype)
        lse = torch.empty(B, Hh, Nq, device=q.device, dtype=torch.float32)
        ops.pooled_attn_fwd(q, k, v, bq, out, lse, k_grid, scale, residual_pooling)
        ctx.save_for_backward(q, k, v, bq, lse)
        ctx.cfg = (tuple(k_grid), scale, residual_pooling)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, bq, lse = ctx.saved_tensors
        k_grid, scale, resid = ctx.cfg
        dq = torch.empty_like(q)
        dk = torch.zeros(k.shape, device=k.device, dtype=torch.float32)
        dv = torch.zeros(v.shape, device=v.device, dtype=torch.float32)
        dbq = torch.empty_like(bq)
        delta = torch.empty_like(lse)
        ops.pooled_attn_bwd(q, k, v, bq, _act(dout, q.dtype).contiguous(), lse, dq, dk, dv, dbq, delta, k_grid, scale, resid)
        return dq, _act(dk, k.dtype), _act(dv, v.dtype), dbq, None, None, None


def pooled_attention(q, k, v, bq, k_grid, scale, residual_pooling=True):
    """q [B, heads, Nq, 96], k / v [B, heads, Nk, 96], bq from rel_pos_projections -> [B, Nq, heads * 96]."""
    return _PooledAttention.apply(q, k, v, bq, tuple(k_grid), float(scale), bool(residual_pooling))


class _MaxPoolSkip(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, grid, kernel, stride, pad):
        x = x.contiguous()
        B, _, D = x.shape
        og = ops.pool_out_grid(grid, kernel, stride, pad)
        Lo = og[0] * og[1] * og[2]
        y = torch.empty(B, 1 + Lo, D, device=x.device, dtype=x.dtype)
        arg = torch.empty(B, Lo, D, device=x.device, dtype=torch.int32)
        ops.maxpool3d_fwd(x, y, arg, grid, kernel, stride, pad)
        ctx.save_for_backward(arg)
        ctx.cfg = (tuple(x.shape), x.dtype, grid, kernel, stride, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        shape, dtype, grid, kernel, stride, pad = ctx.cfg
        dx = torch.zeros(shape, device=dy.device, dtype=torch.float32)
        ops.maxpool3d_bwd(dy.contiguous(), arg, dx, grid, kernel, stride, pad)
        return _act(dx, dtype), None, None, None, None


def max_pool_skip(x, grid, stride):
    """MaxPool3d(kernel = stride + 1 where stride > 1, stride, padding = kernel // 2) over the token grid, cls bypasses
    (attention.py:521-528,537-543)."""
    kernel = tuple(s + 1 if s > 1 else s for s in stride)
    return _MaxPoolSkip.apply(x, tuple(grid), kernel, tuple(stride), tuple(k // 2 for k in kernel))


def conv3d_stem_rows(frames, kernel, stride, pad, dtype, k_align=64):
    """frames [B, Cin, T, H, W] fp32 -> (rows [B * T'*H'*W', Kpad], output grid): the Conv3d stem as a GEMM operand.  The
    window length Cin * kt*kh*kw (441) is zero-padded to a multiple of `k_align` (448) for the TMA operand maps; the
    caller pads the weight the same way.  No gradient flows to the frames."""
    B, Cin, T, H, W = frames.shape
    og = ops.pool_out_grid((T, H, W), kernel, stride, pad)
    K = Cin * math.prod(kernel)
    Kpad = (K + k_align - 1) // k_align * k_align
    rows = torch.empty(B * og[0] * og[1] * og[2], Kpad, device=frames.device, dtype=dtype)
    ops.im2col3d(frames.detach().float().contiguous(), rows, tuple(kernel), tuple(stride), tuple(pad))
    return rows, og
