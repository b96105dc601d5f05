"""CPU oracle for the MViTv2 video encoder of ProcedureVRL (BASELINE config 5).

TEST INFRASTRUCTURE ONLY -- nothing in `procedurevrl_b200/` imports this file.  It restates, as plain functional fp32
torch, what the reference's `MViT_encoder` computes (paths relative to the reference root, lib/models/slowfast_mvit/):

  geometry()            mvit.py:77-228     per-block widths, heads, Q / KV pooling kernels and strides from the MVIT.* keys
  stem                  stem_helper.py:290-322   Conv3d(3 -> EMBED_DIM, PATCH_KERNEL, PATCH_STRIDE, PATCH_PADDING), tokens (t h w)
  pool_tokens()         attention.py:14-48       depth-wise Conv3d over the (t h w) grid per head, cls token bypasses, LayerNorm
  rel_pos_bias()        attention.py:51-159      decomposed relative position bias: q . R_h[dh] + q . R_w[dw] + q . R_t[dt]
  attention()           attention.py:282-411     qkv Linear, pooled q/k/v, scaled scores + bias, softmax, residual pooling, proj
  block()               attention.py:530-567     norm1, attention, widened / max-pooled skip, MLP with erf GELU
  encoder_forward()     mvit.py:338-406          cls token, blocks, final LayerNorm, cls row

It is pinned by `oracle/make_golden_mvit.py`, which runs the UNMODIFIED reference (through `oracle/ref_shims.py`) on the
same seeded parameters / clips and stores its outputs under tests/golden/mvit_*.pt; `tests/test_mvit_oracle_golden.py`
checks this file against them.  The product path (procedurevrl_b200/lib/models/mvit.py on csrc/mvit.cu)
is tested against the same goldens (tests/test_mvit_cpu.py, tests/test_mvit_gpu.py); it never imports this file.

Only what the shipped MViT YAMLs use is restated: MODE conv, POOL_FIRST False, SEPARATE_QKV False, CLS_EMBED_ON True,
USE_ABS_POS False, REL_POS_SPATIAL / REL_POS_TEMPORAL True, RESIDUAL_POOLING True, DIM_MUL_IN_ATT True, NORM layernorm,
no layer scale, dropout 0.  `geometry()` raises on anything else.
"""
import math

import torch
import torch.nn.functional as F

PRE = "model.video_encoder."
LN_EPS = 1e-6          # partial(nn.LayerNorm, eps=1e-6), mvit.py:68-69


def _round_width(width, mult, divisor=1):
    """utils.py:7-20 (min_width = 1)."""
    if not mult:
        return width
    width = width * mult
    out = max(1, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


def geometry(mvit, num_frames, crop):
    """Per-block geometry from the MVIT.* config node (dict-like).  Returns a dict:
    patch (kernel, stride, padding), grid [T, H, W] after the stem, embed_dim, blocks = list of dicts with
    dim, dim_out (= attention width, DIM_MUL_IN_ATT), heads, kernel_q / stride_q, kernel_kv / stride_kv ([] = no pooling),
    input grid, rel_sp (rows of rel_pos_h / rel_pos_w), rel_t (rows of rel_pos_t)."""
    need = dict(MODE="conv", POOL_FIRST=False, SEPARATE_QKV=False, CLS_EMBED_ON=True, USE_ABS_POS=False,
                REL_POS_SPATIAL=True, REL_POS_TEMPORAL=True, RESIDUAL_POOLING=True, DIM_MUL_IN_ATT=True, NORM="layernorm")
    for k, v in need.items():
        if k in mvit and mvit[k] != v:
            raise NotImplementedError(f"mvit_oracle restates MVIT.{k} = {v} only (got {mvit[k]})")
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
    # adaptive KV stride (mvit.py:153-163): starts at POOL_KV_STRIDE_ADAPTIVE and shrinks wherever Q is pooled
    skv = list(mvit["POOL_KV_STRIDE_ADAPTIVE"])
    stride_kv = []
    for i in range(depth):
        if stride_q[i]:
            skv = [max(skv[d] // stride_q[i][d], 1) for d in range(3)]
        stride_kv.append(list(skv))
    blocks, dim, heads, size = [], mvit["EMBED_DIM"], mvit["NUM_HEADS"], list(grid)
    for i in range(depth):
        heads = _round_width(heads, head_mul[i])
        dim_out = _round_width(dim, dim_mul[i], divisor=_round_width(heads, head_mul[i]))
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


def param_shapes(geo, in_chans=3):
    """name -> shape of every `model.video_encoder.*` parameter (the reference's state_dict schema for this geometry)."""
    k = geo["patch"][0]
    D0 = geo["embed_dim"]
    s = {PRE + "cls_token": (1, 1, D0), PRE + "patch_embed.proj.weight": (D0, in_chans, *k), PRE + "patch_embed.proj.bias": (D0,)}
    for i, b in enumerate(geo["blocks"]):
        p = f"{PRE}blocks.{i}."
        d, do, hd = b["dim"], b["dim_out"], b["dim_out"] // b["heads"]
        s[p + "norm1.weight"], s[p + "norm1.bias"] = (d,), (d,)
        s[p + "attn.rel_pos_h"], s[p + "attn.rel_pos_w"], s[p + "attn.rel_pos_t"] = (b["rel_sp"], hd), (b["rel_sp"], hd), (b["rel_t"], hd)
        s[p + "attn.qkv.weight"], s[p + "attn.qkv.bias"] = (3 * do, d), (3 * do,)
        s[p + "attn.proj.weight"], s[p + "attn.proj.bias"] = (do, do), (do,)
        for nm, kern in (("q", b["kernel_q"]), ("k", b["kernel_kv"]), ("v", b["kernel_kv"])):
            if kern and not (math.prod(kern) == 1 and math.prod(b["stride_q" if nm == "q" else "stride_kv"]) == 1):
                s[p + f"attn.pool_{nm}.weight"] = (hd, 1, *kern)
                s[p + f"attn.norm_{nm}.weight"], s[p + f"attn.norm_{nm}.bias"] = (hd,), (hd,)
        s[p + "norm2.weight"], s[p + "norm2.bias"] = (do,), (do,)
        hidden = int(do * geo["mlp_ratio"])
        s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = (hidden, do), (hidden,)
        s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = (do, hidden), (do,)
        if d != do:
            s[p + "proj.weight"], s[p + "proj.bias"] = (do, d), (do,)
    s[PRE + "norm.weight"], s[PRE + "norm.bias"] = (geo["out_dim"],), (geo["out_dim"],)
    return s


def seeded_state(shapes, seed):
    """Deterministic non-degenerate values for a {name: shape} schema (in sorted-name order, one generator):
    LayerNorm scales around 1, pooling kernels large enough to matter, everything else N(0, small)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        r = torch.randn(*shape, generator=g)
        if "norm" in name and name.endswith(".weight"):
            v = 1.0 + 0.1 * r
        elif name.endswith(".bias"):
            v = 0.02 * r
        elif ".pool_" in name:
            v = 0.25 * r
        elif "rel_pos" in name or "cls_token" in name:
            v = 0.05 * r
        else:
            fan_in = math.prod(shape[1:]) if len(shape) > 1 else shape[0]
            v = r / math.sqrt(fan_in)
        out[name] = v
    return out


def synthetic_clips(B, T, crop, seed):
    """uint8 U[0,255] -> /255 -> (x - 0.45) / 0.225, as the loader does (datasets/utils.py:309-326)."""
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (B, 3, T, crop, crop), generator=g, dtype=torch.uint8)
    return (u8.float() / 255.0 - 0.45) / 0.225


# ---------------------------------------------------------------------------------------------------------------- ops
def pool_tokens(x, grid, weight, stride, ln_w=None, ln_b=None, mode="conv"):
    """x [B, heads, 1 + T*H*W, C]: the cls token passes through, the (t h w) grid of every head goes through a depth-wise
    3-D convolution (or a max pool: mode "max", kernel = weight given as a list) -- attention.py:14-48 -- then LayerNorm over C."""
    cls, tok = x[:, :, :1], x[:, :, 1:]
    B, Hh, L, C = tok.shape
    T, H, W = grid
    vol = tok.reshape(B * Hh, T, H, W, C).permute(0, 4, 1, 2, 3)
    if mode == "conv":
        k = weight.shape[2:]
        vol = F.conv3d(vol, weight, None, stride=stride, padding=[s // 2 for s in k], groups=C)
    else:
        vol = F.max_pool3d(vol, weight, stride, [s // 2 for s in weight])
    new_grid = list(vol.shape[2:])
    tok = vol.reshape(B, Hh, C, -1).transpose(2, 3)
    y = torch.cat((cls, tok), dim=2)
    if ln_w is not None:
        y = F.layer_norm(y, (C,), ln_w, ln_b, LN_EPS)
    return y, new_grid


def _rel_table(table, n_q, n_k):
    """R[i, j] = table[(i * rq - j * rk + (n_k - 1) * rk)] with rq = max(n_k / n_q, 1), rk = max(n_q / n_k, 1):
    attention.py:66-80,124-134 (the table is linearly resized first if its length is not 2 * max(n_q, n_k) - 1)."""
    d = 2 * max(n_q, n_k) - 1
    if table.shape[0] != d:
        table = F.interpolate(table.t().unsqueeze(0), size=d, mode="linear").squeeze(0).t()
    rq, rk = max(n_k / n_q, 1.0), max(n_q / n_k, 1.0)
    dist = torch.arange(n_q)[:, None] * rq - torch.arange(n_k)[None, :] * rk + (n_k - 1) * rk
    return table[dist.long()]                       # [n_q, n_k, C]


def rel_pos_bias(q, q_grid, k_grid, rel_h, rel_w, rel_t):
    """Decomposed bias for the non-cls queries x non-cls keys: [B, heads, Tq*Hq*Wq, Tk*Hk*Wk] (attention.py:51-159)."""
    B, Hh, _, C = q.shape
    (qt, qh, qw), (kt, kh, kw) = q_grid, k_grid
    rq = q[:, :, 1:].reshape(B, Hh, qt, qh, qw, C)
    bh = torch.einsum("bnthwc,hkc->bnthwk", rq, _rel_table(rel_h, qh, kh))
    bw = torch.einsum("bnthwc,wkc->bnthwk", rq, _rel_table(rel_w, qw, kw))
    bt = torch.einsum("bnthwc,tkc->bnthwk", rq, _rel_table(rel_t, qt, kt))
    bias = bt[..., :, None, None] + bh[..., None, :, None] + bw[..., None, None, :]
    return bias.reshape(B, Hh, qt * qh * qw, kt * kh * kw)


def attention(p, pre, x, grid, blk):
    """MultiScaleAttention.forward (attention.py:282-411) for POOL_FIRST False / shared conv pooling."""
    B, N, _ = x.shape
    Hh, do = blk["heads"], blk["dim_out"]
    qkv = F.linear(x, p[pre + "qkv.weight"], p[pre + "qkv.bias"]).reshape(B, N, 3, Hh, do // Hh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q_grid = k_grid = list(grid)
    if pre + "pool_q.weight" in p:
        q, q_grid = pool_tokens(q, grid, p[pre + "pool_q.weight"], blk["stride_q"], p[pre + "norm_q.weight"], p[pre + "norm_q.bias"])
    if pre + "pool_k.weight" in p:
        k, k_grid = pool_tokens(k, grid, p[pre + "pool_k.weight"], blk["stride_kv"], p[pre + "norm_k.weight"], p[pre + "norm_k.bias"])
        v, _ = pool_tokens(v, grid, p[pre + "pool_v.weight"], blk["stride_kv"], p[pre + "norm_v.weight"], p[pre + "norm_v.bias"])
    scale = (do // Hh) ** -0.5
    att = (q * scale) @ k.transpose(-2, -1)
    bias = rel_pos_bias(q, q_grid, k_grid, p[pre + "rel_pos_h"], p[pre + "rel_pos_w"], p[pre + "rel_pos_t"])
    att = torch.cat((att[:, :, :1], torch.cat((att[:, :, 1:, :1], att[:, :, 1:, 1:] + bias), dim=3)), dim=2)
    o = att.softmax(dim=-1) @ v
    o = torch.cat((o[:, :, :1], o[:, :, 1:] + q[:, :, 1:]), dim=2)           # residual pooling: every row but cls (:397-401)
    o = o.transpose(1, 2).reshape(B, -1, do)
    return F.linear(o, p[pre + "proj.weight"], p[pre + "proj.bias"]), q_grid


def block(p, pre, x, grid, blk):
    """MultiScaleBlock.forward (attention.py:530-567), DIM_MUL_IN_ATT: the skip is the widened *normalised* input."""
    xn = F.layer_norm(x, (blk["dim"],), p[pre + "norm1.weight"], p[pre + "norm1.bias"], LN_EPS)
    a, new_grid = attention(p, pre + "attn.", xn, grid, blk)
    skip = F.linear(xn, p[pre + "proj.weight"], p[pre + "proj.bias"]) if blk["dim"] != blk["dim_out"] else x
    sq = blk["stride_q"]
    if sq and math.prod(sq) > 1:                                              # MaxPool3d skip (:521-528)
        kern = [s + 1 if s > 1 else s for s in sq]
        skip, _ = pool_tokens(skip.unsqueeze(1), grid, kern, sq, mode="max")
        skip = skip.squeeze(1)
    x = skip + a
    xn = F.layer_norm(x, (blk["dim_out"],), p[pre + "norm2.weight"], p[pre + "norm2.bias"], LN_EPS)
    h = F.gelu(F.linear(xn, p[pre + "mlp.fc1.weight"], p[pre + "mlp.fc1.bias"]))
    return x + F.linear(h, p[pre + "mlp.fc2.weight"], p[pre + "mlp.fc2.bias"]), new_grid


def encoder_forward(p, x, geo, taps=None):
    """MViT_encoder.forward (mvit.py:338-406): clips [B, 3, T, H, W] -> cls feature [B, out_dim]."""
    kern, stride, pad = geo["patch"]
    x = F.conv3d(x, p[PRE + "patch_embed.proj.weight"], p[PRE + "patch_embed.proj.bias"], stride=stride, padding=pad)
    B = x.shape[0]
    grid = list(x.shape[2:])
    assert grid == geo["grid"], (grid, geo["grid"])
    x = torch.cat((p[PRE + "cls_token"].expand(B, -1, -1), x.flatten(2).transpose(1, 2)), dim=1)
    for i, blk in enumerate(geo["blocks"]):
        x, grid = block(p, f"{PRE}blocks.{i}.", x, grid, blk)
        if taps is not None:
            taps.append(x)
    x = F.layer_norm(x, (geo["out_dim"],), p[PRE + "norm.weight"], p[PRE + "norm.bias"], LN_EPS)
    return x[:, 0]


def match_lang_forward(p, x, geo, label_emb, temp=0.02, taps=None):
    """lib/models/mvit.py:111-124 (DEV.MATCH_LANG_EMB): head Linear, L2-normalise, cosine logits / TEMP against the
    (row-normalised) step bank."""
    f = encoder_forward(p, x, geo, taps)
    e = F.linear(f, p["model.head.weight"], p["model.head.bias"])
    e = e / e.norm(dim=1, keepdim=True)
    return e @ label_emb.t() / temp
