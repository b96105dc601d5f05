"""Explicit forward / backward schedule of the TimeSformer divided space-time encoder on the sm_100a kernels.

This replaces `VisionTransformer.forward_features` (reference lib/models/vit.py:365-423), `Block.forward`
(vit.py:120-158) and their autograd with a hand-ordered sequence of C-ABI calls (procedurevrl_b200.ops):
no tracing compiler, no eager tensor math.  torch is used for device memory (the caching allocator hands
out the activation buffers) and for the autograd hook-up (`EncoderFunction`) only.

Data layout in HBM
  x      fp32 [Bc, S = 1 + HW*T, D]   residual stream, token order (h w t) with t fastest, cls first
  act    bf16 (precision "bf16") or fp32 ("bf16x3") row-major [rows, features]: LayerNorm outputs, qkv,
         attention outputs, MLP hidden; the temporal view uses rows (b, hw, t) == the natural order, the
         spatial view rows (b, t, [cls, hw]) -- the permutes of vit.py:131,138-143,150-153 are folded into
         the row maps of the LayerNorm gather and of the GEMM epilogues, never materialised.
  W      bf16 copies of the fp32 master weights, as [N,K] (forward) and [K,N] (dX) -- refreshed when the
         parameters change.

precision = "bf16x3" is the parity mode: activations stay fp32 and every GEMM operand is split into
bf16 hi/lo parts concatenated along the contraction ([hi|hi|lo] x [hi|lo|hi]), so the same tcgen05 kernel
accumulates hi*hi + hi*lo + lo*hi in fp32 (product error ~2^-17 instead of 2^-9).
"""
import math

import os

import torch

from . import ops

LINEARS = ("temporal_attn.qkv", "temporal_attn.proj", "temporal_fc", "attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")


class EncoderEngine:
    def __init__(self, params, depth, num_frames, embed_dim=768, num_heads=12, mlp_ratio=4, patch=16, eps=1e-6,
                 attention_type="divided_space_time", precision="bf16", prefix="model."):
        """params: dict name -> fp32 CUDA tensor/Parameter with the reference state_dict names."""
        assert precision in ("bf16", "bf16x3")
        assert attention_type in ("divided_space_time", "space_only", "joint_space_time"), attention_type
        self.attention_type = attention_type
        self.divided = attention_type == "divided_space_time"
        self.p = params
        self.pre = prefix
        self.depth, self.T0, self.D, self.H = depth, num_frames, embed_dim, num_heads
        self.hidden = int(embed_dim * mlp_ratio)
        self.patch, self.eps = patch, eps
        self.precision = precision
        self.x3 = precision == "bf16x3"
        self.act_dtype = torch.float32 if self.x3 else torch.bfloat16
        self.scale = (embed_dim // num_heads) ** -0.5
        assert embed_dim // num_heads == 64, "kernels are specialised for head_dim 64"
        self.use_tc_attention = True      # tcgen05 spatial attention in bf16 mode (the fp32 parity mode uses CUDA cores)
        self._wepoch = 0
        self.grad_sink = None             # optional dict name -> fp32 tensor: backward accumulates straight into it
        self.pixel_mean, self.pixel_std = (0.45, 0.45, 0.45), (0.225, 0.225, 0.225)   # DATA.MEAN / DATA.STD (defaults.py:510,516)
        # every LayerNorm backward also writes the next sub-layer's dY operand (pvrl_layernorm_bwd_emit) instead of a
        # separate pvrl_gather_cast pass over dx; PVRL_FUSE_GATHER=0 restores the separate passes
        self.fuse_gather = os.environ.get("PVRL_FUSE_GATHER", "1") != "0"
        self.on_block_bwd_done = None     # optional callable(i): block i's parameter gradients are complete (bucketed all-reduce)
        self._wcache = {}      # name -> (version, W operand [N, K'], W^T operand [K, N'])
        self.grad_names = self._grad_names()

    # ------------------------------------------------------------------------------------------ parameters
    def _grad_names(self):
        # space_only models have no time_embed parameter (vit.py:213-215)
        emb = ("cls_token", "pos_embed", "patch_embed.proj.weight", "patch_embed.proj.bias") \
            if self.attention_type == "space_only" else \
            ("cls_token", "pos_embed", "time_embed", "patch_embed.proj.weight", "patch_embed.proj.bias")
        n = [self.pre + k for k in emb]
        for i in range(self.depth):
            b = f"{self.pre}blocks.{i}."
            for ln in (("norm1", "temporal_norm1", "norm2") if self.divided else ("norm1", "norm2")):
                n += [b + ln + ".weight", b + ln + ".bias"]
            for l in LINEARS:
                if self.divided or not l.startswith("temporal"):
                    n += [b + l + ".weight", b + l + ".bias"]
        n += [self.pre + "norm.weight", self.pre + "norm.bias"]
        return n

    def _weight_ops(self, name):
        """GEMM operand copies of linear weight `name` ([N, K] fp32 master): (W for y = x W^T, W^T for dX)."""
        w = self.p[name]
        ver = (w._version, w.data_ptr(), self._wepoch)
        hit = self._wcache.get(name)
        if hit is not None and hit[0] == ver:
            return hit[1], hit[2]
        w2 = w.detach().reshape(w.shape[0], -1)
        N, K = w2.shape
        dev = w2.device
        mul = 3 if self.x3 else 1
        if hit is not None and hit[1].device == dev:
            wb, wt = hit[1], hit[2]                 # refresh in place: operand addresses stay stable (CUDA graphs)
        else:
            wb = torch.empty(N, mul * K, device=dev, dtype=torch.bfloat16)
            wt = torch.empty(K, mul * N, device=dev, dtype=torch.bfloat16)
        if not self.x3:
            ops.cast_weight(w2, wb, wt)
        else:
            wt32 = torch.empty(K, N, device=dev, dtype=torch.float32)
            ops.cast_weight(w2, None, wt32)
            ops.split3(w2.contiguous(), wb, N, K, 1, 1)
            ops.split3(wt32, wt, K, N, 1, 1)
        self._wcache[name] = (ver, wb, wt)
        return wb, wt

    def _patchify(self, frames, A):
        """im2col of the Conv2d patch embedding; uint8 frames are normalised inside the kernel."""
        if frames.dtype == torch.uint8:
            ops.patchify_u8(frames, A, self.patch, self.pixel_mean, self.pixel_std)
        else:
            ops.patchify(frames, A, self.patch)

    def refresh_weights(self):
        """Re-cast every Linear weight whose master changed, in ONE launch (bf16 mode; called at the top of forward)."""
        if self.x3:
            return
        names = [self.pre + "patch_embed.proj.weight"]
        for i in range(self.depth):
            b = f"{self.pre}blocks.{i}."
            names += [b + l + ".weight" for l in LINEARS if self.divided or not l.startswith("temporal")]
        key = tuple((self.p[n]._version, self.p[n].data_ptr()) for n in names) + (self._wepoch,)
        if getattr(self, "_wkey", None) == key:
            return
        triples = []
        for n in names:
            w = self.p[n]
            w2 = w.detach().reshape(w.shape[0], -1)
            hit = self._wcache.get(n)
            if hit is not None and hit[1].device == w2.device:
                wb, wt = hit[1], hit[2]
            else:
                N, K = w2.shape
                wb = torch.empty(N, K, device=w2.device, dtype=torch.bfloat16)
                wt = torch.empty(K, N, device=w2.device, dtype=torch.bfloat16)
            self._wcache[n] = ((w._version, w.data_ptr(), self._wepoch), wb, wt)
            triples.append((w2, wb, wt))
        sig = tuple((t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr()) for t in triples)
        if getattr(self, "_wtable_sig", None) != sig:          # operand addresses are stable -> the table is built once
            self._wtable, self._wtable_sig = ops.cast_weight_table(triples), sig
        ops.cast_weight_multi(self._wtable)
        self._wkey = key

    def invalidate_weights(self):
        """Force the bf16 operand copies to be re-cast on next use (call once per optimizer step when the step is
        captured in a CUDA graph: the cast kernels must be part of every replay)."""
        self._wepoch += 1

    # ------------------------------------------------------------------------------------------ GEMM helpers
    def _A(self, a, M, K):
        """activation -> GEMM A operand (bf16 as is; fp32 -> [hi|hi|lo] split)."""
        if not self.x3:
            return a, K
        a3 = torch.empty(M, 3 * K, device=a.device, dtype=torch.bfloat16)
        ops.split3(a, a3, M, K, 0, 1)
        return a3, 3 * K

    def linear(self, a, wname, out, M, N, K, **epi):
        """out = epilogue(a @ W^T): nn.Linear forward (vit.py:47-60,72-90,118)."""
        A, Keff = self._A(a, M, K)
        wb, _ = self._weight_ops(wname)
        ops.gemm(A, wb, out, M=M, N=N, K=Keff, **epi)

    def linear_dx(self, dy, wname, out, M, N, K, **epi):
        """out[M, N] = epilogue(dy[M, K] @ W) with W [K, N] (the layer's [out, in] weight)."""
        A, Keff = self._A(dy, M, K)
        _, wt = self._weight_ops(wname)
        ops.gemm(A, wt, out, M=M, N=N, K=Keff, **epi)

    def linear_dw(self, dy, x, gw, gb, Mc, N_w, K_w):
        """gw[N_w, K_w] += dy[Mc, N_w]^T x[Mc, K_w];  gb[N_w] += colsum(dy) (gb = None when the kernel that produced
        dy already accumulated the bias gradient -- `colsum=` of ops.gather_cast / ops.gemm)."""
        if gb is not None:
            ops.colsum(dy, gb, Mc, N_w)
        if not self.x3:
            ops.gemm(dy, x, gw, M=N_w, N=K_w, K=Mc, trans=1, epilogue=ops.EPI_ATOMIC, ldo=K_w)
        else:
            a3 = torch.empty(3 * Mc, N_w, device=dy.device, dtype=torch.bfloat16)
            b3 = torch.empty(3 * Mc, K_w, device=dy.device, dtype=torch.bfloat16)
            ops.split3(dy, a3, Mc, N_w, 0, 0)
            ops.split3(x, b3, Mc, K_w, 1, 0)
            ops.gemm(a3, b3, gw, M=N_w, N=K_w, K=3 * Mc, trans=1, epilogue=ops.EPI_ATOMIC, ldo=K_w)

    def _act(self, rows, cols, dev):
        return torch.empty(rows, cols, device=dev, dtype=self.act_dtype)

    # ------------------------------------------------------------------------------------------ embeddings
    def _pos_time(self, HW, T):
        """pos_embed / time_embed rows for this input size; nearest-neighbour resize as vit.py:375-386,398-402."""
        pos = self.p[self.pre + "pos_embed"].detach()[0]
        pos_idx = te_idx = None
        if self.attention_type == "space_only":               # no time embedding (vit.py:393); the caller passes zeros
            te = None
        else:
            te = self.p[self.pre + "time_embed"].detach()[0]
        if pos.shape[0] != HW + 1:
            P0 = int(round(math.sqrt(pos.shape[0] - 1)))
            P1 = int(round(math.sqrt(HW)))
            src = (torch.arange(P1, device=pos.device).float() * (P0 / P1)).floor().long()
            pos_idx = torch.cat((torch.zeros(1, dtype=torch.long, device=pos.device),
                                 1 + (src.view(-1, 1) * P0 + src.view(1, -1)).reshape(-1)))
            pos = pos.index_select(0, pos_idx).contiguous()
        if te is not None and te.shape[0] != T:
            te_idx = (torch.arange(T, device=te.device).float() * (te.shape[0] / T)).floor().long()
            te = te.index_select(0, te_idx).contiguous()
        return pos, te, pos_idx, te_idx

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, frames, drop_scales=None, save=True):
        """frames fp32 [Bc, 3, T, H, W] -> (cls feature fp32 [Bc, D], saved state or None).
        drop_scales: optional list (len depth) of dicts {'temporal': [Bc*HW], 'spatial': [Bc*T], 'mlp': [Bc]}
        of DropPath factors mask/keep (vit_utils.py:140-155); None entries = identity."""
        assert frames.dtype in (torch.float32, torch.uint8)   # ops._p rejects non-CUDA tensors: there is no CPU path
        frames = frames.contiguous()
        self.refresh_weights()
        if not self.divided:
            return self._forward_plain(frames, drop_scales, save)
        Bc, C, T, Hh, Ww = frames.shape
        D, P = self.D, self.patch
        HW = (Hh // P) * (Ww // P)
        L, S = HW * T, 1 + HW * T
        dev = frames.device
        g = dict(T=T, HW=HW)
        pos, te, pos_idx, te_idx = self._pos_time(HW, T)
        KP = 3 * P * P

        A = self._act(Bc * T * HW, KP, dev)
        self._patchify(frames, A)
        x = torch.empty(Bc, S, D, device=dev, dtype=torch.float32)
        self.linear(A, self.pre + "patch_embed.proj.weight", x, Bc * T * HW, D, KP, epilogue=ops.EPI_RESID,
                    bias=self.p[self.pre + "patch_embed.proj.bias"], map=ops.MAP_PATCH, add_pos=pos, add_time=te,
                    ldo=D, **g)
        ops.cls_init(x, self.p[self.pre + "cls_token"], pos)

        blocks = []
        for i in range(self.depth):
            dp = (drop_scales[i] if drop_scales is not None else None) or {}
            x, saved = self._block_fwd(i, x, Bc, T, HW, dp, save)
            blocks.append(saved)

        feat = torch.empty(Bc, D, device=dev, dtype=torch.float32)
        st_f = torch.empty(Bc, 2, device=dev, dtype=torch.float32)
        ops.layernorm_fwd(x, self.p[self.pre + "norm.weight"], self.p[self.pre + "norm.bias"], feat, st_f, Bc, D,
                          self.eps, ops.MAP_CLS, **g)
        if not save:
            return feat, None
        return feat, dict(Bc=Bc, T=T, HW=HW, A=A, x_last=x, st_f=st_f, blocks=blocks, pos_idx=pos_idx, te_idx=te_idx)

    def _block_fwd(self, i, x0, Bc, T, HW, dp, save):
        """Block.forward, divided_space_time branch (vit.py:128-158)."""
        D, H, Hd = self.D, self.H, self.hidden
        L, S = HW * T, 1 + HW * T
        Mt, Ms, Mm = Bc * L, Bc * T * (HW + 1), Bc * S
        dev = x0.device
        b = f"{self.pre}blocks.{i}."
        P = self.p
        g = dict(T=T, HW=HW)
        f32 = dict(device=dev, dtype=torch.float32)

        # ---- temporal attention over the T frames of each spatial token (vit.py:130-135)
        ln_t, st_t = self._act(Mt, D, dev), torch.empty(Mt, 2, **f32)
        ops.layernorm_fwd(x0, P[b + "temporal_norm1.weight"], P[b + "temporal_norm1.bias"], ln_t, st_t, Mt, D, self.eps,
                          ops.MAP_SKIPCLS, **g)
        qkv_t = self._act(Mt, 3 * D, dev)
        self.linear(ln_t, b + "temporal_attn.qkv.weight", qkv_t, Mt, 3 * D, D, bias=P[b + "temporal_attn.qkv.bias"])
        o_t, lse_t = self._act(Mt, D, dev), torch.empty(Bc * HW, H, T, **f32)
        ops.attn_fwd(qkv_t, o_t, lse_t, Bc * HW, T, H, self.scale)
        p_t = self._act(Mt, D, dev)
        self.linear(o_t, b + "temporal_attn.proj.weight", p_t, Mt, D, D, bias=P[b + "temporal_attn.proj.bias"],
                    rowscale=dp.get("temporal"), rs_div=T)
        x1 = torch.empty(Bc, S, D, **f32)          # token rows only; the cls row of x1 is never read
        self.linear(p_t, b + "temporal_fc.weight", x1, Mt, D, D, epilogue=ops.EPI_RESID, bias=P[b + "temporal_fc.bias"],
                    map=ops.MAP_SKIPCLS, resid=x0, ldo=D, **g)

        # ---- spatial attention over [cls, HW] tokens of each frame (vit.py:138-153)
        ln_s, st_s = self._act(Ms, D, dev), torch.empty(Ms, 2, **f32)
        ops.layernorm_fwd(x1, P[b + "norm1.weight"], P[b + "norm1.bias"], ln_s, st_s, Ms, D, self.eps, ops.MAP_SPATIAL,
                          x_cls=x0, **g)
        qkv_s = self._act(Ms, 3 * D, dev)
        self.linear(ln_s, b + "attn.qkv.weight", qkv_s, Ms, 3 * D, D, bias=P[b + "attn.qkv.bias"])
        o_s, lse_s = self._act(Ms, D, dev), torch.empty(Bc * T, H, HW + 1, **f32)
        self.spatial_attn_fwd(qkv_s, o_s, lse_s, Bc * T, HW + 1)
        x2, side = torch.empty(Bc, S, D, **f32), torch.empty(Bc * T, D, **f32)
        self.linear(o_s, b + "attn.proj.weight", x2, Ms, D, D, epilogue=ops.EPI_RESID, bias=P[b + "attn.proj.bias"],
                    rowscale=dp.get("spatial"), rs_div=HW + 1, map=ops.MAP_SPATIAL, resid=x1, out2=side, ldo=D, **g)
        ops.cls_merge(x0, side, x2, Bc, T, S, D)   # cls = init cls + mean over frames (vit.py:147-149,156)

        # ---- MLP (vit.py:157)
        ln_m, st_m = self._act(Mm, D, dev), torch.empty(Mm, 2, **f32)
        ops.layernorm_fwd(x2, P[b + "norm2.weight"], P[b + "norm2.bias"], ln_m, st_m, Mm, D, self.eps, ops.MAP_IDENT)
        dact, hid = self._act(Mm, Hd, dev), self._act(Mm, Hd, dev)      # gelu'(fc1) (for the backward), gelu(fc1)
        self.linear(ln_m, b + "mlp.fc1.weight", dact, Mm, Hd, D, epilogue=ops.EPI_GELU, bias=P[b + "mlp.fc1.bias"],
                    out2=hid)
        x3 = torch.empty(Bc, S, D, **f32)
        self.linear(hid, b + "mlp.fc2.weight", x3, Mm, D, Hd, epilogue=ops.EPI_RESID, bias=P[b + "mlp.fc2.bias"],
                    rowscale=dp.get("mlp"), rs_div=S, map=ops.MAP_IDENT, resid=x2, ldo=D)
        if not save:
            return x3, None
        return x3, dict(x0=x0, x1=x1, x2=x2, ln_t=ln_t, st_t=st_t, qkv_t=qkv_t, o_t=o_t, lse_t=lse_t, p_t=p_t,
                        ln_s=ln_s, st_s=st_s, qkv_s=qkv_s, o_s=o_s, lse_s=lse_s, ln_m=ln_m, st_m=st_m, dact=dact,
                        hid=hid, dp=dp)

    # spatial attention dispatch (the tcgen05 kernel plugs in here)
    def spatial_attn_fwd(self, qkv, out, lse, n_seq, seq):
        if not self.x3 and seq <= 256 and self.use_tc_attention:
            ops.attn_tc_fwd(qkv, out, lse, n_seq, seq, self.H, self.scale)
        else:
            ops.attn_fwd(qkv, out, lse, n_seq, seq, self.H, self.scale)

    def spatial_attn_bwd(self, qkv, out, dout, lse, dqkv, n_seq, seq):
        if not self.x3 and seq <= 256 and self.use_tc_attention:
            ops.attn_tc_bwd(qkv, out, dout, lse, dqkv, n_seq, seq, self.H, self.scale)
        else:
            ops.attn_bwd(qkv, out, dout, lse, dqkv, n_seq, seq, self.H, self.scale)

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, st, dfeat):
        """dfeat fp32 [Bc, D] -> dict name -> fp32 gradient (views of one flat zero-initialised buffer)."""
        if not self.divided:
            return self._backward_plain(st, dfeat)
        Bc, T, HW = st["Bc"], st["T"], st["HW"]
        D = self.D
        L, S = HW * T, 1 + HW * T
        dev = dfeat.device
        g = dict(T=T, HW=HW)
        P = self.p
        if self.grad_sink is not None:
            G = self.grad_sink                     # kernels accumulate (+=) into the caller's gradient buffers
        else:
            sizes = [P[n].numel() for n in self.grad_names]
            flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
            G, off = {}, 0
            for n, sz in zip(self.grad_names, sizes):
                G[n] = flat[off:off + sz].view(P[n].shape)
                off += sz

        dx = torch.zeros(Bc, S, D, device=dev, dtype=torch.float32)
        ops.layernorm_bwd(dfeat.contiguous(), st["x_last"], P[self.pre + "norm.weight"], st["st_f"], dx,
                          G[self.pre + "norm.weight"], G[self.pre + "norm.bias"], Bc, D, ops.MAP_CLS, **g)
        KP = 3 * self.patch * self.patch
        Mp = Bc * T * HW
        dY = None                                  # block i's MLP operand, emitted by block i+1's last LayerNorm backward
        for i in reversed(range(self.depth)):
            # what this block's last LayerNorm backward hands to its successor in the backward order: the MLP dY of
            # block i-1 (DropPath factors and fc2 bias gradient of THAT block), or the patch-embedding dY below block 0
            if not self.fuse_gather:
                nxt = None
            elif i > 0:
                nxt = (self._act(Bc * S, D, dev), ops.MAP_IDENT, st["blocks"][i - 1]["dp"].get("mlp"), S,
                       G[f"{self.pre}blocks.{i - 1}.mlp.fc2.bias"])
            else:
                nxt = (self._act(Mp, D, dev), ops.MAP_PATCH, None, 0, G[self.pre + "patch_embed.proj.bias"])
            self._block_bwd(i, st["blocks"][i], dx, G, Bc, T, HW, dY, nxt)
            dY = nxt[0] if nxt is not None else None
            st["blocks"][i] = None                 # free this block's activations
            if self.on_block_bwd_done is not None:
                self.on_block_bwd_done(i)

        # embeddings (vit.py:370-407): patch conv, cls token, pos / time embeddings
        dYp = dY
        if dYp is None:
            dYp = self._act(Mp, D, dev)
            ops.gather_cast(dx, dYp, Mp, D, ops.MAP_PATCH, colsum=G[self.pre + "patch_embed.proj.bias"], **g)
        self.linear_dw(dYp, st["A"], G[self.pre + "patch_embed.proj.weight"].view(D, KP), None, Mp, D, KP)
        gpos, gte = G[self.pre + "pos_embed"][0], G[self.pre + "time_embed"][0]
        dpos = gpos if st["pos_idx"] is None else torch.zeros(HW + 1, D, device=dev)
        dte = gte if st["te_idx"] is None else torch.zeros(T, D, device=dev)
        ops.embed_bwd(dx, G[self.pre + "cls_token"].view(D), dpos, dte, Bc, D, T, HW)
        if st["pos_idx"] is not None:
            gpos.index_add_(0, st["pos_idx"], dpos)
        if st["te_idx"] is not None:
            gte.index_add_(0, st["te_idx"], dte)
        return G

    def _block_bwd(self, i, sv, dx, G, Bc, T, HW, dY=None, nxt=None):
        """dY: this block's MLP operand if the previous call already emitted it; nxt = (out, map, rowscale, rs_div, colsum):
        what the temporal LayerNorm backward emits for the next call (see `backward`).  With `fuse_gather` every
        LayerNorm backward writes the next sub-layer's GEMM operand itself (pvrl_layernorm_bwd_emit); otherwise a
        pvrl_gather_cast pass re-reads dx."""
        D, H, Hd = self.D, self.H, self.hidden
        L, S = HW * T, 1 + HW * T
        Mt, Ms, Mm = Bc * L, Bc * T * (HW + 1), Bc * S
        dev = dx.device
        b = f"{self.pre}blocks.{i}."
        P = self.p
        g = dict(T=T, HW=HW)
        dp = sv["dp"]
        fuse = self.fuse_gather

        # ---- MLP: x3 = x2 + s_m * (fc2(gelu(fc1(LN(x2)))))
        if dY is None:
            dY = self._act(Mm, D, dev)
            ops.gather_cast(dx, dY, Mm, D, ops.MAP_IDENT, rowscale=dp.get("mlp"), rs_div=S, colsum=G[b + "mlp.fc2.bias"])
        self.linear_dw(dY, sv["hid"], G[b + "mlp.fc2.weight"], None, Mm, D, Hd)
        d_pre = self._act(Mm, Hd, dev)
        self.linear_dx(dY, b + "mlp.fc2.weight", d_pre, Mm, Hd, D, epilogue=ops.EPI_DGELU, aux=sv["dact"],
                       colsum=G[b + "mlp.fc1.bias"])
        self.linear_dw(d_pre, sv["ln_m"], G[b + "mlp.fc1.weight"], None, Mm, Hd, D)
        d_ln = self._act(Mm, D, dev)
        self.linear_dx(d_pre, b + "mlp.fc1.weight", d_ln, Mm, D, Hd)
        del d_pre, dY
        dYs = self._act(Ms, D, dev)
        emit = (dYs, ops.MAP_SPATIAL, dp.get("spatial"), HW + 1, G[b + "attn.proj.bias"]) if fuse else None
        ops.layernorm_bwd(d_ln, sv["x2"], P[b + "norm2.weight"], sv["st_m"], dx, G[b + "norm2.weight"],
                          G[b + "norm2.bias"], Mm, D, ops.MAP_IDENT, emit=emit, **g)

        # ---- spatial: tokens x2 = x1 + s_s * proj(attn(LN(gather(x0 cls, x1)))), cls = x0 cls + mean_t(...)
        if not fuse:
            ops.gather_cast(dx, dYs, Ms, D, ops.MAP_SPATIAL, rowscale=dp.get("spatial"), rs_div=HW + 1,
                            colsum=G[b + "attn.proj.bias"], **g)
        self.linear_dw(dYs, sv["o_s"], G[b + "attn.proj.weight"], None, Ms, D, D)
        d_o = self._act(Ms, D, dev)
        self.linear_dx(dYs, b + "attn.proj.weight", d_o, Ms, D, D)
        dqkv = self._act(Ms, 3 * D, dev)
        self.spatial_attn_bwd(sv["qkv_s"], sv["o_s"], d_o, sv["lse_s"], dqkv, Bc * T, HW + 1)
        self.linear_dw(dqkv, sv["ln_s"], G[b + "attn.qkv.weight"], G[b + "attn.qkv.bias"], Ms, 3 * D, D)
        d_ln = self._act(Ms, D, dev)
        self.linear_dx(dqkv, b + "attn.qkv.weight", d_ln, Ms, D, 3 * D)
        del dYs
        dYf = self._act(Mt, D, dev)
        emit = (dYf, ops.MAP_SKIPCLS, None, 0, G[b + "temporal_fc.bias"]) if fuse else None
        ops.layernorm_bwd(d_ln, sv["x1"], P[b + "norm1.weight"], sv["st_s"], dx, G[b + "norm1.weight"],
                          G[b + "norm1.bias"], Ms, D, ops.MAP_SPATIAL, x_cls=sv["x0"], emit=emit, **g)

        # ---- temporal: x1 = x0[:,1:] + fc(s_t * proj(attn(LN(x0[:,1:]))))
        if not fuse:
            ops.gather_cast(dx, dYf, Mt, D, ops.MAP_SKIPCLS, colsum=G[b + "temporal_fc.bias"], **g)
        self.linear_dw(dYf, sv["p_t"], G[b + "temporal_fc.weight"], None, Mt, D, D)
        d_p = self._act(Mt, D, dev)
        self.linear_dx(dYf, b + "temporal_fc.weight", d_p, Mt, D, D, rowscale=dp.get("temporal"), rs_div=T,
                       colsum=G[b + "temporal_attn.proj.bias"])
        self.linear_dw(d_p, sv["o_t"], G[b + "temporal_attn.proj.weight"], None, Mt, D, D)
        d_o = self._act(Mt, D, dev)
        self.linear_dx(d_p, b + "temporal_attn.proj.weight", d_o, Mt, D, D)
        dqkv = self._act(Mt, 3 * D, dev)
        ops.attn_bwd(sv["qkv_t"], sv["o_t"], d_o, sv["lse_t"], dqkv, Bc * HW, T, H, self.scale)
        self.linear_dw(dqkv, sv["ln_t"], G[b + "temporal_attn.qkv.weight"], G[b + "temporal_attn.qkv.bias"], Mt, 3 * D, D)
        d_ln = self._act(Mt, D, dev)
        self.linear_dx(dqkv, b + "temporal_attn.qkv.weight", d_ln, Mt, D, 3 * D)
        ops.layernorm_bwd(d_ln, sv["x0"], P[b + "temporal_norm1.weight"], sv["st_t"], dx,
                          G[b + "temporal_norm1.weight"], G[b + "temporal_norm1.bias"], Mt, D, ops.MAP_SKIPCLS,
                          emit=nxt, **g)


# ---------------------------------------------------------------------------------------------- plain ViT schedules
def _grad_buffers(eng, dev):
    P = eng.p
    if eng.grad_sink is not None:
        return eng.grad_sink
    sizes = [P[n].numel() for n in eng.grad_names]
    flat = torch.zeros(sum(sizes), device=dev, dtype=torch.float32)
    G, off = {}, 0
    for n, sz in zip(eng.grad_names, sizes):
        G[n] = flat[off:off + sz].view(P[n].shape)
        off += sz
    return G


def _forward_plain(self, frames, drop_scales, save):
    """TIMESFORMER.ATTENTION_TYPE space_only / joint_space_time (vit.py:124-127, :393, :414-416): every block is
    x += attn(norm1(x)); x += mlp(norm2(x)) over sequences of Sx tokens --
      space_only        Bx = Bc*T sequences of 1 + HW tokens (one per frame, no time embedding), frames averaged at the end;
      joint_space_time  Bx = Bc sequences of 1 + HW*T tokens (same embedding as the divided model)."""
    Bc, C, T, Hh, Ww = frames.shape
    D, P, Hd = self.D, self.patch, self.hidden
    HW = (Hh // P) * (Ww // P)
    so = self.attention_type == "space_only"
    Bx, Tx = (Bc * T, 1) if so else (Bc, T)               # geometry of the residual stream as the kernels see it
    Sx = 1 + HW * Tx
    Mx = Bx * Sx
    dev = frames.device
    g = dict(T=Tx, HW=HW)
    f32 = dict(device=dev, dtype=torch.float32)
    pos, te, pos_idx, te_idx = self._pos_time(HW, T)
    if so:
        te = torch.zeros(1, D, **f32)                     # space_only adds no time embedding (vit.py:393)
    KP = 3 * P * P
    A = self._act(Bc * T * HW, KP, dev)
    self._patchify(frames, A)
    x = torch.empty(Bx, Sx, D, **f32)
    self.linear(A, self.pre + "patch_embed.proj.weight", x, Bc * T * HW, D, KP, epilogue=ops.EPI_RESID,
                bias=self.p[self.pre + "patch_embed.proj.bias"], map=ops.MAP_PATCH, add_pos=pos, add_time=te, ldo=D, **g)
    ops.cls_init(x, self.p[self.pre + "cls_token"], pos)
    Pm = self.p
    blocks = []
    for i in range(self.depth):
        dp = (drop_scales[i] if drop_scales is not None else None) or {}
        b = f"{self.pre}blocks.{i}."
        x0 = x
        ln_a, st_a = self._act(Mx, D, dev), torch.empty(Mx, 2, **f32)
        ops.layernorm_fwd(x0, Pm[b + "norm1.weight"], Pm[b + "norm1.bias"], ln_a, st_a, Mx, D, self.eps, ops.MAP_IDENT)
        qkv = self._act(Mx, 3 * D, dev)
        self.linear(ln_a, b + "attn.qkv.weight", qkv, Mx, 3 * D, D, bias=Pm[b + "attn.qkv.bias"])
        o, lse = self._act(Mx, D, dev), torch.empty(Bx, self.H, Sx, **f32)
        self.spatial_attn_fwd(qkv, o, lse, Bx, Sx)
        x1 = torch.empty(Bx, Sx, D, **f32)
        self.linear(o, b + "attn.proj.weight", x1, Mx, D, D, epilogue=ops.EPI_RESID, bias=Pm[b + "attn.proj.bias"],
                    rowscale=dp.get("attn"), rs_div=Sx, map=ops.MAP_IDENT, resid=x0, ldo=D)
        ln_m, st_m = self._act(Mx, D, dev), torch.empty(Mx, 2, **f32)
        ops.layernorm_fwd(x1, Pm[b + "norm2.weight"], Pm[b + "norm2.bias"], ln_m, st_m, Mx, D, self.eps, ops.MAP_IDENT)
        dact, hid = self._act(Mx, Hd, dev), self._act(Mx, Hd, dev)
        self.linear(ln_m, b + "mlp.fc1.weight", dact, Mx, Hd, D, epilogue=ops.EPI_GELU, bias=Pm[b + "mlp.fc1.bias"],
                    out2=hid)
        x = torch.empty(Bx, Sx, D, **f32)
        self.linear(hid, b + "mlp.fc2.weight", x, Mx, D, Hd, epilogue=ops.EPI_RESID, bias=Pm[b + "mlp.fc2.bias"],
                    rowscale=dp.get("mlp"), rs_div=Sx, map=ops.MAP_IDENT, resid=x1, ldo=D)
        blocks.append(dict(x0=x0, x1=x1, ln_a=ln_a, st_a=st_a, qkv=qkv, o=o, lse=lse, ln_m=ln_m, st_m=st_m, dact=dact,
                           hid=hid, dp=dp) if save else None)
    # final norm on the cls rows; space_only first averages the frames of a clip (vit.py:414-418)
    feat = torch.empty(Bc, D, **f32)
    st_f = torch.empty(Bc, 2, **f32)
    if so:
        xm = x.view(Bc, T, Sx, D)[:, :, 0].mean(1).contiguous()
        ops.layernorm_fwd(xm, Pm[self.pre + "norm.weight"], Pm[self.pre + "norm.bias"], feat, st_f, Bc, D, self.eps,
                          ops.MAP_IDENT)
    else:
        xm = x
        ops.layernorm_fwd(x, Pm[self.pre + "norm.weight"], Pm[self.pre + "norm.bias"], feat, st_f, Bc, D, self.eps,
                          ops.MAP_CLS, **g)
    if not save:
        return feat, None
    return feat, dict(Bc=Bc, T=T, HW=HW, A=A, xm=xm, st_f=st_f, blocks=blocks, pos_idx=pos_idx, te_idx=te_idx)


def _backward_plain(self, st, dfeat):
    Bc, T, HW = st["Bc"], st["T"], st["HW"]
    D, Hd = self.D, self.hidden
    so = self.attention_type == "space_only"
    Bx, Tx = (Bc * T, 1) if so else (Bc, T)
    Sx = 1 + HW * Tx
    Mx = Bx * Sx
    dev = dfeat.device
    g = dict(T=Tx, HW=HW)
    P = self.p
    G = _grad_buffers(self, dev)
    dx = torch.zeros(Bx, Sx, D, device=dev, dtype=torch.float32)
    if so:
        dxm = torch.zeros(Bc, D, device=dev, dtype=torch.float32)
        ops.layernorm_bwd(dfeat.contiguous(), st["xm"], P[self.pre + "norm.weight"], st["st_f"], dxm,
                          G[self.pre + "norm.weight"], G[self.pre + "norm.bias"], Bc, D, ops.MAP_IDENT)
        dx.view(Bc, T, Sx, D)[:, :, 0] = (dxm / T).unsqueeze(1)          # backward of the mean over frames
    else:
        ops.layernorm_bwd(dfeat.contiguous(), st["xm"], P[self.pre + "norm.weight"], st["st_f"], dx,
                          G[self.pre + "norm.weight"], G[self.pre + "norm.bias"], Bc, D, ops.MAP_CLS, **g)
    for i in reversed(range(self.depth)):
        sv = st["blocks"][i]
        b = f"{self.pre}blocks.{i}."
        dp = sv["dp"]
        # ---- MLP
        dY = self._act(Mx, D, dev)
        ops.gather_cast(dx, dY, Mx, D, ops.MAP_IDENT, rowscale=dp.get("mlp"), rs_div=Sx, colsum=G[b + "mlp.fc2.bias"])
        self.linear_dw(dY, sv["hid"], G[b + "mlp.fc2.weight"], None, Mx, D, Hd)
        d_pre = self._act(Mx, Hd, dev)
        self.linear_dx(dY, b + "mlp.fc2.weight", d_pre, Mx, Hd, D, epilogue=ops.EPI_DGELU, aux=sv["dact"],
                       colsum=G[b + "mlp.fc1.bias"])
        self.linear_dw(d_pre, sv["ln_m"], G[b + "mlp.fc1.weight"], None, Mx, Hd, D)
        d_ln = self._act(Mx, D, dev)
        self.linear_dx(d_pre, b + "mlp.fc1.weight", d_ln, Mx, D, Hd)
        del d_pre
        ops.layernorm_bwd(d_ln, sv["x1"], P[b + "norm2.weight"], sv["st_m"], dx, G[b + "norm2.weight"],
                          G[b + "norm2.bias"], Mx, D, ops.MAP_IDENT)
        # ---- attention
        dYa = self._act(Mx, D, dev)
        ops.gather_cast(dx, dYa, Mx, D, ops.MAP_IDENT, rowscale=dp.get("attn"), rs_div=Sx, colsum=G[b + "attn.proj.bias"])
        self.linear_dw(dYa, sv["o"], G[b + "attn.proj.weight"], None, Mx, D, D)
        d_o = self._act(Mx, D, dev)
        self.linear_dx(dYa, b + "attn.proj.weight", d_o, Mx, D, D)
        dqkv = self._act(Mx, 3 * D, dev)
        self.spatial_attn_bwd(sv["qkv"], sv["o"], d_o, sv["lse"], dqkv, Bx, Sx)
        self.linear_dw(dqkv, sv["ln_a"], G[b + "attn.qkv.weight"], G[b + "attn.qkv.bias"], Mx, 3 * D, D)
        d_ln = self._act(Mx, D, dev)
        self.linear_dx(dqkv, b + "attn.qkv.weight", d_ln, Mx, D, 3 * D)
        ops.layernorm_bwd(d_ln, sv["x0"], P[b + "norm1.weight"], sv["st_a"], dx, G[b + "norm1.weight"],
                          G[b + "norm1.bias"], Mx, D, ops.MAP_IDENT)
        st["blocks"][i] = None
    # embeddings
    KP = 3 * self.patch * self.patch
    Mp = Bc * T * HW
    dYp = self._act(Mp, D, dev)
    ops.gather_cast(dx, dYp, Mp, D, ops.MAP_PATCH, colsum=G[self.pre + "patch_embed.proj.bias"], **g)
    self.linear_dw(dYp, st["A"], G[self.pre + "patch_embed.proj.weight"].view(D, KP), None, Mp, D, KP)
    gpos = G[self.pre + "pos_embed"][0]
    gte = None if so else G[self.pre + "time_embed"][0]
    dpos = gpos if st["pos_idx"] is None else torch.zeros(HW + 1, D, device=dev)
    dte = None if so else (gte if st["te_idx"] is None else torch.zeros(T, D, device=dev))
    ops.embed_bwd(dx, G[self.pre + "cls_token"].view(D), dpos, dte, Bx, D, Tx, HW)
    if st["pos_idx"] is not None:
        gpos.index_add_(0, st["pos_idx"], dpos)
    if dte is not None and st["te_idx"] is not None:
        gte.index_add_(0, st["te_idx"], dte)
    return G


EncoderEngine._forward_plain = _forward_plain
EncoderEngine._backward_plain = _backward_plain


class EncoderFunction(torch.autograd.Function):
    """autograd boundary: forward_features as one node whose backward is EncoderEngine.backward."""

    @staticmethod
    def forward(ctx, engine, frames, drop_scales, need, *params):
        feat, saved = engine.forward(frames, drop_scales, save=need)
        ctx.engine, ctx.saved = engine, saved
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        eng = ctx.engine
        if ctx.saved is None:
            raise RuntimeError("EncoderFunction.backward called without saved activations")
        G = eng.backward(ctx.saved, dfeat)
        ctx.saved = None
        if eng.grad_sink is not None:              # already accumulated in place
            return (None, None, None, None) + (None,) * len(eng.grad_names)
        return (None, None, None, None) + tuple(G[n] for n in eng.grad_names)


def encode(engine, frames, drop_scales=None):
    """cls feature [Bc, D] with autograd through the engine's parameters."""
    params = [engine.p[n] for n in engine.grad_names]
    need = torch.is_grad_enabled() and any(p.requires_grad for p in params)   # Function.forward runs under no_grad
    return EncoderFunction.apply(engine, frames, drop_scales, need, *params)
