"""Lane-level CPU model of csrc/mvit_attn_mma.cu (pooled attention forward on mma.sync.m16n8k16 + ldmatrix.x4.trans).

The kernel's arithmetic is trivial; what can be wrong is its *index algebra* -- which element of Q / K / V / S / P / O lives
in which register of which lane.  This file restates the PTX fragment layouts of `mma.sync.aligned.m16n8k16.row.col`
(A: a0a1 = (row g, cols 2t, 2t+1), a2a3 = (row g+8, ..), a4a5 = (row g, cols 2t+8, +9), a6a7 = (row g+8, ..);
B: b0b1 = (rows 2t, 2t+1, col g), b2b3 = (rows 2t+8, +9, col g); C: c0c1 = (row g, cols 2t, 2t+1), c2c3 = (row g+8, ..);
g = lane >> 2, t = lane & 3) and of `ldmatrix.m8n8.x4.trans` (lane l supplies the address of row l & 7 of matrix l >> 3;
register i of a lane = elements (2t, g), (2t+1, g) of matrix i), then executes the kernel's addressing with them, warp by
warp, and compares with the plain attention of tests/shadow_ops.py.  A wrong fragment index, chunk tail, mask or bias column
shows up here without a GPU (the same procedure preceded the first GPU run of attention_t32.cu)."""
import numpy as np
import pytest
import torch

import shadow_ops as S

KC, C = 64, 96            # MM_KC, MM_C of mvit_attn_mma.cu


def bf16(x):
    return torch.from_numpy(np.asarray(x, dtype=np.float32)).to(torch.bfloat16).float().numpy()


def mma16816(c, a, b0, b1):
    """c [32, 4] += A B with A (16 x 16) and B (16 x 8) assembled from the lanes' fragments."""
    A, Bm = np.zeros((16, 16), np.float32), np.zeros((16, 8), np.float32)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, 2 * t:2 * t + 2], A[g + 8, 2 * t:2 * t + 2] = a[lane, 0], a[lane, 1]
        A[g, 2 * t + 8:2 * t + 10], A[g + 8, 2 * t + 8:2 * t + 10] = a[lane, 2], a[lane, 3]
        Bm[2 * t:2 * t + 2, g], Bm[2 * t + 8:2 * t + 10, g] = b0[lane], b1[lane]
    Cm = A.astype(np.float64) @ Bm.astype(np.float64)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        c[lane] += np.array([Cm[g, 2 * t], Cm[g, 2 * t + 1], Cm[g + 8, 2 * t], Cm[g + 8, 2 * t + 1]], np.float32)


def ldsm_x4_trans(smem, addr):
    """addr(lane) -> (row, col) of the 8-element shared-memory row that lane points at.  Returns regs [4][32, 2]."""
    regs = []
    for m in range(4):
        M = np.stack([smem[addr(8 * m + r)[0], addr(8 * m + r)[1]:addr(8 * m + r)[1] + 8] for r in range(8)])
        regs.append(np.stack([np.array([M[2 * (l & 3), l >> 2], M[2 * (l & 3) + 1, l >> 2]], np.float32) for l in range(32)]))
    return regs


def quad(v, op):
    out = v.copy()
    for mask in (1, 2):
        out = op(out, out[np.arange(32) ^ mask])
    return out


def emulate_warp(q, k, v, bq, bh, block_x, wid, heads, kg, scale, resid, out, lse):
    """One warp of pooled_attn_fwd_mma_kernel: q [BH, Nq, C], k / v [BH, Nk, C] (bf16 values), bq [BH, Nq - 1, KB]."""
    Nq, Nk = q.shape[1], k.shape[1]
    Kt, Kh, Kw = kg
    lanes = np.arange(32)
    gq, t = lanes >> 2, lanes & 3
    r0 = block_x * 64 + wid * 16 + gq
    r1 = r0 + 8
    qa = np.zeros((C // 16, 32, 4, 2), np.float32)
    for kk in range(C // 16):
        for lane in range(32):
            d = kk * 16 + 2 * t[lane]
            if r0[lane] < Nq:
                qa[kk, lane, 0], qa[kk, lane, 2] = q[bh, r0[lane], d:d + 2], q[bh, r0[lane], d + 8:d + 10]
            if r1[lane] < Nq:
                qa[kk, lane, 1], qa[kk, lane, 3] = q[bh, r1[lane], d:d + 2], q[bh, r1[lane], d + 8:d + 10]
    m0 = np.full(32, -3.0e38, np.float32)
    m1, l0, l1 = m0.copy(), np.zeros(32, np.float32), np.zeros(32, np.float32)
    o = np.zeros((C // 8, 32, 4), np.float32)
    for j0 in range(0, Nk, KC):
        ks, vs = np.zeros((KC, 104), np.float32), np.zeros((KC, 104), np.float32)
        kcomp = np.full(KC, -1, np.int64)
        for row in range(KC):
            j = j0 + row
            if j < Nk:
                ks[row, :C], vs[row, :C] = k[bh, j], v[bh, j]
            if 0 < j < Nk:
                jj = j - 1
                kcomp[row] = (jj // (Kw * Kh)) | ((Kt + (jj // Kw) % Kh) << 8) | ((Kt + Kh + jj % Kw) << 16)
        s = np.zeros((KC // 8, 32, 4), np.float32)
        for n in range(KC // 8):
            for kk in range(C // 16):
                b0 = np.stack([ks[n * 8 + gq[l], kk * 16 + 2 * t[l]:kk * 16 + 2 * t[l] + 2] for l in range(32)])
                b1 = np.stack([ks[n * 8 + gq[l], kk * 16 + 8 + 2 * t[l]:kk * 16 + 10 + 2 * t[l]] for l in range(32)])
                mma16816(s[n], qa[kk], b0, b1)
        mx0, mx1 = np.full(32, -3.0e38, np.float32), np.full(32, -3.0e38, np.float32)
        for n in range(KC // 8):
            for e in range(2):
                for lane in range(32):
                    jl = n * 8 + 2 * t[lane] + e
                    c = kcomp[jl]
                    a0, a1 = s[n, lane, e] * scale, s[n, lane, 2 + e] * scale
                    if c >= 0:
                        cols = [c & 0xff, (c >> 8) & 0xff, (c >> 16) & 0xff]
                        if 0 < r0[lane] < Nq:
                            a0 += bq[bh, r0[lane] - 1, cols].sum()
                        if 0 < r1[lane] < Nq:
                            a1 += bq[bh, r1[lane] - 1, cols].sum()
                    if j0 + jl >= Nk:
                        a0 = a1 = -3.0e38
                    s[n, lane, e], s[n, lane, 2 + e] = a0, a1
                    mx0[lane], mx1[lane] = max(mx0[lane], a0), max(mx1[lane], a1)
        mn0, mn1 = np.maximum(m0, quad(mx0, np.maximum)), np.maximum(m1, quad(mx1, np.maximum))
        with np.errstate(under="ignore"):
            corr0, corr1 = np.exp(m0 - mn0), np.exp(m1 - mn1)
            m0, m1 = mn0, mn1
            s[:, :, 0:2] = np.exp(s[:, :, 0:2] - mn0[None, :, None])
            s[:, :, 2:4] = np.exp(s[:, :, 2:4] - mn1[None, :, None])
        l0 = l0 * corr0 + s[:, :, 0:2].sum(axis=(0, 2))
        l1 = l1 * corr1 + s[:, :, 2:4].sum(axis=(0, 2))
        o[:, :, 0:2] *= corr0[None, :, None]
        o[:, :, 2:4] *= corr1[None, :, None]
        for kk in range(KC // 16):
            pa = np.zeros((32, 4, 2), np.float32)
            pa[:, 0], pa[:, 1] = bf16(s[2 * kk][:, 0:2]), bf16(s[2 * kk][:, 2:4])
            pa[:, 2], pa[:, 3] = bf16(s[2 * kk + 1][:, 0:2]), bf16(s[2 * kk + 1][:, 2:4])
            for npair in range(C // 16):
                regs = ldsm_x4_trans(vs, lambda l: (kk * 16 + ((l >> 3) & 1) * 8 + (l & 7), npair * 16 + ((l >> 3) >> 1) * 8))
                mma16816(o[2 * npair], pa, regs[0], regs[1])
                mma16816(o[2 * npair + 1], pa, regs[2], regs[3])
    l0, l1 = quad(l0, np.add), quad(l1, np.add)
    b, h = bh // heads, bh % heads
    for n in range(C // 8):
        for lane in range(32):
            d = n * 8 + 2 * t[lane]
            for r, lo, l in ((r0[lane], 0, l0[lane]), (r1[lane], 2, l1[lane])):
                if r < Nq:
                    val = o[n, lane, lo:lo + 2] / l
                    if resid and r > 0:
                        val = val + q[bh, r, d:d + 2]
                    out[b, r, h * C + d:h * C + d + 2] = bf16(val)
    for lane in range(0, 32, 4):
        if r0[lane] < Nq:
            lse[bh, r0[lane]] = m0[lane] + np.log(l0[lane])
        if r1[lane] < Nq:
            lse[bh, r1[lane]] = m1[lane] + np.log(l1[lane])


CASES = [  # B, heads, q grid, k grid, residual pooling
    (1, 2, (1, 5, 5), (1, 3, 3), True),          # Nq = 26 (tail rows in the second warp), Nk = 10 (one partial chunk)
    (1, 1, (2, 6, 6), (2, 7, 7), True),          # Nq = 73 (two blocks), Nk = 99 (two chunks, tail of 35)
    (2, 1, (1, 4, 4), (2, 8, 4), False),         # Nk = 65: a chunk holding one key; no residual pooling
]


@pytest.mark.parametrize("B,heads,qg,kg,resid", CASES)
def test_mma_forward_index_algebra(B, heads, qg, kg, resid):
    gen = torch.Generator().manual_seed(sum(qg) * 7 + sum(kg))
    Nq, Nk, KB = 1 + qg[0] * qg[1] * qg[2], 1 + kg[0] * kg[1] * kg[2], sum(kg)
    q, k, v = (torch.randn(B, heads, n, C, generator=gen).to(torch.bfloat16) for n in (Nq, Nk, Nk))
    bq = 0.5 * torch.randn(B, heads, Nq - 1, KB, generator=gen)
    scale = C ** -0.5
    ref, ref_lse = S._pooled_attn(q.float(), k.float(), v.float(), bq, kg, scale, resid)
    out = np.full((B, Nq, heads * C), np.nan, np.float32)
    lse = np.full((B * heads, Nq), np.nan, np.float32)
    qn, kn, vn = (x.float().reshape(B * heads, -1, C).numpy() for x in (q, k, v))
    bqn = bq.reshape(B * heads, Nq - 1, KB).numpy()
    for bh in range(B * heads):
        for block_x in range((Nq + 63) // 64):
            for wid in range(4):
                emulate_warp(qn, kn, vn, bqn, bh, block_x, wid, heads, kg, scale, resid, out, lse)
    assert not np.isnan(out).any() and not np.isnan(lse).any(), "an output element was never written"
    err = np.abs(out - ref.numpy()).max() / np.abs(ref.numpy()).max()
    assert err < 1.5e-2, err                                   # P and the output are rounded to bf16
    assert np.abs(lse.reshape(B, heads, Nq) - ref_lse.numpy()).max() < 1e-4
