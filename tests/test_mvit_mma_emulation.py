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


# ------------------------------------------------------------------------------------------------ backward kernels
BC = 32      # MB_C: keys (dQ pass) / queries (dK/dV pass) per chunk


def afrag(rows0, rows1, t, valid0, valid1):
    """load_afrag: [C // 16][32 lanes][4 regs][2] from the two rows each lane owns (zeros where the row does not exist)."""
    a = np.zeros((C // 16, 32, 4, 2), np.float32)
    for kk in range(C // 16):
        for lane in range(32):
            d = kk * 16 + 2 * t[lane]
            if valid0[lane]:
                a[kk, lane, 0], a[kk, lane, 2] = rows0[lane][d:d + 2], rows0[lane][d + 8:d + 10]
            if valid1[lane]:
                a[kk, lane, 1], a[kk, lane, 3] = rows1[lane][d:d + 2], rows1[lane][d + 8:d + 10]
    return a


def bfrag(smem, n, kk, gq, t):
    b0 = np.stack([smem[n * 8 + gq[l], kk * 16 + 2 * t[l]:kk * 16 + 2 * t[l] + 2] for l in range(32)])
    b1 = np.stack([smem[n * 8 + gq[l], kk * 16 + 8 + 2 * t[l]:kk * 16 + 10 + 2 * t[l]] for l in range(32)])
    return b0, b1


def pack_a(x, kk):
    pa = np.zeros((32, 4, 2), np.float32)
    pa[:, 0], pa[:, 1] = bf16(x[2 * kk][:, 0:2]), bf16(x[2 * kk][:, 2:4])
    pa[:, 2], pa[:, 3] = bf16(x[2 * kk + 1][:, 0:2]), bf16(x[2 * kk + 1][:, 2:4])
    return pa


def ldsm_addr(kk, npair):
    return lambda l: (kk * 16 + ((l >> 3) & 1) * 8 + (l & 7), npair * 16 + ((l >> 3) >> 1) * 8)


def emulate_bwd_q_warp(q, k, v, bq, do, lse, bh, block_x, wid, kg, scale, resid, dq, dbq, delta):
    """One warp of pooled_attn_bwd_q_mma_kernel.  do [BH, Nq, C] is the (clip, head) slice of dout."""
    Nq, Nk = q.shape[1], k.shape[1]
    Kt, Kh, Kw = kg
    KB = Kt + Kh + Kw
    lanes = np.arange(32)
    gq, t = lanes >> 2, lanes & 3
    r0 = block_x * 64 + wid * 16 + gq
    r1 = r0 + 8
    rows = lambda x, r: [x[bh, min(ri, Nq - 1)] for ri in r]
    qa = afrag(rows(q, r0), rows(q, r1), t, r0 < Nq, r1 < Nq)
    da = afrag(rows(do, r0), rows(do, r1), t, r0 < Nq, r1 < Nq)
    ls0 = np.array([lse[bh, r] if r < Nq else 0.0 for r in r0], np.float32)
    ls1 = np.array([lse[bh, r] if r < Nq else 0.0 for r in r1], np.float32)
    nbt = (KB + 7) // 8
    dl0, dl1 = np.zeros(32, np.float32), np.zeros(32, np.float32)
    dqa, dba = np.zeros((C // 8, 32, 4), np.float32), np.zeros((8, 32, 4), np.float32)
    for pas in range(2):
        for j0 in range(0, Nk, BC):
            ks, vs, sel = np.zeros((BC, 104), np.float32), np.zeros((BC, 104), np.float32), np.zeros((BC, 72), np.float32)
            kcomp = np.full(BC, -1, np.int64)
            for row in range(BC):
                j = j0 + row
                if j < Nk:
                    ks[row, :C], vs[row, :C] = k[bh, j], v[bh, j]
                if 0 < j < Nk:
                    jj = j - 1
                    cols = (jj // (Kw * Kh), Kt + (jj // Kw) % Kh, Kt + Kh + jj % Kw)
                    kcomp[row] = cols[0] | (cols[1] << 8) | (cols[2] << 16)
                    if pas == 1:
                        sel[row, list(cols)] = 1.0
            s, dp = np.zeros((BC // 8, 32, 4), np.float32), np.zeros((BC // 8, 32, 4), np.float32)
            for n in range(BC // 8):
                for kk in range(C // 16):
                    mma16816(s[n], qa[kk], *bfrag(ks, n, kk, gq, t))
                    mma16816(dp[n], da[kk], *bfrag(vs, n, kk, gq, t))
            for n in range(BC // 8):
                for e in range(2):
                    for lane in range(32):
                        jl = n * 8 + 2 * t[lane] + e
                        cc = kcomp[jl]
                        a0, a1 = s[n, lane, e] * scale, s[n, lane, 2 + e] * scale
                        if cc >= 0:
                            cols = [cc & 0xff, (cc >> 8) & 0xff, (cc >> 16) & 0xff]
                            if 0 < r0[lane] < Nq:
                                a0 += bq[bh, r0[lane] - 1, cols].sum()
                            if 0 < r1[lane] < Nq:
                                a1 += bq[bh, r1[lane] - 1, cols].sum()
                        live = j0 + jl < Nk
                        p0 = np.exp(a0 - ls0[lane]) if live and r0[lane] < Nq else 0.0
                        p1 = np.exp(a1 - ls1[lane]) if live and r1[lane] < Nq else 0.0
                        if pas == 0:
                            dl0[lane] += p0 * dp[n, lane, e]
                            dl1[lane] += p1 * dp[n, lane, 2 + e]
                        else:
                            s[n, lane, e] = p0 * (dp[n, lane, e] - dl0[lane])
                            s[n, lane, 2 + e] = p1 * (dp[n, lane, 2 + e] - dl1[lane])
            if pas == 1:
                for kk in range(BC // 16):
                    pa = pack_a(s, kk)
                    for npair in range(C // 16):
                        regs = ldsm_x4_trans(ks, ldsm_addr(kk, npair))
                        mma16816(dqa[2 * npair], pa, regs[0], regs[1])
                        mma16816(dqa[2 * npair + 1], pa, regs[2], regs[3])
                    for nb in range(4):
                        if 2 * nb < nbt:
                            regs = ldsm_x4_trans(sel, ldsm_addr(kk, nb))
                            mma16816(dba[2 * nb], pa, regs[0], regs[1])
                            mma16816(dba[2 * nb + 1], pa, regs[2], regs[3])
        if pas == 0:
            dl0, dl1 = quad(dl0, np.add), quad(dl1, np.add)
            for lane in range(0, 32, 4):
                if r0[lane] < Nq:
                    delta[bh, r0[lane]] = dl0[lane]
                if r1[lane] < Nq:
                    delta[bh, r1[lane]] = dl1[lane]
    for n in range(C // 8):
        for lane in range(32):
            d = n * 8 + 2 * t[lane]
            for r, lo in ((r0[lane], 0), (r1[lane], 2)):
                if r < Nq:
                    val = dqa[n, lane, lo:lo + 2] * scale
                    if resid and r > 0:
                        val = val + do[bh, r, d:d + 2]
                    dq[bh, r, d:d + 2] = bf16(val)
    for n in range(8):
        for e in range(2):
            for lane in range(32):
                col = n * 8 + 2 * t[lane] + e
                if col < KB:
                    if 0 < r0[lane] < Nq:
                        dbq[bh, r0[lane] - 1, col] = dba[n, lane, e]
                    if 0 < r1[lane] < Nq:
                        dbq[bh, r1[lane] - 1, col] = dba[n, lane, 2 + e]


def emulate_bwd_kv_warp(q, k, v, bq, do, lse, delta, bh, block_x, wid, zslice, qsplit, kg, scale, dk, dv):
    """One warp of pooled_attn_bwd_kv_mma_kernel for query slice `zslice` of `qsplit`; dk / dv accumulate (the atomics)."""
    Nq, Nk = q.shape[1], k.shape[1]
    Kt, Kh, Kw = kg
    lanes = np.arange(32)
    gq, t = lanes >> 2, lanes & 3
    j_0 = block_x * 64 + wid * 16 + gq
    j_1 = j_0 + 8
    rows = lambda x, r: [x[bh, min(ri, Nk - 1)] for ri in r]
    ka = afrag(rows(k, j_0), rows(k, j_1), t, j_0 < Nk, j_1 < Nk)
    va = afrag(rows(v, j_0), rows(v, j_1), t, j_0 < Nk, j_1 < Nk)

    def comps(j):
        if not 0 < j < Nk:
            return None
        jj = j - 1
        return [jj // (Kw * Kh), Kt + (jj // Kw) % Kh, Kt + Kh + jj % Kw]
    c0, c1 = [comps(j) for j in j_0], [comps(j) for j in j_1]
    dka, dva = np.zeros((C // 8, 32, 4), np.float32), np.zeros((C // 8, 32, 4), np.float32)
    per = (((Nq + qsplit - 1) // qsplit) + BC - 1) // BC * BC
    ibeg, iend = zslice * per, min(Nq, zslice * per + per)
    for i0 in range(ibeg, iend, BC):
        qs, dos = np.zeros((BC, 104), np.float32), np.zeros((BC, 104), np.float32)
        lss, dls = np.zeros(BC, np.float32), np.zeros(BC, np.float32)
        for row in range(BC):
            if i0 + row < iend:
                qs[row, :C], dos[row, :C] = q[bh, i0 + row], do[bh, i0 + row]
                lss[row], dls[row] = lse[bh, i0 + row], delta[bh, i0 + row]
        st, dpt = np.zeros((BC // 8, 32, 4), np.float32), np.zeros((BC // 8, 32, 4), np.float32)
        for n in range(BC // 8):
            for kk in range(C // 16):
                mma16816(st[n], ka[kk], *bfrag(qs, n, kk, gq, t))
                mma16816(dpt[n], va[kk], *bfrag(dos, n, kk, gq, t))
        for n in range(BC // 8):
            for e in range(2):
                for lane in range(32):
                    il = n * 8 + 2 * t[lane] + e
                    i = i0 + il
                    a0, a1 = st[n, lane, e] * scale, st[n, lane, 2 + e] * scale
                    if 0 < i < iend:
                        if c0[lane] is not None:
                            a0 += bq[bh, i - 1, c0[lane]].sum()
                        if c1[lane] is not None:
                            a1 += bq[bh, i - 1, c1[lane]].sum()
                    p0 = np.exp(a0 - lss[il]) if i < iend and j_0[lane] < Nk else 0.0
                    p1 = np.exp(a1 - lss[il]) if i < iend and j_1[lane] < Nk else 0.0
                    d0, d1 = p0 * (dpt[n, lane, e] - dls[il]), p1 * (dpt[n, lane, 2 + e] - dls[il])
                    st[n, lane, e], st[n, lane, 2 + e] = p0, p1
                    dpt[n, lane, e], dpt[n, lane, 2 + e] = d0, d1
        for kk in range(BC // 16):
            pa, sa = pack_a(st, kk), pack_a(dpt, kk)
            for npair in range(C // 16):
                regs = ldsm_x4_trans(dos, ldsm_addr(kk, npair))
                mma16816(dva[2 * npair], pa, regs[0], regs[1])
                mma16816(dva[2 * npair + 1], pa, regs[2], regs[3])
                regs = ldsm_x4_trans(qs, ldsm_addr(kk, npair))
                mma16816(dka[2 * npair], sa, regs[0], regs[1])
                mma16816(dka[2 * npair + 1], sa, regs[2], regs[3])
    for n in range(C // 8):
        for lane in range(32):
            d = n * 8 + 2 * t[lane]
            for j, lo in ((j_0[lane], 0), (j_1[lane], 2)):
                if j < Nk:
                    dk[bh, j, d:d + 2] += dka[n, lane, lo:lo + 2] * scale
                    dv[bh, j, d:d + 2] += dva[n, lane, lo:lo + 2]


@pytest.mark.parametrize("B,heads,qg,kg,resid", CASES)
def test_mma_backward_index_algebra(B, heads, qg, kg, resid):
    gen = torch.Generator().manual_seed(sum(qg) * 11 + sum(kg))
    Nq, Nk, KB = 1 + qg[0] * qg[1] * qg[2], 1 + kg[0] * kg[1] * kg[2], sum(kg)
    BH = B * heads
    q, k, v = (torch.randn(B, heads, n, C, generator=gen).to(torch.bfloat16) for n in (Nq, Nk, Nk))
    bq = 0.5 * torch.randn(B, heads, Nq - 1, KB, generator=gen)
    dout = torch.randn(B, Nq, heads * C, generator=gen).to(torch.bfloat16)
    scale = C ** -0.5
    # reference: autograd of the plain attention (tests/shadow_ops.py)
    lse_t = torch.empty(B, heads, Nq)
    S.pooled_attn_fwd(q, k, v, bq, torch.empty(B, Nq, heads * C), lse_t, kg, scale, resid)
    rdq = torch.empty(B, heads, Nq, C)
    rdk, rdv = torch.zeros(B, heads, Nk, C), torch.zeros(B, heads, Nk, C)
    rdbq, rdelta = torch.empty_like(bq), torch.empty(B, heads, Nq)
    S.pooled_attn_bwd(q, k, v, bq, dout, lse_t, rdq, rdk, rdv, rdbq, rdelta, kg, scale, resid)

    qn, kn, vn = (x.float().reshape(BH, -1, C).numpy() for x in (q, k, v))
    bqn, lsn = bq.reshape(BH, Nq - 1, KB).numpy(), lse_t.reshape(BH, Nq).numpy()
    don = dout.float().reshape(B, Nq, heads, C).permute(0, 2, 1, 3).reshape(BH, Nq, C).numpy()
    dq = np.full((BH, Nq, C), np.nan, np.float32)
    dbq = np.full((BH, Nq - 1, KB), np.nan, np.float32)
    delta = np.full((BH, Nq), np.nan, np.float32)
    dk, dv = np.zeros((BH, Nk, C), np.float32), np.zeros((BH, Nk, C), np.float32)
    for bh in range(BH):
        for bx in range((Nq + 63) // 64):
            for wid in range(4):
                emulate_bwd_q_warp(qn, kn, vn, bqn, don, lsn, bh, bx, wid, kg, scale, resid, dq, dbq, delta)
    assert not (np.isnan(dq).any() or np.isnan(dbq).any() or np.isnan(delta).any()), "an output element was never written"
    qsplit = 2                      # two query slices: their partial sums meet in dk / dv
    for bh in range(BH):
        for bx in range((Nk + 63) // 64):
            for wid in range(4):
                for z in range(qsplit):
                    emulate_bwd_kv_warp(qn, kn, vn, bqn, don, lsn, delta, bh, bx, wid, z, qsplit, kg, scale, dk, dv)

    def relerr(a, r):
        r = r.reshape(a.shape).numpy()
        return np.abs(a - r).max() / (np.abs(r).max() + 1e-12)
    assert relerr(delta, rdelta) < 1e-4, relerr(delta, rdelta)           # fp32 end to end
    for name, a, r in (("dq", dq, rdq), ("dbq", dbq, rdbq), ("dk", dk, rdk), ("dv", dv, rdv)):
        assert relerr(a, r) < 2e-2, (name, relerr(a, r))                 # dS / P rounded to bf16 before their MMAs
