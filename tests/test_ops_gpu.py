"""GPU unit tests: every C-ABI op of libpvrl_sm100.so against a plain fp32 torch restatement of the same op
(run on the B200 with `-m gpu`).  Tolerances are stated per test: bf16-operand GEMMs are compared with
the exact product of the bf16-rounded operands (so only accumulation order differs), fp32 ops to ~1e-5."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from procedurevrl_b200 import ops as O
    O.lib()
    return O


def _rand(*shape, scale=1.0, seed=0, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


def _report(name, got, ref):
    err = (got.float() - ref.float()).abs()
    rel = err.max().item() / (ref.float().abs().max().item() + 1e-12)
    print(f"[{name}] max_abs_err={err.max().item():.3e} rel_to_max={rel:.3e} ref_absmax={ref.abs().max().item():.3e}")
    return err, rel


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 192), (1000, 768, 768), (4097, 2304, 768),
                                   (300, 768, 3072), (130, 32, 72)])
def test_gemm_nt_store(ops, M, N, K):
    A = _rand(M, K, seed=1, dtype=torch.bfloat16)
    B = _rand(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    bias = _rand(N, seed=3)
    ref = A.float() @ B.float().t() + bias
    for dt in (torch.float32, torch.bfloat16):
        out = torch.full((M, N), float("nan"), device="cuda", dtype=dt)
        ops.gemm(A, B, out, M=M, N=N, K=K, bias=bias)
        torch.cuda.synchronize()
        err, rel = _report(f"gemm_nt {M}x{N}x{K} {dt}", out, ref)
        if rel > 1e-2:  # help debugging a wrong descriptor: where are the errors?
            bad = (err > 0.05 * ref.abs().max()).nonzero()
            print("first bad idx:", bad[:8].tolist(), "n_bad", bad.shape[0], "of", M * N)
            print("got", out[:2, :8].float().tolist(), "ref", ref[:2, :8].tolist())
        tol = 2e-3 if dt == torch.float32 else 1e-2
        assert rel < tol
        # fused bias gradient: colsum += column sums of the stored (row-scaled) output
        rs = torch.rand((M + 6) // 7, device="cuda") + 0.5
        cs = torch.full((N,), 0.25, device="cuda")
        ops.gemm(A, B, out, M=M, N=N, K=K, bias=bias, rowscale=rs, rs_div=7, colsum=cs)
        ref2 = ref * rs[torch.arange(M, device="cuda") // 7].unsqueeze(1)
        assert _report("gemm colsum", cs, 0.25 + ref2.double().sum(0).float())[1] < 2e-3
        assert _report("gemm rowscale out", out, ref2)[1] < tol


@pytest.mark.parametrize("Mc,M,N", [(64, 128, 256), (512, 256, 256), (1000, 768, 768), (4100, 2304, 768), (999, 768, 3072)])
def test_gemm_tn_atomic(ops, Mc, M, N):
    """dW[M,N] += dY[Mc,M]^T X[Mc,N]  (MN-major operands, split-K, fp32 red.add)."""
    dY = _rand(Mc, M, seed=4, dtype=torch.bfloat16)
    X = _rand(Mc, N, seed=5, dtype=torch.bfloat16)
    base = _rand(M, N, seed=6)
    out = base.clone()
    ops.gemm(dY, X, out, M=M, N=N, K=Mc, trans=1, epilogue=ops.EPI_ATOMIC)
    torch.cuda.synchronize()
    ref = base + dY.float().t() @ X.float()
    err, rel = _report(f"gemm_tn {Mc}:{M}x{N}", out, ref)
    if rel > 1e-2:
        bad = (err > 0.05 * ref.abs().max()).nonzero()
        print("first bad idx:", bad[:8].tolist(), "n_bad", bad.shape[0], "of", M * N)
    assert rel < 2e-3
    # forced split counts agree
    for ks in (1, 3):
        o2 = base.clone()
        ops.gemm(dY, X, o2, M=M, N=N, K=Mc, trans=1, epilogue=ops.EPI_ATOMIC, k_splits=ks)
        assert (o2 - ref).abs().max().item() / ref.abs().max().item() < 2e-3


def test_gemm_gelu_dgelu(ops):
    M, N, K = 777, 3072, 768
    A = _rand(M, K, seed=1, dtype=torch.bfloat16)
    B = _rand(N, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    bias = _rand(N, seed=3, scale=0.1)
    pre_ref = A.float() @ B.float().t() + bias
    for dt in (torch.float32, torch.bfloat16):
        dact = torch.empty(M, N, device="cuda", dtype=dt)
        act = torch.empty(M, N, device="cuda", dtype=dt)
        ops.gemm(A, B, dact, M=M, N=N, K=K, bias=bias, epilogue=ops.EPI_GELU, out2=act)
        tol = 2e-3 if dt == torch.float32 else 1e-2
        z = pre_ref.clone().requires_grad_(True)
        torch.nn.functional.gelu(z).sum().backward()
        assert _report("gelu.dact", dact, z.grad)[1] < tol          # out = gelu'(z), saved for the backward GEMM
        assert _report("gelu.act", act, torch.nn.functional.gelu(pre_ref))[1] < tol
        # dgelu: out = (dY @ W) * gelu'(z) with the saved derivative as aux
        dY = _rand(M, 768, seed=7, dtype=torch.bfloat16)
        Wt = _rand(N, 768, seed=8, scale=0.05, dtype=torch.bfloat16)     # [N_out=3072, K=768]
        out = torch.empty(M, N, device="cuda", dtype=dt)
        ops.gemm(dY, Wt, out, M=M, N=N, K=768, epilogue=ops.EPI_DGELU, aux=dact)
        p = pre_ref.clone().requires_grad_(True)
        torch.nn.functional.gelu(p).backward(dY.float() @ Wt.float().t())
        assert _report("dgelu", out, p.grad)[1] < tol


def test_gemm_resid_maps(ops):
    Bc, T, HW, D = 2, 4, 49, 256
    L, S = T * HW, 1 + T * HW
    K = 192
    W = _rand(D, K, seed=2, scale=0.05, dtype=torch.bfloat16)
    bias = _rand(D, seed=3, scale=0.1)
    # SKIPCLS + rowscale per sequence of T tokens
    A = _rand(Bc * L, K, seed=1, dtype=torch.bfloat16)
    resid = _rand(Bc, S, D, seed=4)
    scale = (torch.rand(Bc * HW, device="cuda") > 0.3).float() / 0.7
    out = torch.zeros(Bc, S, D, device="cuda")
    ops.gemm(A, W, out, M=Bc * L, N=D, K=K, epilogue=ops.EPI_RESID, bias=bias, rowscale=scale, rs_div=T,
             map=ops.MAP_SKIPCLS, resid=resid, T=T, HW=HW, ldo=D)
    y = (A.float() @ W.float().t() + bias).view(Bc * HW, T, D) * scale.view(-1, 1, 1)
    ref = torch.zeros_like(out)
    ref[:, 1:] = resid[:, 1:] + y.view(Bc, L, D)
    assert _report("resid.skipcls", out, ref)[1] < 2e-3
    # SPATIAL scatter + cls side buffer
    A = _rand(Bc * T * (HW + 1), K, seed=5, dtype=torch.bfloat16)
    scale = (torch.rand(Bc * T, device="cuda") > 0.3).float() / 0.7
    out = torch.zeros(Bc, S, D, device="cuda")
    side = torch.zeros(Bc * T, D, device="cuda")
    ops.gemm(A, W, out, M=Bc * T * (HW + 1), N=D, K=K, epilogue=ops.EPI_RESID, bias=bias, rowscale=scale,
             rs_div=HW + 1, map=ops.MAP_SPATIAL, resid=resid, out2=side, T=T, HW=HW, ldo=D)
    y = (A.float() @ W.float().t() + bias).view(Bc * T, HW + 1, D) * scale.view(-1, 1, 1)
    ref = torch.zeros_like(out)
    ref[:, 1:] = resid[:, 1:] + y[:, 1:].reshape(Bc, T, HW, D).permute(0, 2, 1, 3).reshape(Bc, L, D)
    assert _report("resid.spatial", out, ref)[1] < 2e-3
    assert _report("resid.spatial.cls", side, y[:, 0])[1] < 2e-3
    # PATCH: + pos + time
    A = _rand(Bc * T * HW, K, seed=6, dtype=torch.bfloat16)
    pos = _rand(1 + HW, D, seed=7)
    tim = _rand(T, D, seed=8)
    out = torch.zeros(Bc, S, D, device="cuda")
    ops.gemm(A, W, out, M=Bc * T * HW, N=D, K=K, epilogue=ops.EPI_RESID, bias=bias, map=ops.MAP_PATCH, add_pos=pos,
             add_time=tim, T=T, HW=HW, ldo=D)
    y = (A.float() @ W.float().t() + bias).view(Bc, T, HW, D) + pos[1:].view(1, 1, HW, D) + tim.view(1, T, 1, D)
    ref = torch.zeros_like(out)
    ref[:, 1:] = y.permute(0, 2, 1, 3).reshape(Bc, L, D)
    assert _report("resid.patch", out, ref)[1] < 2e-3


def test_split3_gemm_precision(ops):
    """bf16x3: [hi|hi|lo] x [hi|lo|hi] over a 3x longer contraction recovers ~fp32 products."""
    M, N, K = 512, 768, 768
    A, B = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.05)
    A3 = torch.empty(M, 3 * K, device="cuda", dtype=torch.bfloat16)
    B3 = torch.empty(N, 3 * K, device="cuda", dtype=torch.bfloat16)
    ops.split3(A, A3, M, K, 0, 1)
    ops.split3(B, B3, N, K, 1, 1)
    out = torch.empty(M, N, device="cuda")
    ops.gemm(A3, B3, out, M=M, N=N, K=3 * K)
    ref = (A.double() @ B.double().t()).float()
    err, rel = _report("bf16x3", out, ref)
    one = torch.empty(M, N, device="cuda")
    ops.gemm(A.bfloat16(), B.bfloat16(), one, M=M, N=N, K=K)
    _report("bf16x1", one, ref)
    assert rel < 2e-5
    # along rows (for dW = dY^T X)
    A3r = torch.empty(3 * M, K, device="cuda", dtype=torch.bfloat16)
    ops.split3(A, A3r, M, K, 0, 0)
    hi = A.bfloat16()
    assert torch.equal(A3r[:M], hi) and torch.equal(A3r[M:2 * M], hi)
    assert torch.equal(A3r[2 * M:], (A - hi.float()).bfloat16())


# ------------------------------------------------------------------------------------------------ elementwise
def test_patchify(ops):
    Bc, T = 2, 4
    x = _rand(Bc, 3, T, 224, 224, seed=1)
    for dt in (torch.float32, torch.bfloat16):
        out = torch.empty(Bc * T * 196, 768, device="cuda", dtype=dt)
        ops.patchify(x, out)
        ref = x.permute(0, 2, 1, 3, 4).reshape(Bc * T, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(-1, 768)
        assert torch.equal(out, ref.to(dt))


@pytest.mark.parametrize("map_name", ["ident", "skipcls", "spatial", "cls"])
def test_layernorm_fwd_bwd(ops, map_name):
    Bc, T, HW, D = 2, 4, 49, 768
    L, S = T * HW, 1 + T * HW
    x = _rand(Bc, S, D, seed=1, scale=2.0) + 0.5
    x_cls = _rand(Bc, S, D, seed=9, scale=2.0)
    w, b = 1 + _rand(D, seed=2, scale=0.1), _rand(D, seed=3, scale=0.1)
    xr = x.clone().requires_grad_(True)
    xc = x_cls.clone().requires_grad_(True)
    if map_name == "ident":
        M, mp, src = Bc * S, ops.MAP_IDENT, xr.view(-1, D)
    elif map_name == "skipcls":
        M, mp, src = Bc * L, ops.MAP_SKIPCLS, xr[:, 1:].reshape(-1, D)
    elif map_name == "cls":
        M, mp, src = Bc, ops.MAP_CLS, xr[:, 0]
    else:
        M, mp = Bc * T * (HW + 1), ops.MAP_SPATIAL
        tok = xr[:, 1:].reshape(Bc, HW, T, D).permute(0, 2, 1, 3)                   # b t n d
        cls = xc[:, 0].view(Bc, 1, 1, D).expand(Bc, T, 1, D)
        src = torch.cat((cls, tok), 2).reshape(-1, D)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(src, (D,), wr, br, 1e-6)
    for dt in (torch.float32, torch.bfloat16):
        y = torch.empty(M, D, device="cuda", dtype=dt)
        stats = torch.empty(M, 2, device="cuda")
        ops.layernorm_fwd(x, w, b, y, stats, M, D, 1e-6, mp, x_cls=x_cls, T=T, HW=HW)
        assert _report(f"ln_fwd.{map_name}.{dt}", y, ref.detach())[0].max() < (2e-5 if dt == torch.float32 else 3e-2)
    dy = _rand(M, D, seed=4)
    gx, gc, gw, gb = torch.autograd.grad(ref, (xr, xc, wr, br), dy, allow_unused=True)
    dx = torch.zeros(Bc, S, D, device="cuda")
    dw, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ops.layernorm_bwd(dy, x, w, stats, dx, dw, db, M, D, mp, x_cls=x_cls, T=T, HW=HW)
    ref_dx = gx + (gc if gc is not None else 0)
    assert _report(f"ln_bwd.dx.{map_name}", dx, ref_dx)[1] < 1e-5
    assert _report(f"ln_bwd.dw.{map_name}", dw, gw)[1] < 1e-5
    assert _report(f"ln_bwd.db.{map_name}", db, gb)[1] < 1e-5


def test_gather_cast_clsmerge_colsum_castweight_embedbwd(ops):
    Bc, T, HW, D = 2, 4, 49, 256
    L, S = T * HW, 1 + T * HW
    dx = _rand(Bc, S, D, seed=1)
    scale = torch.rand(Bc * T, device="cuda") + 0.5
    out = torch.empty(Bc * T * (HW + 1), D, device="cuda")
    cs = torch.full((D,), 0.5, device="cuda")
    ops.gather_cast(dx, out, Bc * T * (HW + 1), D, ops.MAP_SPATIAL, rowscale=scale, rs_div=HW + 1, T=T, HW=HW, colsum=cs)
    tok = dx[:, 1:].reshape(Bc, HW, T, D).permute(0, 2, 1, 3)
    cls = (dx[:, 0] / T).view(Bc, 1, 1, D).expand(Bc, T, 1, D)
    ref = torch.cat((cls, tok), 2).reshape(Bc * T, HW + 1, D) * scale.view(-1, 1, 1)
    assert _report("gather.spatial", out, ref.reshape(-1, D))[0].max() < 1e-6
    assert _report("gather.colsum", cs, 0.5 + ref.reshape(-1, D).double().sum(0).float())[1] < 1e-5   # fused bias gradient
    outb = torch.empty(Bc * L, D, device="cuda", dtype=torch.bfloat16)
    ops.gather_cast(dx, outb, Bc * L, D, ops.MAP_SKIPCLS, T=T, HW=HW)
    assert torch.equal(outb, dx[:, 1:].reshape(-1, D).bfloat16())
    outp = torch.empty(Bc * T * HW, D, device="cuda")
    ops.gather_cast(dx, outp, Bc * T * HW, D, ops.MAP_PATCH, T=T, HW=HW)
    assert torch.equal(outp, tok.reshape(-1, D))
    # cls merge
    x0, side, x2 = _rand(Bc, S, D, seed=2), _rand(Bc * T, D, seed=3), torch.zeros(Bc, S, D, device="cuda")
    ops.cls_merge(x0, side, x2, Bc, T, S, D)
    assert _report("cls_merge", x2[:, 0], x0[:, 0] + side.view(Bc, T, D).mean(1))[0].max() < 1e-6
    # colsum
    for dt in (torch.float32, torch.bfloat16):
        a = _rand(1000, 768, seed=4, dtype=dt)
        o = torch.ones(768, device="cuda")
        ops.colsum(a, o, 1000, 768)
        assert _report("colsum", o, 1 + a.float().sum(0))[1] < 1e-5
    # cast weight
    w = _rand(300, 200, seed=5)
    for dt in (torch.float32, torch.bfloat16):
        wo, wt = torch.empty(300, 200, device="cuda", dtype=dt), torch.empty(200, 300, device="cuda", dtype=dt)
        ops.cast_weight(w, wo, wt)
        assert torch.equal(wo, w.to(dt)) and torch.equal(wt, w.t().contiguous().to(dt))
    # embed bwd
    dcls, dpos, dtime = torch.zeros(D, device="cuda"), torch.zeros(1 + HW, D, device="cuda"), torch.zeros(T, D, device="cuda")
    ops.embed_bwd(dx, dcls, dpos, dtime, Bc, D, T, HW)
    t4 = dx[:, 1:].reshape(Bc, HW, T, D)
    assert _report("dcls", dcls, dx[:, 0].sum(0))[1] < 1e-5
    assert _report("dpos", dpos, torch.cat((dx[:, 0].sum(0, keepdim=True), t4.sum((0, 2))), 0))[1] < 1e-5
    assert _report("dtime", dtime, t4.sum((0, 1)))[1] < 1e-5


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("n_seq,seq,H", [(37, 8, 12), (5, 4, 12), (3, 7, 5), (1, 1, 1), (3528, 8, 12), (3, 32, 12), (4, 197, 12), (1, 460, 2),
                                         (50, 32, 12), (7, 16, 12), (5, 9, 3), (3, 17, 5), (2, 24, 12), (3, 29, 2), (1, 12, 1), (393, 32, 12)])
def test_attn_simt(ops, n_seq, seq, H):
    C = H * 64
    scale = 64 ** -0.5
    for dt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)):
        qkv = _rand(n_seq * seq, 3 * C, seed=1, dtype=dt)
        qkv_r = qkv.float().requires_grad_(True)
        qr, kr, vr = (t.reshape(n_seq, seq, H, 64).transpose(1, 2) for t in qkv_r.split(C, dim=1))
        att = (qr @ kr.transpose(-1, -2)) * scale
        ref = (att.softmax(-1) @ vr).transpose(1, 2).reshape(n_seq * seq, C)
        out = torch.empty(n_seq * seq, C, device="cuda", dtype=dt)
        lse = torch.empty(n_seq, H, seq, device="cuda")
        ops.attn_fwd(qkv, out, lse, n_seq, seq, H, scale)
        assert _report(f"attn_fwd {seq} {dt}", out, ref.detach())[1] < tol
        assert _report("lse", lse, torch.logsumexp(att.detach(), -1))[0].max() < 1e-3
        do = _rand(n_seq * seq, C, seed=2, dtype=dt)      # seq > 208 takes the streaming (flash-style) backward
        (g,) = torch.autograd.grad(ref, qkv_r, do.float())
        dqkv = torch.empty_like(qkv)
        ops.attn_bwd(qkv, out, do, lse, dqkv, n_seq, seq, H, scale)
        assert _report(f"attn_bwd {seq} {dt}", dqkv, g)[1] < (1e-4 if dt == torch.float32 else 3e-2)


# ------------------------------------------------------------------------------------------------ head / loss
def test_head_ops(ops):
    M, K, N, C = 26, 768, 512, 1234
    x, w, b = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.05), _rand(N, seed=3)
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    lab = torch.nn.functional.normalize(_rand(C, N, seed=4), dim=1)
    y_ref = xr @ wr.t() + br
    e_ref = y_ref / y_ref.norm(dim=1, keepdim=True)
    lg_ref = e_ref @ lab.t() / 0.02
    y, e, nrm, lg = (torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda"), torch.empty(M, device="cuda"),
                     torch.empty(M, C, device="cuda"))
    ops.linear_small_fwd(x, w, b, y)
    ops.l2norm_fwd(y, e, nrm)
    ops.sim_logits_fwd(e, lab, lg, 50.0)
    assert _report("head.logits", lg, lg_ref.detach())[1] < 1e-5
    dlg = _rand(M, C, seed=5)
    gx, gw, gb = torch.autograd.grad(lg_ref, (xr, wr, br), dlg)
    de, dy, dx = torch.zeros(M, N, device="cuda"), torch.empty(M, N, device="cuda"), torch.empty(M, K, device="cuda")
    dw, db = torch.zeros(N, K, device="cuda"), torch.zeros(N, device="cuda")
    ops.sim_logits_bwd(dlg, lab, de, 50.0)
    ops.l2norm_bwd(e, nrm, de, dy)
    ops.linear_small_bwd(x, w, dy, dx, dw, db)
    assert _report("head.dx", dx, gx)[1] < 1e-4
    assert _report("head.dw", dw, gw)[1] < 1e-4
    assert _report("head.db", db, gb)[1] < 1e-4


@pytest.mark.parametrize("topk", [5, 0])
def test_kl_topk_loss(ops, topk):
    import timesformer_oracle as O
    M, K = 26, 9871
    pred = _rand(M, K, seed=1, scale=3.0)
    teach = _rand(M, K, seed=2, scale=3.0)
    pr = pred.clone().requires_grad_(True)
    loss_ref, _, _ = O.pretrain_loss(pr, teach, [torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")], topk)
    loss_ref.backward()
    row, dp, tout = torch.empty(M, device="cuda"), torch.empty(M, K, device="cuda"), torch.empty(M, K, device="cuda")
    ops.kl_topk_loss(pred, teach, row, dp, tout, topk)
    assert abs(row.sum().item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    assert _report("kl.teacher", tout, O.topk_teacher(teach, topk))[0].max() < 1e-6
    assert _report("kl.dpred", dp, pr.grad)[0].max() < 1e-6
    sm = torch.empty(M, K, device="cuda")
    ops.softmax_rows(pred, sm)
    assert _report("softmax", sm, pred.softmax(1))[0].max() < 1e-6


@pytest.mark.parametrize("n_seq,seq,H", [(3, 197, 12), (2, 128, 4), (2, 50, 2), (1, 256, 2), (144, 197, 12),
                                          (13, 197, 12), (30, 101, 12), (20, 145, 12), (37, 256, 5), (40, 17, 12)])
def test_attn_tensor_core(ops, n_seq, seq, H):
    """tcgen05 spatial attention (bf16) against the fp32 torch restatement and the CUDA-core kernel."""
    C = H * 64
    scale = 64 ** -0.5
    qkv = _rand(n_seq * seq, 3 * C, seed=1, dtype=torch.bfloat16)
    qkv_r = qkv.float().requires_grad_(True)
    qr, kr, vr = (t.reshape(n_seq, seq, H, 64).transpose(1, 2) for t in qkv_r.split(C, dim=1))
    att = (qr @ kr.transpose(-1, -2)) * scale
    ref = (att.softmax(-1) @ vr).transpose(1, 2).reshape(n_seq * seq, C)
    out = torch.full((n_seq * seq, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((n_seq, H, seq), float("nan"), device="cuda")
    ops.attn_tc_fwd(qkv, out, lse, n_seq, seq, H, scale)
    torch.cuda.synchronize()
    err, rel = _report(f"attn_tc_fwd {n_seq}x{seq}x{H}", out, ref.detach())
    if rel > 2e-2:
        bad = (err > 0.05 * ref.abs().max()).nonzero()
        print("first bad idx:", bad[:10].tolist(), "n_bad", bad.shape[0], "of", out.numel())
    assert rel < 2e-2
    assert _report("attn_tc lse", lse, torch.logsumexp(att.detach(), -1))[0].max() < 2e-3
    do = _rand(n_seq * seq, C, seed=2, dtype=torch.bfloat16)
    (g,) = torch.autograd.grad(ref, qkv_r, do.float())
    dqkv = torch.full_like(qkv, float("nan"))
    ops.attn_tc_bwd(qkv, out, do, lse, dqkv, n_seq, seq, H, scale)
    torch.cuda.synchronize()
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        err, rel = _report(f"attn_tc_bwd.{name} {n_seq}x{seq}x{H}", dqkv[:, sl], g[:, sl])
        if rel > 3e-2:
            bad = (err > 0.05 * g[:, sl].abs().max()).nonzero()
            print("first bad idx:", bad[:10].tolist(), "n_bad", bad.shape[0])
        assert rel < 3e-2


# ------------------------------------------------------------------------------------------------ order transformer
# Each pvrl_ot_* kernel against its torch restatement (tests/shadow_ops.py, run on the same device), fp32, rtol 1e-4;
# then one ResidualAttentionBlock level end to end against torch.nn.MultiheadAttention / LayerNorm autograd.
def _close(name, got, ref, tol=2e-4):
    err, rel = _report(name, got, ref)
    assert rel < tol, name


@pytest.mark.parametrize("M,N,K", [(18, 1536, 512), (18, 512, 2048), (45, 2048, 512), (3, 512, 128)])
def test_ot_linear(ops, M, N, K):
    import shadow_ops as S
    x, W, b, res = _rand(M, K, seed=1), _rand(N, K, seed=2, scale=0.05), _rand(N, seed=3), _rand(M, N, seed=4)
    lw, lb = 1 + 0.1 * _rand(K, seed=5), 0.1 * _rand(K, seed=6)
    for mode in (0, 1, 2):
        if mode == 1 and K > 512:
            continue
        y, yr = torch.empty(M, N, device="cuda"), torch.empty(M, N, device="cuda")
        xh, rs = torch.empty(M, K, device="cuda"), torch.empty(M, device="cuda")
        xhr, rsr = torch.empty_like(xh), torch.empty_like(rs)
        act, actr = torch.empty_like(y), torch.empty_like(yr)
        ops.ot_linear_fwd(x, W, b, y, mode, lw, lb, 1e-5, xh, rs, resid=res, act_out=act)
        S.ot_linear_fwd(x, W, b, yr, mode, lw, lb, 1e-5, xhr, rsr, resid=res, act_out=actr)
        _close(f"ot_linear_fwd mode{mode}", y, yr)
        _close(f"ot_linear_fwd mode{mode} QuickGELU(y)", act, actr)
        if mode == 1:
            _close("ot xhat", xh, xhr), _close("ot rstd", rs, rsr)
    dY = _rand(M, N, seed=7)
    pre = _rand(M, K, seed=8)
    for p in (None, pre):
        dA, dAr = torch.empty(M, K, device="cuda"), torch.empty(M, K, device="cuda")
        ops.ot_linear_dx(dY, W, dA, pre=p)
        S.ot_linear_dx(dY, W, dAr, pre=p)
        _close("ot_linear_dx", dA, dAr)
    for mode in (0, 1, 2):
        dW, db = _rand(N, K, seed=9), _rand(N, seed=10)
        dWr, dbr = dW.clone(), db.clone()
        ops.ot_linear_dw(dY, x, dW, db, mode, lw, lb)
        S.ot_linear_dw(dY, x, dWr, dbr, mode, lw, lb)
        _close(f"ot_linear_dw mode{mode}", dW, dWr), _close("ot db", db, dbr)


def test_ot_ln_attn_embed(ops):
    import shadow_ops as S
    B, Sq, H, C = 3, 9, 8, 512
    M = B * Sq
    dA, xh, rs, w = _rand(M, C, seed=1), _rand(M, C, seed=2), _rand(M, seed=3).abs() + 0.5, 1 + 0.1 * _rand(C, seed=4)
    outs = []
    for fn in (ops.ot_ln_bwd, S.ot_ln_bwd):
        dh, dw, db = _rand(M, C, seed=5), _rand(C, seed=6), _rand(C, seed=7)
        fn(dA, xh, rs, w, dh, dw, db)
        outs.append((dh, dw, db))
    for a, b, n in zip(outs[0], outs[1], ("dh", "dw", "db")):
        _close("ot_ln_bwd " + n, a, b)
    qkv, dO = _rand(M, 3 * C, seed=8), _rand(M, C, seed=9)
    pad = torch.tensor([9, 4, 6], device="cuda")
    res = []
    for mod in (ops, S):
        probs, o, dqkv = torch.empty(B, H, Sq, Sq, device="cuda"), torch.empty(M, C, device="cuda"), torch.empty(M, 3 * C, device="cuda")
        mod.ot_attn_fwd(qkv, pad, probs, o, B, Sq, H)
        mod.ot_attn_bwd(qkv, probs, dO, dqkv, B, Sq, H)
        res.append((probs, o, dqkv))
    for a, b, n in zip(res[0], res[1], ("probs", "o", "dqkv")):
        _close("ot_attn " + n, a, b)
    # attention against torch.nn.functional.multi_head_attention-style autograd
    q, k, v = (qkv[:, i * C:(i + 1) * C].reshape(B, Sq, H, 64).permute(0, 2, 1, 3).clone().requires_grad_(True) for i in range(3))
    msk = (torch.arange(Sq, device="cuda").view(1, 1, 1, Sq) >= pad.view(B, 1, 1, 1))
    oo = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=~msk)
    oo.backward(dO.reshape(B, Sq, H, 64).permute(0, 2, 1, 3))
    _close("ot_attn o vs sdpa", res[0][1], oo.permute(0, 2, 1, 3).reshape(M, C))
    _close("ot_attn dq vs sdpa", res[0][2][:, :C], q.grad.permute(0, 2, 1, 3).reshape(M, C))
    _close("ot_attn dv vs sdpa", res[0][2][:, 2 * C:], v.grad.permute(0, 2, 1, 3).reshape(M, C))
    # level input and its backward
    video, src, noise = _rand(M, C, seed=11), _rand(B, C, seed=12), _rand(B, C, seed=13)
    tw, pw, padw, tv = _rand(2, C, seed=14), _rand(Sq, C, seed=15), _rand(1, C, seed=16), _rand(C, seed=17)
    mk = torch.tensor([8, 1, 3], device="cuda")
    res = []
    for mod in (ops, S):
        h = torch.empty(M, C, device="cuda")
        mod.ot_embed_fwd(video, src, noise, 0.9, 0.3, mk, pad, tw, pw, padw, tv, h, B, Sq)
        dv_, dt_, dp_, dpad_, dtv_ = _rand(M, C, seed=18), _rand(2, C, seed=19), _rand(Sq, C, seed=20), _rand(1, C, seed=21), torch.empty(C, device="cuda")
        mod.ot_embed_bwd(dA, mk, pad, dv_, dt_, dp_, dpad_, dtv_, B, Sq)
        res.append((h, dv_, dt_, dp_, dpad_, dtv_))
    for a, b, n in zip(res[0], res[1], ("h", "dvideo", "dtype", "dpos", "dpad", "dtvec")):
        _close("ot_embed " + n, a, b)


def test_order_levels_vs_torch_modules(ops):
    """The whole pre-training branch of the order transformer on the kernels vs the same module evaluated with torch's
    own MultiheadAttention / LayerNorm / autograd (the op-by-op expression the reference uses, tfm_model.py:165-204)."""
    from procedurevrl_b200.lib.config import get_cfg
    from procedurevrl_b200.lib.models.order_tfm import DiffusionTransformer
    torch.manual_seed(3)
    m = DiffusionTransformer(num_seg=8, tfm_layers=4, hidden_size=512, cfg=get_cfg()).cuda().train()
    B, S, C, L = 2, m.max_len, 512, 4
    x = torch.nn.functional.normalize(_rand(B * S, C, seed=1), dim=1).requires_grad_(True)
    mask, pad, noise = torch.tensor([2, 8], device="cuda"), torch.tensor([5, 9], device="cuda"), _rand(L, B, C, seed=2)
    m.fixed_draws = (mask, pad, noise)
    wts = _rand(L * B, C, seed=5)
    den, _, mse, inter = m(x, is_pretrain=True)
    ((inter * wts).sum() + torch.nn.functional.mse_loss(mse[0], mse[1])).backward()
    got = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    gx = x.grad.clone()
    # torch-module restatement of the same levels
    m.zero_grad(set_to_none=True)
    x2 = x.detach().clone().requires_grad_(True)
    feats = x2.reshape(B, S, C).transpose(0, 1)
    pos = torch.arange(S, device="cuda").unsqueeze(1)
    is_mask, padm = pos == mask.unsqueeze(0), pos >= pad.unsqueeze(0)
    x0 = (feats * is_mask.unsqueeze(-1)).sum(0)
    feats = torch.where(padm.unsqueeze(-1), m.pad_embedding.weight[0], feats)
    outs, d = [], None
    for lvl in range(L):
        t = L - 1 - lvl
        src = (x0 if lvl == 0 else d).detach()
        noisy = m.sqrt_alphas_cumprod[t] * src + m.sqrt_one_minus_alphas_cumprod[t] * noise[lvl]
        d = m._level(torch.where(is_mask.unsqueeze(-1), noisy.unsqueeze(0), feats), is_mask, t, padm.t())
        outs.append(d)
    inter_ref = torch.cat(outs)
    x0t = x0.unsqueeze(0).expand(L, -1, -1).reshape(-1, C)
    ((inter_ref * wts).sum() + torch.nn.functional.mse_loss(x0t, inter_ref)).backward()
    _close("order inter", inter, inter_ref, 5e-4)
    _close("order dx", gx, x2.grad, 2e-3)
    for n, p in m.named_parameters():
        if p.grad is not None:
            _close("order grad " + n, got[n], p.grad, 2e-3)


def test_patchify_u8(ops):
    """uint8 frames + fused normalisation == patchify of the normalised fp32 frames."""
    u8 = torch.randint(0, 256, (2, 3, 4, 64, 96), device="cuda", dtype=torch.uint8)
    mean, std = (0.45, 0.40, 0.50), (0.225, 0.2, 0.25)
    xf = (u8.float() / 255.0 - torch.tensor(mean, device="cuda").view(1, 3, 1, 1, 1)) / torch.tensor(std, device="cuda").view(1, 3, 1, 1, 1)
    rows, KP = 2 * 4 * 4 * 6, 768
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 1e-2)):
        a, b = torch.empty(rows, KP, device="cuda", dtype=dt), torch.empty(rows, KP, device="cuda", dtype=dt)
        ops.patchify_u8(u8, a, 16, mean, std)
        ops.patchify(xf.contiguous(), b, 16)
        assert _report(f"patchify_u8 {dt}", a, b)[1] < tol


def test_cast_weight_multi(ops):
    """All Linear weights of a model in one launch: bf16 copy and transposed bf16 copy of every matrix."""
    shapes = [(768, 768), (2304, 768), (768, 3072), (130, 66), (64, 64)]
    ws = [_rand(r, c, seed=i) for i, (r, c) in enumerate(shapes)]
    outs = [(torch.zeros(r, c, device="cuda", dtype=torch.bfloat16), torch.zeros(c, r, device="cuda", dtype=torch.bfloat16))
            for r, c in shapes]
    table = ops.cast_weight_table([(w, o, oT) for w, (o, oT) in zip(ws, outs)])
    ops.cast_weight_multi(table)
    for w, (o, oT) in zip(ws, outs):
        assert torch.equal(o, w.bfloat16())
        assert torch.equal(oT, w.t().contiguous().bfloat16())


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("pair", ["mlp->spatial", "spatial->temporal", "temporal->mlp", "temporal->patch"])
def test_layernorm_bwd_emit(ops, pair, dt):
    """pvrl_layernorm_bwd_emit == pvrl_layernorm_bwd followed by pvrl_gather_cast: same dx / dw / db, the emitted operand
    bit-equal to the separate gather (same arithmetic on the same fp32 row), its column sums equal up to summation order."""
    Bc, T, HW, D = 3, 5, 7, 768
    L, S = T * HW, 1 + T * HW
    lmap, emap = {"mlp->spatial": (ops.MAP_IDENT, ops.MAP_SPATIAL), "spatial->temporal": (ops.MAP_SPATIAL, ops.MAP_SKIPCLS),
                  "temporal->mlp": (ops.MAP_SKIPCLS, ops.MAP_IDENT), "temporal->patch": (ops.MAP_SKIPCLS, ops.MAP_PATCH)}[pair]
    rows = {ops.MAP_IDENT: Bc * S, ops.MAP_SPATIAL: Bc * T * (HW + 1), ops.MAP_SKIPCLS: Bc * L, ops.MAP_PATCH: Bc * L}
    M, M2 = rows[lmap], rows[emap]
    rs_div = {ops.MAP_SPATIAL: HW + 1, ops.MAP_IDENT: S}.get(emap, 0)
    x, x_cls, w = _rand(Bc, S, D, seed=1), _rand(Bc, S, D, seed=2), _rand(D, seed=3)
    dy = _rand(M, D, seed=4, dtype=dt)
    stats = torch.empty(M, 2, device="cuda")
    y = torch.empty(M, D, device="cuda", dtype=dt)
    ops.layernorm_fwd(x, w, _rand(D, seed=5), y, stats, M, D, 1e-6, lmap, x_cls=x_cls, T=T, HW=HW)
    for use_rs in ([False, True] if rs_div else [False]):
        rowscale = (torch.rand(M2 // rs_div, device="cuda") > 0.3).float() / 0.7 if use_rs else None
        res = []
        for fused in (False, True):
            dx = _rand(Bc, S, D, seed=6)
            dw, db, cs = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda"), torch.full((D,), 0.5, device="cuda")
            out = torch.full((M2, D), float("nan"), device="cuda", dtype=dt)
            emit = (out, emap, rowscale, rs_div, cs) if fused else None
            ops.layernorm_bwd(dy, x, w, stats, dx, dw, db, M, D, lmap, x_cls=x_cls, T=T, HW=HW, emit=emit)
            if not fused:
                ops.gather_cast(dx, out, M2, D, emap, rowscale=rowscale, rs_div=rs_div, T=T, HW=HW, colsum=cs)
            torch.cuda.synchronize()
            res.append((dx, dw, db, out, cs))
        (dx0, dw0, db0, out0, cs0), (dx1, dw1, db1, out1, cs1) = res
        assert not torch.isnan(out1.float()).any()
        if lmap == ops.MAP_SPATIAL:      # the cls rows are accumulated with atomics: order-dependent in the last bits
            torch.testing.assert_close(dx1, dx0, rtol=1e-5, atol=1e-5)
        else:
            assert torch.equal(dx1, dx0)
        assert torch.equal(out1, out0)
        assert _report(f"emit colsum {pair}", cs1, cs0)[1] < 1e-5
        assert _report("emit dw", dw1, dw0)[1] < 1e-5 and _report("emit db", db1, db0)[1] < 1e-5
    with pytest.raises(RuntimeError, match="cannot emit"):
        ops.layernorm_bwd(dy, x, w, stats, dx, dw, db, M, D, lmap, x_cls=x_cls, T=T, HW=HW, emit=(out, ops.MAP_CLS, None, 0, None))
