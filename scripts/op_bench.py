#!/usr/bin/env python
"""Per-kernel micro-benchmark at the benchmark shapes (18 clips of 8x224, D 768): every C-ABI op of one block's
forward/backward timed alone with CUDA events (mean of --iters back-to-back launches after warm-up) and reported as
TFLOP/s (tensor-bound ops) and algorithmic GB/s (HBM-bound ops).  Development tool; bench.py is the contract."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from procedurevrl_b200 import ops  # noqa: E402


WARM = 3
COLD = None      # when set: a buffer larger than L2 that is rewritten between timed launches (every launch sees a cold L2,
                 # as inside the real step where ~14 GB stream through between two uses of a tensor)


def timeit(fn, iters, warm=None):
    warm = WARM if warm is None else warm
    if COLD is not None:
        for _ in range(warm):
            fn()
        tot = 0.0
        for _ in range(iters):
            COLD.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / iters * 1e3
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--clips", type=int, default=18)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--only", default="")
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--json", default="")
    ap.add_argument("--cold", action="store_true", help="flush L2 (rewrite a 512 MB buffer) before every timed launch")
    a = ap.parse_args()
    global WARM, COLD
    WARM = a.warm
    if a.cold:
        COLD = torch.zeros(128 * 1024 * 1024, device="cuda")
    dev = torch.device("cuda")
    Bc, T, HW, D, H, Hd = a.clips, a.frames, 196, 768, 12, 3072
    L, S = HW * T, 1 + HW * T
    Mt, Ms, Mm = Bc * L, Bc * T * (HW + 1), Bc * S
    g = dict(T=T, HW=HW)
    bf = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)
    f32 = lambda *s: torch.randn(*s, device=dev) * 0.5
    rows = []

    def add(name, fn, flops=0.0, bytes_=0.0):
        if a.only and a.only not in name:
            return
        us = timeit(fn, a.iters)
        rows.append(dict(op=name, us=round(us, 1), tflops=round(flops / us / 1e6, 1) if flops else None,
                         gbs=round(bytes_ / us / 1e3, 1) if bytes_ else None))
        print(f"{name:34s} {us:9.1f} us  {rows[-1]['tflops'] or '':>8} TF/s  {rows[-1]['gbs'] or '':>8} GB/s", flush=True)

    # ---- GEMMs (forward)
    w_qkv, w_d, w_fc1, w_fc2 = bf(3 * D, D), bf(D, D), bf(Hd, D), bf(D, Hd)
    bias3, bias1, biash = f32(3 * D), f32(D), f32(Hd)
    x0, x1, x2 = f32(Bc, S, D), f32(Bc, S, D), f32(Bc, S, D)
    side = f32(Bc * T, D)
    a_t, a_s, a_m = bf(Mt, D), bf(Ms, D), bf(Mm, D)
    qkv_t, qkv_s = bf(Mt, 3 * D), bf(Ms, 3 * D)
    o_t, o_s = bf(Mt, D), bf(Ms, D)
    pre, hid = bf(Mm, Hd), bf(Mm, Hd)
    add("gemm qkv_t  STORE  M28224 N2304 K768", lambda: ops.gemm(a_t, w_qkv, qkv_t, M=Mt, N=3 * D, K=D, bias=bias3),
        2.0 * Mt * 3 * D * D, Mt * D * 2 + Mt * 3 * D * 2)
    add("gemm proj_t STORE  M28224 N768 K768", lambda: ops.gemm(a_t, w_d, o_t, M=Mt, N=D, K=D, bias=bias1),
        2.0 * Mt * D * D, Mt * D * 4)
    add("gemm fc_t   RESID  M28224 N768 K768",
        lambda: ops.gemm(a_t, w_d, x1, M=Mt, N=D, K=D, epilogue=ops.EPI_RESID, bias=bias1, map=ops.MAP_SKIPCLS,
                         resid=x0, ldo=D, **g), 2.0 * Mt * D * D, Mt * D * (2 + 4 + 4))
    add("gemm proj_s RESID(spatial scatter)",
        lambda: ops.gemm(a_s, w_d, x2, M=Ms, N=D, K=D, epilogue=ops.EPI_RESID, bias=bias1, map=ops.MAP_SPATIAL,
                         resid=x1, out2=side, ldo=D, **g), 2.0 * Ms * D * D, Ms * D * (2 + 4 + 4))
    add("gemm fc1    GELU   M28242 N3072 K768",
        lambda: ops.gemm(a_m, w_fc1, pre, M=Mm, N=Hd, K=D, epilogue=ops.EPI_GELU, bias=biash, out2=hid),
        2.0 * Mm * Hd * D, Mm * D * 2 + 2 * Mm * Hd * 2)
    add("gemm fc2    RESID  M28242 N768 K3072",
        lambda: ops.gemm(hid, w_fc2, x0, M=Mm, N=D, K=Hd, epilogue=ops.EPI_RESID, bias=bias1, map=ops.MAP_IDENT,
                         resid=x2, ldo=D), 2.0 * Mm * Hd * D, Mm * Hd * 2 + Mm * D * 8)
    # ---- GEMMs (backward)
    d_pre = bf(Mm, Hd)
    add("gemm dfc2   DGELU  M28242 N3072 K768",
        lambda: ops.gemm(a_m, w_fc1, d_pre, M=Mm, N=Hd, K=D, epilogue=ops.EPI_DGELU, aux=pre),
        2.0 * Mm * Hd * D, Mm * D * 2 + 2 * Mm * Hd * 2)
    add("gemm dfc1   STORE  M28242 N768 K3072", lambda: ops.gemm(hid, w_fc2, a_m, M=Mm, N=D, K=Hd),
        2.0 * Mm * Hd * D, Mm * Hd * 2 + Mm * D * 2)
    add("gemm dqkv   STORE  M28224 N768 K2304", lambda: ops.gemm(qkv_t, bf(D, 3 * D), a_t, M=Mt, N=D, K=3 * D),
        2.0 * Mt * 3 * D * D, Mt * 3 * D * 2 + Mt * D * 2)
    gw_fc, gw_d, gw_qkv = f32(D, Hd), f32(D, D), f32(3 * D, D)
    gw_fc1 = f32(Hd, D)
    add("gemm dW fc2  TN M768 N3072 K28242",
        lambda: ops.gemm(a_m, hid, gw_fc, M=D, N=Hd, K=Mm, trans=1, epilogue=ops.EPI_ATOMIC, ldo=Hd),
        2.0 * Mm * Hd * D, Mm * D * 2 + Mm * Hd * 2)
    add("gemm dW fc1  TN M3072 N768 K28242",
        lambda: ops.gemm(hid, a_m, gw_fc1, M=Hd, N=D, K=Mm, trans=1, epilogue=ops.EPI_ATOMIC, ldo=D),
        2.0 * Mm * Hd * D, Mm * D * 2 + Mm * Hd * 2)
    add("gemm dW proj TN M768 N768 K28224",
        lambda: ops.gemm(a_t, o_t, gw_d, M=D, N=D, K=Mt, trans=1, epilogue=ops.EPI_ATOMIC, ldo=D),
        2.0 * Mt * D * D, Mt * D * 4)
    add("gemm dW qkv  TN M2304 N768 K28224",
        lambda: ops.gemm(qkv_t, a_t, gw_qkv, M=3 * D, N=D, K=Mt, trans=1, epilogue=ops.EPI_ATOMIC, ldo=D),
        2.0 * Mt * 3 * D * D, Mt * D * 8)

    # ---- LayerNorm
    lw, lb = f32(D), f32(D)
    st_t, st_s, st_m = f32(Mt, 2).abs() + 0.5, f32(Ms, 2).abs() + 0.5, f32(Mm, 2).abs() + 0.5
    add("layernorm_fwd temporal (SKIPCLS)", lambda: ops.layernorm_fwd(x0, lw, lb, a_t, st_t, Mt, D, 1e-6, ops.MAP_SKIPCLS, **g),
        0, Mt * D * 6)
    add("layernorm_fwd spatial (gather)", lambda: ops.layernorm_fwd(x1, lw, lb, a_s, st_s, Ms, D, 1e-6, ops.MAP_SPATIAL, x_cls=x0, **g),
        0, Ms * D * 6)
    add("layernorm_fwd mlp (IDENT)", lambda: ops.layernorm_fwd(x2, lw, lb, a_m, st_m, Mm, D, 1e-6, ops.MAP_IDENT), 0, Mm * D * 6)
    dx = f32(Bc, S, D)
    dlw, dlb = f32(D), f32(D)
    add("layernorm_bwd temporal", lambda: ops.layernorm_bwd(a_t, x0, lw, st_t, dx, dlw, dlb, Mt, D, ops.MAP_SKIPCLS, **g),
        0, Mt * D * (2 + 4 + 4 + 4))
    add("layernorm_bwd spatial", lambda: ops.layernorm_bwd(a_s, x1, lw, st_s, dx, dlw, dlb, Ms, D, ops.MAP_SPATIAL, x_cls=x0, **g),
        0, Ms * D * (2 + 4 + 4 + 4))
    add("layernorm_bwd mlp", lambda: ops.layernorm_bwd(a_m, x2, lw, st_m, dx, dlw, dlb, Mm, D, ops.MAP_IDENT), 0, Mm * D * 14)

    # ---- attention
    lse_t, lse_s = f32(Bc * HW, H, T), f32(Bc * T, H, HW + 1)
    scale = 0.125
    add("attn_fwd temporal (T=8)", lambda: ops.attn_fwd(qkv_t, o_t, lse_t, Bc * HW, T, H, scale),
        4.0 * Bc * HW * H * T * T * 64, Mt * D * 8)
    ops.attn_fwd(qkv_t, o_t, lse_t, Bc * HW, T, H, scale)
    dqkv_t = bf(Mt, 3 * D)
    add("attn_bwd temporal (T=8)", lambda: ops.attn_bwd(qkv_t, o_t, a_t, lse_t, dqkv_t, Bc * HW, T, H, scale),
        10.0 * Bc * HW * H * T * T * 64, Mt * D * (6 + 2 + 2 + 6))
    add("attn_tc_fwd spatial (197)", lambda: ops.attn_tc_fwd(qkv_s, o_s, lse_s, Bc * T, HW + 1, H, scale),
        4.0 * Bc * T * H * 197 * 197 * 64, Ms * D * 8)
    ops.attn_tc_fwd(qkv_s, o_s, lse_s, Bc * T, HW + 1, H, scale)
    dqkv_s = bf(Ms, 3 * D)
    add("attn_tc_bwd spatial (197)", lambda: ops.attn_tc_bwd(qkv_s, o_s, a_s, lse_s, dqkv_s, Bc * T, HW + 1, H, scale),
        10.0 * Bc * T * H * 197 * 197 * 64, Ms * D * 16)

    # ---- gradient plumbing
    add("gather_cast IDENT", lambda: ops.gather_cast(dx, a_m, Mm, D, ops.MAP_IDENT), 0, Mm * D * 6)
    add("gather_cast SPATIAL", lambda: ops.gather_cast(dx, a_s, Ms, D, ops.MAP_SPATIAL, **g), 0, Ms * D * 6)
    gb = f32(Hd)
    add("colsum [28242, 3072]", lambda: ops.colsum(pre, gb, Mm, Hd), 0, Mm * Hd * 2)
    add("colsum [28224, 768]", lambda: ops.colsum(a_t, dlb, Mt, D), 0, Mt * D * 2)
    add("colsum [28224, 2304]", lambda: ops.colsum(qkv_t, f32(3 * D), Mt, 3 * D), 0, Mt * 3 * D * 2)
    wm, wb, wt = f32(Hd, D), bf(Hd, D), bf(D, Hd)
    add("cast_weight [3072, 768]", lambda: ops.cast_weight(wm, wb, wt), 0, Hd * D * 8)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
