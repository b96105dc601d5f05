#!/usr/bin/env python
"""Cost of the order transformer's pre-training branch (4 levels x 4 blocks, fwd + bwd) replayed from a CUDA graph:
pvrl_ot_* kernels vs the op-by-op torch expression.  Development tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from procedurevrl_b200 import ops  # noqa: E402
from procedurevrl_b200.lib.config import get_cfg  # noqa: E402
from procedurevrl_b200.lib.models.order_tfm import DiffusionTransformer  # noqa: E402


def graph_time(fn, iters=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    torch.manual_seed(0)
    m = DiffusionTransformer(num_seg=8, tfm_layers=4, hidden_size=512, cfg=get_cfg()).cuda().train()
    B, S, C, L = 2, m.max_len, 512, 4
    x = torch.nn.functional.normalize(torch.randn(B * S, C, device="cuda"), dim=1).requires_grad_(True)
    m.fixed_draws = (torch.tensor([2, 8], device="cuda"), torch.tensor([5, 9], device="cuda"),
                     torch.randn(L, B, C, device="cuda"))
    w = torch.randn(L * B, C, device="cuda")

    def kernels():
        for p in m.parameters():
            p.grad = None
        x.grad = None
        n0 = ops.launch_count()
        _, _, mse, inter = m(x, is_pretrain=True)
        ((inter * w).sum() + torch.nn.functional.mse_loss(mse[0], mse[1])).backward()
        return ops.launch_count() - n0

    def torch_path():
        for p in m.parameters():
            p.grad = None
        x.grad = None
        mask, pad, noise = m.fixed_draws
        feats = x.reshape(B, S, C).transpose(0, 1)
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
        inter = torch.cat(outs)
        x0t = x0.unsqueeze(0).expand(L, -1, -1).reshape(-1, C)
        ((inter * w).sum() + torch.nn.functional.mse_loss(x0t, inter)).backward()

    print("pvrl launches per fwd+bwd:", kernels())
    print(f"pvrl_ot kernels : {graph_time(kernels):9.1f} us per fwd+bwd (graph replay)")
    print(f"torch op-by-op  : {graph_time(torch_path):9.1f} us per fwd+bwd (graph replay)")
    # per-kernel timings, back to back (warm caches / clocks)
    M = B * S
    f = lambda *s: torch.randn(*s, device="cuda")
    xx, W1, b1, y1 = f(M, 512), f(1536, 512), f(1536), f(M, 1536)
    W2, y2, u = f(512, 2048), f(M, 512), f(M, 2048)
    lw, lb, xh, rs = f(512), f(512), f(M, 512), f(M)
    def t(fn, n=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    print("fwd LN+in_proj  [18,512]x[1536,512]:", round(t(lambda: ops.ot_linear_fwd(xx, W1, b1, y1, 1, lw, lb, 1e-5, xh, rs)), 1), "us (eager back-to-back incl. launch)")
    print("fwd qgelu+c_proj [18,2048]x[512,2048]:", round(t(lambda: ops.ot_linear_fwd(u, W2, None, y2, 2, resid=xx)), 1), "us")
    dA, dU = f(M, 512), f(M, 2048)
    print("dx in_proj N1536 K512:", round(t(lambda: ops.ot_linear_dx(y1, W1, dA)), 1), "us")
    print("dx c_proj  N512 K2048 (+qgelu'):", round(t(lambda: ops.ot_linear_dx(y2, W2, dU, pre=u)), 1), "us")
    gW, gb = f(1536, 512), f(1536)
    print("dw in_proj:", round(t(lambda: ops.ot_linear_dw(y1, xh, gW, gb, 1, lw, lb)), 1), "us")


if __name__ == "__main__":
    main()
