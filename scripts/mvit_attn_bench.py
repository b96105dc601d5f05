"""Per-shape micro-benchmark of the MViTv2 pooled-attention kernels at the seven (heads, queries, keys) shapes of MViTv2-S
16 x 224 with 9 clips: forward and backward, mma.sync vs CUDA-core variants (PVRL_MVIT_ATTN_MMA / PVRL_MVIT_ATTN_MMA_BWD), CUDA
events, 5 launches after 2 warm-ups.  Prints one JSON line per shape (us per launch, TFLOP/s of 4 Nq Nk 96 per (clip, head)
forward, 2.5x that backward) -- the loop to iterate on these kernels without paying for a whole training step.

    python scripts/mvit_attn_bench.py [--clips 9]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from procedurevrl_b200 import ops  # noqa: E402

# (blocks of MViTv2-S that have this shape, heads, query grid, key grid)
SHAPES = [("0", 1, (8, 56, 56), (8, 7, 7)), ("1", 2, (8, 28, 28), (8, 14, 14)), ("2", 2, (8, 28, 28), (8, 7, 7)),
          ("3", 4, (8, 14, 14), (8, 14, 14)), ("4-13", 4, (8, 14, 14), (8, 7, 7)), ("14", 8, (8, 7, 7), (8, 14, 14)),
          ("15", 8, (8, 7, 7), (8, 7, 7))]


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=9)
    a = ap.parse_args()
    B, C = a.clips, 96
    for blocks, heads, qg, kg in SHAPES:
        Nq, Nk = 1 + qg[0] * qg[1] * qg[2], 1 + kg[0] * kg[1] * kg[2]
        g = torch.Generator().manual_seed(Nq + Nk)
        q, k, v = (torch.randn(B, heads, n, C, generator=g).cuda().bfloat16() for n in (Nq, Nk, Nk))
        bq = (0.5 * torch.randn(B, heads, Nq - 1, sum(kg), generator=g)).cuda()
        dout = torch.randn(B, Nq, heads * C, generator=g).cuda().bfloat16()
        out, lse = torch.empty_like(dout), torch.empty(B, heads, Nq, device="cuda")
        dq, dbq, delta = torch.empty_like(q), torch.empty_like(bq), torch.empty_like(lse)
        dk, dv = torch.zeros(B, heads, Nk, C, device="cuda"), torch.zeros(B, heads, Nk, C, device="cuda")
        scale = C ** -0.5
        row = {"blocks": blocks, "heads": heads, "Nq": Nq, "Nk": Nk}
        flop = 4.0 * B * heads * Nq * Nk * C
        for name, env in (("simt", "0"), ("mma", "1")):      # the proven variant first: a fault in an unproven one keeps its numbers
            os.environ["PVRL_MVIT_ATTN_MMA"] = env
            us = timed(lambda: ops.pooled_attn_fwd(q, k, v, bq, out, lse, kg, scale, True))
            row[f"fwd_{name}_us"], row[f"fwd_{name}_tflops"] = round(us, 1), round(flop / us / 1e6, 1)
            os.environ["PVRL_MVIT_ATTN_MMA_BWD"] = env
            us = timed(lambda: ops.pooled_attn_bwd(q, k, v, bq, dout, lse, dq, dk, dv, dbq, delta, kg, scale, True))
            row[f"bwd_{name}_us"], row[f"bwd_{name}_tflops"] = round(us, 1), round(2.5 * flop / us / 1e6, 1)
        for e in ("PVRL_MVIT_ATTN_MMA", "PVRL_MVIT_ATTN_MMA_BWD"):
            os.environ.pop(e, None)
        print(json.dumps(row))


if __name__ == "__main__":
    main()
