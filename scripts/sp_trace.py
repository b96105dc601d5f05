#!/usr/bin/env python
"""PVRL_SP_TRACE=1 python scripts/sp_trace.py [fwd|bwd]: per-phase cycle counts of CTA 0 of the persistent spatial
attention kernels at the benchmark shape (development aid; see pvrl_debug_sp_trace in include/pvrl.h)."""
import os
import sys

os.environ["PVRL_SP_TRACE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from procedurevrl_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
n_seq, seq, H = 144, 197, 12
C = H * 64
qkv = (torch.randn(n_seq * seq, 3 * C, device="cuda") * 0.5).to(torch.bfloat16)
out = torch.empty(n_seq * seq, C, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(n_seq, H, seq, device="cuda")
do = (torch.randn(n_seq * seq, C, device="cuda") * 0.5).to(torch.bfloat16)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    ops.attn_tc_fwd(qkv, out, lse, n_seq, seq, H, 0.125)
    if which == "bwd":
        ops.attn_tc_bwd(qkv, out, do, lse, dqkv, n_seq, seq, H, 0.125)
ops.debug_sp_trace()
if which == "bwd":
    ops.attn_tc_bwd(qkv, out, do, lse, dqkv, n_seq, seq, H, 0.125)
else:
    ops.attn_tc_fwd(qkv, out, lse, n_seq, seq, H, 0.125)
tr = ops.debug_sp_trace()
t0 = tr[tr > 0].min()
print("rows: warp; per problem the stamps relative to the CTA's first stamp (cycles)")
for w in range(20):
    if not (tr[w] > 0).any():
        continue
    print(f"warp {w}")
    for it in range(12):
        row = tr[w, it]
        if not (row > 0).any():
            continue
        print(f"  it {it:2d}: " + " ".join(f"{(v - t0) if v > 0 else -1:7d}" for v in row))
