"""First hardware run of the mma.sync dQ / dK/dV passes of the MViTv2 pooled attention (csrc/mvit_attn_mma.cu, opt-in:
PVRL_MVIT_ATTN_MMA_BWD=1) and of the one-channel-per-lane pooling weight-gradient kernel (csrc/mvit.cu, PVRL_POOL_DW=2).  These two kernels were written after the round's GPU time was spent: their fragment algebra is
pinned by the lane-level CPU emulation (tests/test_mvit_mma_emulation.py, the procedure that the forward kernel passed
before its first -- green -- GPU run), but they have not executed on a GPU.  So this file (a) sorts last, (b) runs the
kernels in a CHILD process -- a fault in an unproven kernel must not poison the CUDA context of the rest of the suite --
and (c) is marked xfail(strict=False): a pass is reported as XPASS, a miss does not turn the suite red.  The default path
(CUDA-core backward) is what every other test exercises."""
import json
import os
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

CHILD = r"""
import json, os, sys
sys.path[:0] = [%(root)r, %(here)r]
import torch
import shadow_ops as S
from procedurevrl_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
C = 96
rel = lambda a, r: ((a - r).abs().max() / (r.abs().max() + 1e-12)).item()
# the one-channel-per-lane weight-gradient kernel of the Q / K / V pooling (PVRL_POOL_DW=2, csrc/mvit.cu)
pool = []
for B, heads, grid, stride in [(2, 2, (4, 8, 8), (1, 2, 2)), (1, 4, (2, 7, 7), (1, 1, 1)), (9, 1, (8, 56, 56), (1, 1, 1))]:
    L = grid[0] * grid[1] * grid[2]
    g = torch.Generator().manual_seed(L)
    src = torch.randn(B, 1 + L, 3 * heads * C, generator=g).cuda().bfloat16()
    w = (0.3 * torch.randn(C, 27, generator=g)).cuda()
    og = ops.pool_out_grid(grid, (3, 3, 3), stride, (1, 1, 1))
    dout = torch.randn(B, heads, 1 + og[0] * og[1] * og[2], C, generator=g).cuda().bfloat16()
    dws, times = {}, {}
    for name, env in (("v1", "1"), ("cg", "2")):
        os.environ["PVRL_POOL_DW"] = env
        dw, din = torch.zeros(C, 27, device="cuda"), torch.empty_like(src)
        ops.pool3d_bwd(dout, src, heads * C, w, din, dw, heads, C, grid, (3, 3, 3), stride, (1, 1, 1))
        torch.cuda.synchronize()
        dws[name] = dw
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.pool3d_bwd(dout, src, heads * C, w, din, torch.zeros(C, 27, device="cuda"), heads, C, grid, (3, 3, 3), stride, (1, 1, 1))
        e1.record()
        torch.cuda.synchronize()
        times[name] = e0.elapsed_time(e1) / 3 * 1e3
    os.environ.pop("PVRL_POOL_DW")
    pool.append({"shape": [B, heads, list(grid), list(stride)], "dw_cg_vs_v1": rel(dws["cg"], dws["v1"]),
                 "us_both_kernels_v1": times["v1"], "us_both_kernels_cg": times["cg"]})
print("RESULT_POOL " + json.dumps(pool), flush=True)
out = []
for B, heads, qg, kg in [(2, 2, (2, 8, 8), (2, 4, 4)), (1, 1, (1, 5, 5), (1, 3, 3)), (1, 3, (4, 14, 14), (4, 7, 7)),
                         (1, 1, (2, 6, 6), (2, 7, 7)), (9, 4, (8, 14, 14), (8, 7, 7))]:
    Nq, Nk = 1 + qg[0] * qg[1] * qg[2], 1 + kg[0] * kg[1] * kg[2]
    g = torch.Generator().manual_seed(Nq + Nk)
    q, k, v = ((torch.randn(B, heads, n, C, generator=g)).cuda().bfloat16() for n in (Nq, Nk, Nk))
    bq = (0.5 * torch.randn(B, heads, Nq - 1, sum(kg), generator=g)).cuda()
    dout = torch.randn(B, Nq, heads * C, generator=g).cuda().bfloat16()
    scale = C ** -0.5
    res = {}
    for name, o, env in (("mma", ops, "1"), ("simt", ops, "0"), ("ref", S, "0")):
        o_, lse = torch.empty(B, Nq, heads * C, device="cuda", dtype=torch.bfloat16), torch.empty(B, heads, Nq, device="cuda")
        o.pooled_attn_fwd(q, k, v, bq, o_, lse, kg, scale, True)
        dq = torch.empty_like(q)
        dk, dv = torch.zeros(B, heads, Nk, C, device="cuda"), torch.zeros(B, heads, Nk, C, device="cuda")
        dbq, delta = torch.empty_like(bq), torch.empty_like(lse)
        os.environ["PVRL_MVIT_ATTN_MMA_BWD"] = env
        o.pooled_attn_bwd(q, k, v, bq, dout, lse, dq, dk, dv, dbq, delta, kg, scale, True)
        torch.cuda.synchronize()
        res[name] = dict(dq=dq.float(), dk=dk, dv=dv, dbq=dbq, delta=delta)
    row = {"shape": [B, heads, Nq, Nk]}
    for key in ("dq", "dk", "dv", "dbq", "delta"):
        row[key] = rel(res["mma"][key], res["ref"][key])
        row[key + "_simt"] = rel(res["simt"][key], res["ref"][key])
    if B == 9:
        for env in ("1", "0"):
            os.environ["PVRL_MVIT_ATTN_MMA_BWD"] = env
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ops.pooled_attn_bwd(q, k, v, bq, dout, lse, dq, dk, dv, dbq, delta, kg, scale, True)
            e1.record()
            torch.cuda.synchronize()
            row["us_mma" if env == "1" else "us_simt"] = e0.elapsed_time(e1) / 3 * 1e3
    out.append(row)
print("RESULT_ATTN " + json.dumps(out), flush=True)
"""


@pytest.mark.gpu
def test_mma_backward_first_hardware_run():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "here": HERE}], capture_output=True, text=True, timeout=300)
    got = {l.split(" ", 1)[0]: json.loads(l.split(" ", 1)[1]) for l in r.stdout.splitlines() if l.startswith("RESULT_")}
    print(r.stdout[-3000:], r.stderr[-2000:])
    res = {"pool": got.get("RESULT_POOL", []), "attn": got.get("RESULT_ATTN", [])}
    assert r.returncode == 0 and res["pool"] and res["attn"], "the child process failed (partial results are printed above)"
    for row in res["pool"]:
        print("[pool dw variant]", row)
        assert row["dw_cg_vs_v1"] < 1e-3, row
    for row in res["attn"]:
        print("[mma bwd]", row)
        for key in ("dq", "dk", "dv", "dbq"):
            assert row[key] < 3e-2, (row["shape"], key, row[key])
        assert row["delta"] < 1e-3, (row["shape"], row["delta"])
