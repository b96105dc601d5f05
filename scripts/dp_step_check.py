#!/usr/bin/env python
"""Data-parallel pre-training step on N GPUs (depth 2, one video x 9 clips per rank, CUDA graph): a few replayed steps with
the gradient exchange named by PVRL_GRAD_EXCHANGE / PVRL_SPLIT_UPDATE, then a checksum of the parameters and a NORMAL exit
(graph and trainer released, destroy_process_group, no os._exit).  tests/test_zy_multi_gpu.py launches it under
torch.distributed.run for every exchange variant and compares the lines.
   torchrun --nproc-per-node 2 scripts/dp_step_check.py [steps]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    D = bench.Dist(argparse.Namespace(gpus=world))
    h = bench.build_step(D, T=8, depth=2, precision="bf16", Bv=1, use_graph=True)
    tr = h["trainer"]
    losses = []
    for _ in range(steps):
        losses.append(float(tr()))
    torch.cuda.synchronize()
    flat = tr.opt.flat_param.double()
    line = {"world": world, "exchange": tr.exchange_kind, "split_update": bool(getattr(tr, "_split_update", False)),
            "blocks_per_bucket": tr.blocks_per_bucket, "mode": h["mode"], "losses": losses,
            "param_sum": flat.sum().item(), "param_abs_sum": flat.abs().sum().item(),
            "param_sq_sum": (flat * flat).sum().item()}
    # every rank must hold the same parameters after the exchange
    chk = torch.tensor([line["param_sum"], line["param_abs_sum"]], device=D.dev, dtype=torch.float64)
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        D.dist.all_reduce(lo, op=D.dist.ReduceOp.MIN)
        D.dist.all_reduce(hi, op=D.dist.ReduceOp.MAX)
    line["ranks_agree"] = bool(torch.equal(lo, hi))
    if D.rank == 0:
        print(json.dumps(line), flush=True)
    del tr, h
    bench.release()
    D.barrier()
    if world > 1:
        D.dist.destroy_process_group()


if __name__ == "__main__":
    main()
