#!/usr/bin/env python
"""Time the step's one collective alone: NCCL all-reduce (AVG) of the flat fp32 gradient buffer (134.6 M elements = 538 MB),
10 repetitions after 3 warm-ups, CUDA events, max over ranks.  Launch with torch.distributed.run; NCCL_* variables select
the algorithm / CTA budget.  Development aid for the 1 -> 8 GPU scaling curve (bench.py is the contract)."""
import os
import sys

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 134_600_000
x = torch.randn(n, device="cuda")
for _ in range(3):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
torch.cuda.synchronize()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dist.all_reduce(x, op=dist.ReduceOp.AVG)
e1.record()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
w = dist.get_world_size()
if dist.get_rank() == 0:
    ms = t.item()
    print(f"allreduce {n * 4 / 1e6:.0f} MB x{w}: {ms:.3f} ms  algbw {n * 4 / ms / 1e6:.0f} GB/s  busbw {n * 4 / ms / 1e6 * 2 * (w - 1) / w:.0f} GB/s  "
          f"env {{{', '.join(k + '=' + v for k, v in os.environ.items() if k.startswith('NCCL_'))}}}", flush=True)
dist.destroy_process_group()
