#!/usr/bin/env python
"""Copy-engine gradient exchange (procedurevrl_b200/grad_exchange.py) against NCCL on N GPUs: same result, timing alone
and inside a CUDA graph.   torchrun --nproc-per-node N scripts/ce_allreduce_test.py"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from procedurevrl_b200.grad_exchange import PeerGradExchange  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 134_600_000
ex = PeerGradExchange(n, dist.group.WORLD, dev)
ex.reserve(n)
g = torch.Generator(device="cuda").manual_seed(100 + rank)


def fill():
    ex.buffer.copy_(torch.randn(n, device=dev, generator=g))


# ---- correctness: whole buffer and an odd sub-range
for (a, b) in ((0, n), (12345, n - 77), (5, 5 + 3)):
    fill()
    ref = ex.buffer.clone()
    dist.all_reduce(ref[a:b], op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    dist.barrier()
    ex.all_reduce_mean(a, b)
    ex.join()
    torch.cuda.synchronize()
    err = (ex.buffer - ref).abs().max().item()
    if rank == 0:
        print(f"range [{a}, {b}): max |ce - nccl| = {err:.3e}", flush=True)
    assert err < 1e-5, err

# ---- timing, eager
def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def ce():
    ex.all_reduce_mean(0, n)
    ex.join()


t_nccl = timeit(lambda: dist.all_reduce(ex.buffer, op=dist.ReduceOp.AVG))
t_ce = timeit(ce)
if rank == 0:
    print(f"{n * 4 / 1e6:.0f} MB x{world}: NCCL {t_nccl:.3f} ms, copy-engine exchange {t_ce:.3f} ms (eager)", flush=True)

# ---- inside a CUDA graph
fill()
ref = ex.buffer.clone()
dist.all_reduce(ref, op=dist.ReduceOp.AVG)
torch.cuda.synchronize()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
graph = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    ce()
    torch.cuda.synchronize()
    dist.barrier()
    fill()
    ref = ex.buffer.clone()
    dist.all_reduce(ref, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, stream=s):
        ce()
torch.cuda.synchronize()
dist.barrier()
graph.replay()
torch.cuda.synchronize()
err = (ex.buffer - ref).abs().max().item()
t_graph = timeit(graph.replay)
if rank == 0:
    print(f"graph replay: max |ce - nccl| = {err:.3e}; {t_graph:.3f} ms per replay", flush=True)
assert err < 1e-5
del graph
dist.barrier()
dist.destroy_process_group()
