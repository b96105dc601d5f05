"""Gradient exchange of the data-parallel replicas over NVLink peer memory, driven by the copy engines.

The path's only collective is the mean of the flat fp32 gradient buffer (reference: the NCCL all-reduce
behind DistributedDataParallel, lib/models/build.py:49-53).  NCCL's all-reduce is a kernel: run beside the backward it
takes SMs away from the persistent tcgen05 GEMMs (measured on 2 and 8 B200s: no gain from overlapping it), run after
the backward its ~1.6-2.9 ms are fully exposed.  Here the exchange needs NO SMs for the transfers:

  * the flat gradient buffer lives in symmetric memory (every rank maps every peer's buffer through NVLink);
  * a range [a, b) of it (a bucket of encoder blocks, as soon as the backward has finished them) is cut into `world`
    chunks; rank r PULLS chunk r from every peer into a local staging buffer with plain device-to-device copies --
    DMA engines, NVLink reads -- sums them into its own chunk with one small kernel (`pvrl_reduce_chunks`), then pulls
    the other ranks' reduced chunks back (all-gather, again DMA);
  * three stream-ordered device barriers per range (signal pads in symmetric memory) order the phases across ranks:
    gradients complete -> pull;  chunks reduced -> gather;  all pulls of my chunk done -> the optimizer may clear it.

Everything is stream-ordered (copies, one kernel, barrier kernels), so a whole training step including its exchanges is
captured in one CUDA graph, and the copies of bucket i run under the backward GEMMs of the blocks below it.
"""
import torch
import torch.distributed as dist

from . import ops


class PeerGradExchange:
    def __init__(self, numel, group=None, device=None, n_copy_streams=4):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.numel = int(numel)
        # the gradient buffer itself: symmetric, zero-initialised (the dW kernels accumulate into it)
        self.buffer = symm.empty(self.numel, dtype=torch.float32, device=self.device)
        self.buffer.zero_()
        self.hdl = symm.rendezvous(self.buffer, self.group)
        self.peers = [self.buffer if p == self.rank else self.hdl.get_buffer(p, (self.numel,), torch.float32, 0)
                      for p in range(self.world)]
        self.comm = torch.cuda.Stream(device=self.device)
        self.copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(max(1, n_copy_streams))]
        self._stage = None
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)

    @staticmethod
    def chunk_bounds(a, b, world):
        """`world` chunks of [a, b): equal sizes rounded up to 4 elements (16-byte pieces), the last ones may be short / empty."""
        n = b - a
        c = -(-n // world)
        c = (c + 3) // 4 * 4
        return [(min(b, a + r * c), min(b, a + (r + 1) * c)) for r in range(world)]

    def _staging(self, n):
        if self._stage is None or self._stage.shape[1] < n:
            self._stage = torch.empty(self.world - 1, (n + 3) // 4 * 4, device=self.device, dtype=torch.float32)
        return self._stage

    def reserve(self, max_range):
        """Allocate the staging buffer for ranges of up to `max_range` elements (call before CUDA-graph capture)."""
        lo, hi = self.chunk_bounds(0, max_range, self.world)[0]
        self._staging(hi - lo)

    def all_reduce_mean(self, a, b, after=None):
        """buffer[a:b] <- mean over ranks, on the exchange stream.  `after`: the stream whose work produced the range
        (default: the current stream).  Returns nothing; call `join()` before consuming / clearing the buffer."""
        main = after if after is not None else torch.cuda.current_stream(self.device)
        bounds = self.chunk_bounds(a, b, self.world)
        lo, hi = bounds[self.rank]
        n_own = hi - lo
        others = [p for p in range(self.world) if p != self.rank]
        stage = self._staging(max(e - s for s, e in bounds))
        self.comm.wait_stream(main)
        with torch.cuda.stream(self.comm):
            self.hdl.barrier(channel=0)                       # every rank's gradients of this range are complete
            if n_own > 0:
                used = []
                for k, p in enumerate(others):                # reduce-scatter: pull my chunk from every peer (DMA over NVLink)
                    s = self.copy_streams[k % len(self.copy_streams)]
                    s.wait_stream(self.comm)
                    with torch.cuda.stream(s):
                        stage[k, :n_own].copy_(self.peers[p][lo:hi], non_blocking=True)
                    used.append(s)
                for s in set(used):
                    self.comm.wait_stream(s)
                ops.reduce_chunks(self.buffer[lo:hi], stage, len(others), stage.stride(0), n_own, 1.0 / self.world)
            self.hdl.barrier(channel=1)                       # every chunk is reduced
            used = []
            for k, p in enumerate(others):                    # all-gather: pull the other ranks' reduced chunks
                s0, e0 = bounds[p]
                if e0 > s0:
                    s = self.copy_streams[k % len(self.copy_streams)]
                    s.wait_stream(self.comm)
                    with torch.cuda.stream(s):
                        self.buffer[s0:e0].copy_(self.peers[p][s0:e0], non_blocking=True)
                    used.append(s)
            for s in set(used):
                self.comm.wait_stream(s)
            self.hdl.barrier(channel=2)                       # nobody still reads my chunk: it may be consumed / cleared

    def mark(self):
        """Event after everything enqueued on the exchange stream so far (capturable): lets the caller consume the ranges
        already exchanged while a later `all_reduce_mean` is still running."""
        ev = torch.cuda.Event()
        ev.record(self.comm)
        return ev

    def join(self, stream=None):
        (stream if stream is not None else torch.cuda.current_stream(self.device)).wait_stream(self.comm)
