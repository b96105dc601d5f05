"""One pre-training step as a single replayable unit: forward + KL/MSE loss + backward + gradient all-reduce +
AdamW, optionally captured in a CUDA graph (B200-first replacement for the per-op Python dispatch of
tools/train_net.py:146-191; no tracing compiler involved -- the graph is a recording of our own kernel launches).

Gradients live in ONE flat fp32 buffer (`.grad`s are views of it): the engine's dW/db kernels accumulate straight
into it, and data-parallel replicas synchronise with a single NCCL all-reduce over that buffer (SURVEY.md 8e:
the only collective on the path)."""
import os

import torch

from . import functional as PF
from .lib.models.optimizer import FlatOptimizer


def bucket_ranges(named_sizes, depth, blocks_per_bucket):
    """Partition of the flat gradient buffer (parameters in `named_sizes` order) for the overlapped exchange:
    ({first block index of a bucket -> (start, end)}, front_end, tail_start).  A bucket is fired when the backward has
    finished its first (lowest) block; [0, front_end) -- embeddings and the first blocks -- and [tail_start, total) --
    final norm, head, order transformer -- are exchanged at the end.  Buckets, front and tail tile the buffer exactly."""
    offs, off = {}, 0
    for n, sz in named_sizes:
        offs[n] = (off, off + sz)
        off += sz
    blk = []
    for i in range(depth):
        r = [v for k, v in offs.items() if f".blocks.{i}." in k]
        blk.append((min(a for a, _ in r), max(b for _, b in r)))
    assert all(blk[i][1] == blk[i + 1][0] for i in range(depth - 1)), "encoder blocks must be contiguous in the flat buffer"
    ranges, front_end, hi = {}, blk[0][0], depth
    while hi > 0:
        lo = max(0, hi - blocks_per_bucket)
        if lo > 0:
            ranges[lo] = (blk[lo][0], blk[hi - 1][1])
        else:
            front_end = blk[hi - 1][1]
        hi = lo
    return ranges, front_end, blk[depth - 1][1]


class PretrainStep:
    def __init__(self, model, cfg, lr=5e-5, weight_decay=1e-4, process_group=None, use_graph=True, optimizer=None,
                 accum_steps=1):
        self.model = model
        self.inner = model.model                         # VisionTransformer mirror
        self.cfg = cfg
        self.topk = cfg.TRAIN.TOPK
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        # Data-parallel gradient exchange (the only collective).  Default ("ce"): the flat gradient buffer lives
        # in symmetric memory and buckets of encoder blocks are averaged over NVLink by the copy engines while the backward
        # continues (grad_exchange.PeerGradExchange: DMA pulls + one small reduction kernel + device barriers, no SMs taken
        # from the persistent GEMMs).  PVRL_GRAD_EXCHANGE=nccl: NCCL all-reduce(s) of the same buffer.
        self.exchange = None
        self.exchange_kind = os.environ.get("PVRL_GRAD_EXCHANGE", "ce") if self.world > 1 else "none"
        grad_buffer = None
        self.exchange_note = None
        if self.exchange_kind == "ce" and optimizer is None:
            from .grad_exchange import PeerGradExchange
            try:
                self.exchange = PeerGradExchange(sum(p.numel() for p in self.params), process_group)
                grad_buffer = self.exchange.buffer
            except Exception as e:      # no peer access / no symmetric-memory support on this box: NCCL does the same mean
                if os.environ.get("PVRL_GRAD_EXCHANGE") == "ce":
                    raise               # asked for explicitly: fail loudly
                self.exchange, self.exchange_kind = None, "nccl"
                self.exchange_note = f"copy-engine exchange unavailable ({type(e).__name__}: {e}); NCCL all-reduce instead"[:300]
        elif self.exchange_kind == "ce":
            self.exchange_kind = "nccl"           # a caller-built optimizer owns its gradient buffer
        # Parameters, gradients and AdamW state in flat fp32 buffers (`p.data` / `p.grad` become views): the engine's
        # dW / db kernels accumulate straight into flat_grad, the update + the clearing of the gradients for the next
        # backward is one `pvrl_adam_flat` launch (procedurevrl_adamw.yaml SOLVER: one group, uniform decay).
        self.opt = optimizer if optimizer is not None else \
            FlatOptimizer([{"params": self.params, "lr_mult": 1.0}], "adamw", lr=lr, weight_decay=weight_decay,
                          grad_buffer=grad_buffer)
        self.flat_grad = self.opt.flat_grad
        # Gradient accumulation (train_net.py:99-100,182-192: GLOBAL_BATCH_SIZE // (NUM_SHARDS * TRAIN.BATCH_SIZE) micro-steps,
        # `p.grad /= num_iters`, then step + zero_grad): the dW kernels accumulate across micro-steps by construction,
        # the division rides in the optimizer pass (grad_scale), and -- unlike the reference's DDP, which all-reduces on
        # every micro-step -- the replicas exchange the accumulated gradient once, before the update (same mean).
        self.accum_steps = int(accum_steps)
        assert self.accum_steps >= 1
        self.opt.grad_scale = 1.0 / self.accum_steps
        self._micro = 0
        eng = self.inner.engine()
        by_name = dict(self.inner.named_parameters())
        eng.grad_sink = {n: by_name[n].grad for n in eng.grad_names}
        assert all(g is not None for g in eng.grad_sink.values()), "every encoder parameter must be trainable here"
        if getattr(self.inner, "order_tfm", None) is not None:
            self.inner.order_tfm.grad_into_params = True      # same for the order transformer's 48 block parameters
        # Buckets: PVRL_AR_BLOCKS_PER_BUCKET = n exchanges n encoder blocks at a time as soon as the backward has finished
        # them (part of the captured CUDA graph).  Copy-engine exchange: default n = 1 (the transfers cost the GEMMs nothing; only block 0 + the embeddings stay exposed).
        # NCCL: default 0 = ONE all-reduce after the backward -- overlapped NCCL buckets were measured on 2 and on 8 B200s
        # (profiles/README.md): its kernels take SMs from the persistent GEMMs for as long as the exposed exchange costs.
        default_bpb = "1" if self.exchange is not None else "0"
        self.blocks_per_bucket = int(os.environ.get("PVRL_AR_BLOCKS_PER_BUCKET", default_bpb))
        if self.accum_steps > 1:
            self.blocks_per_bucket = 0            # under accumulation the replicas exchange once, before the update
        self._ar_stream = torch.cuda.Stream() if self.world > 1 else None
        self._ar_ranges = None
        if self.world > 1 and self.blocks_per_bucket > 0:
            name_of = {id(p): n for n, p in model.named_parameters()}
            names = [(name_of[id(p)], p.numel()) for p in self.opt._params]       # the flat buffers' layout order
            self._ar_ranges, self._ar_front_end, self._ar_tail_start = bucket_ranges(names, eng.depth,
                                                                                     self.blocks_per_bucket)
            eng.on_block_bwd_done = self._bucket_ready
            if self.exchange is not None:
                self.exchange.reserve(max([b - a for a, b in self._ar_ranges.values()] +
                                          [self._ar_front_end, self.flat_grad.numel() - self._ar_tail_start]))
        elif self.exchange is not None:
            self.exchange.reserve(self.flat_grad.numel())
        self._split_update = os.environ.get("PVRL_SPLIT_UPDATE", "1") != "0"
        self._depth = eng.depth
        self.use_graph = use_graph
        self.graph = None
        self.static = None

    def _bucket_ready(self, i):
        """Encoder block i's parameter gradients are complete (called by the engine's backward, top block first)."""
        if self.exchange is not None:
            # final norm / head / order transformer: complete once the encoder backward has begun (their kernels were
            # enqueued before it) -- exchanged first, under the whole encoder backward
            if i == self._depth - 1:
                self.exchange.all_reduce_mean(self._ar_tail_start, self.flat_grad.numel())
            rng = self._ar_ranges.get(i)
            if rng is not None:
                self.exchange.all_reduce_mean(rng[0], rng[1])
            return
        rng = self._ar_ranges.get(i)
        if rng is None:
            return
        self._ar_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._ar_stream):
            torch.distributed.all_reduce(self.flat_grad[rng[0]:rng[1]], op=torch.distributed.ReduceOp.AVG, group=self.pg)

    def _eager(self, frames, meta, update=True):
        """One micro-step: forward, loss, backward (gradients accumulate into flat_grad); with `update`, also the
        gradient exchange and the optimizer pass, which leaves flat_grad cleared for the next backward."""
        self.inner.engine().invalidate_weights()     # flat_grad is clean or holds the earlier micro-steps' sum
        pred, teacher, mse = self.model([frames, meta])
        loss, _, _ = PF.pretrain_loss(pred, teacher, mse, topk=self.topk)
        loss.backward()
        if not update:
            return loss.detach()
        if self.exchange is not None:
            if self._ar_ranges is None:
                self.exchange.all_reduce_mean(0, self.flat_grad.numel())
                self.exchange.join()
            elif self._split_update:
                # embeddings + lowest blocks are the only part of the exchange the backward cannot cover: it runs under
                # the optimizer pass over everything above them, whose buckets are complete at `done`
                done = self.exchange.mark()
                self.exchange.all_reduce_mean(0, self._ar_front_end)
                torch.cuda.current_stream().wait_event(done)
                self.opt.step(zero_grad=True, split=(self._ar_front_end, self.exchange.join))
                return loss.detach()
            else:
                self.exchange.all_reduce_mean(0, self._ar_front_end)
                self.exchange.join()
        elif self.world > 1:
            if self._ar_ranges is None:
                torch.distributed.all_reduce(self.flat_grad, op=torch.distributed.ReduceOp.AVG, group=self.pg)
            else:      # what the block buckets did not cover: [embeddings | first blocks) and (norm | head | order transformer]
                AVG = torch.distributed.ReduceOp.AVG
                torch.distributed.all_reduce(self.flat_grad[:self._ar_front_end], op=AVG, group=self.pg)
                torch.distributed.all_reduce(self.flat_grad[self._ar_tail_start:], op=AVG, group=self.pg)
                torch.cuda.current_stream().wait_stream(self._ar_stream)
        self.opt.step(zero_grad=True)
        return loss.detach()

    def capture(self, frames, meta, warmup=3):
        """Warm up on a side stream, then record one step into a CUDA graph with `frames` / `meta` as static inputs
        (with accumulation: a second graph of the update-free micro-step)."""
        assert not (self.accum_steps > 1 and self._ar_ranges is not None), \
            "bucketed exchange during the backward and gradient accumulation are not combined"
        self.static = (frames, meta)
        # The warm-up steps are real optimizer steps on the capture batch: snapshot parameters, optimizer state and step
        # counter and put them back afterwards, so that capturing is invisible to the training trajectory.
        snap = (self.opt.flat_param.clone(), [b.clone() for b in self.opt._state], self.opt._step_dev.clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager(frames, meta)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.no_grad():
            self.opt.flat_param.copy_(snap[0])
            for b, b0 in zip(self.opt._state, snap[1]):
                b.copy_(b0)
            self.opt._step_dev.copy_(snap[2])
            self.flat_grad.zero_()
        del snap
        self.inner.engine().invalidate_weights()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager(frames, meta)
        self.micro_graph = None
        if self.accum_steps > 1:
            self.micro_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.micro_graph, pool=self.graph.pool()):
                self.static_micro_loss = self._eager(frames, meta, update=False)
        return self

    def __call__(self, frames=None, meta=None):
        """Run one (micro-)step; every `accum_steps`-th call exchanges the gradients and updates the parameters.  With a
        captured graph, `frames` (if given) is copied into the static input buffer."""
        self._micro += 1
        update = self._micro % self.accum_steps == 0
        if self.graph is None:
            if update:
                self.opt.sync_hyper()
            return self._eager(frames, meta, update)
        if frames is not None and frames.data_ptr() != self.static[0].data_ptr():
            self.static[0].copy_(frames, non_blocking=True)
        if not update:
            self.micro_graph.replay()
            return self.static_micro_loss
        self.opt.sync_hyper()                       # learning-rate changes reach the graph through a device scalar
        self.graph.replay()
        # The replay cast the bf16 operand copies at its start and updated the fp32 masters at its end, through raw
        # pointers: no `_version` changed.  Mark the copies stale so that an eager forward after training (evaluation,
        # validation) re-casts them instead of running on weights one optimizer step old.
        self.inner.engine().invalidate_weights()
        return self.static_loss


class AutogradPretrainStep:
    """The same pre-training step for a model whose encoder is scheduled through autograd (the MViTv2 mirror, BASELINE
    config 5): forward + KL/MSE loss + `loss.backward()` into the flat gradient buffer of a `FlatOptimizer`, ONE NCCL
    all-reduce of that buffer, one `pvrl_adam_flat` pass that also clears the gradients.  Eager dispatch (no CUDA graph):
    the MViT schedule still contains torch glue between the C-ABI ops."""

    def __init__(self, model, cfg, lr=5e-5, weight_decay=1e-4, process_group=None, use_graph=False, optimizer=None):
        self.model, self.cfg, self.topk = model, cfg, cfg.TRAIN.TOPK
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        params = [p for p in model.parameters() if p.requires_grad]
        self.opt = optimizer if optimizer is not None else \
            FlatOptimizer([{"params": params, "lr_mult": 1.0}], "adamw", lr=lr, weight_decay=weight_decay)
        self.flat_grad = self.opt.flat_grad
        self.graph = None

    def capture(self, frames, meta, warmup=2):
        for _ in range(warmup):
            self(frames, meta)
        return self

    def _eager(self, frames, meta, update=True):
        pred, teacher, mse = self.model([frames, meta])
        loss, _, _ = PF.pretrain_loss(pred, teacher, mse, topk=self.topk)
        loss.backward()
        if update:
            if self.world > 1:
                torch.distributed.all_reduce(self.flat_grad, op=torch.distributed.ReduceOp.AVG, group=self.pg)
            self.opt.step(zero_grad=True)
        return loss.detach()

    def __call__(self, frames=None, meta=None):
        return self._eager(frames, meta)
