"""One pre-training step as a single replayable unit: forward + KL/MSE loss + backward + gradient all-reduce +
AdamW, optionally captured in a CUDA graph (B200-first replacement for the per-op Python dispatch of
tools/train_net.py:146-191; no tracing compiler involved -- the graph is a recording of our own kernel launches).

Gradients live in ONE flat fp32 buffer (`.grad`s are views of it): the engine's dW/db kernels accumulate straight
into it, and data-parallel replicas synchronise with a single NCCL all-reduce over that buffer (SURVEY.md 8e:
the only collective on the path)."""
import torch

from . import functional as PF


class PretrainStep:
    def __init__(self, model, cfg, lr=5e-5, weight_decay=1e-4, process_group=None, use_graph=True):
        self.model = model
        self.inner = model.model                         # VisionTransformer mirror
        self.cfg = cfg
        self.topk = cfg.TRAIN.TOPK
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        self.flat_grad = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        eng = self.inner.engine()
        by_name = dict(self.inner.named_parameters())
        eng.grad_sink = {n: by_name[n].grad for n in eng.grad_names}
        assert all(g is not None for g in eng.grad_sink.values()), "every encoder parameter must be trainable here"
        self.opt = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay, fused=True, capturable=use_graph)
        self.use_graph = use_graph
        self.graph = None
        self.static = None

    def _eager(self, frames, meta):
        self.inner.engine().invalidate_weights()
        self.flat_grad.zero_()
        pred, teacher, mse = self.model([frames, meta])
        loss, _, _ = PF.pretrain_loss(pred, teacher, mse, topk=self.topk)
        loss.backward()
        if self.world > 1:
            torch.distributed.all_reduce(self.flat_grad, op=torch.distributed.ReduceOp.AVG, group=self.pg)
        self.opt.step()
        return loss.detach()

    def capture(self, frames, meta, warmup=3):
        """Warm up on a side stream, then record one step into a CUDA graph with `frames` / `meta` as static inputs."""
        self.static = (frames, meta)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager(frames, meta)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager(frames, meta)
        return self

    def __call__(self, frames=None, meta=None):
        """Run one step.  With a captured graph, `frames` (if given) is copied into the static input buffer."""
        if self.graph is None:
            return self._eager(frames, meta)
        if frames is not None and frames.data_ptr() != self.static[0].data_ptr():
            self.static[0].copy_(frames, non_blocking=True)
        self.graph.replay()
        return self.static_loss
