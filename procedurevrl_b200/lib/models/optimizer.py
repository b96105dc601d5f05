"""Optimizer boundary (reference lib/models/optimizer.py:10-140, lib/utils/lr_policy.py:8-94).

`construct_optimizer(model, cfg)` keeps the reference's parameter grouping (names containing "bn" / "text_model" /
"head" / "order", TRAIN.LINEAR freezing, TRAIN.MULT learning-rate multipliers, BN.WEIGHT_DECAY vs SOLVER.WEIGHT_DECAY)
and its SOLVER.OPTIMIZING_METHOD switch (sgd | adam | adamw), but returns a `FlatOptimizer`: parameters, gradients
and optimizer state of all groups live in flat fp32 buffers (every `p.data` / `p.grad` is a view), and `step()` is ONE
`pvrl_adam_flat` / `pvrl_sgd_flat` launch per group (csrc/optim.cu) instead of torch's multi-tensor-apply chain.
Learning rate and step count are device scalars, so the step can sit inside a replayed CUDA graph (trainer.PretrainStep)
while `set_lr` keeps changing the rate every iteration as tools/train_net.py:123-124 does.

`state_dict()` / `load_state_dict()` use torch.optim's layout (per-parameter `step` / `exp_avg` / `exp_avg_sq` or
`momentum_buffer`, `param_groups` with integer ids) so `optimizer_state` entries of reference checkpoints
(lib/utils/checkpoint.py:129,397) round-trip."""
import math

import torch

from ... import ops


# ------------------------------------------------------------------------------------------------ lr policy
def lr_func_cosine(cfg, cur_epoch):
    """lr_policy.py:30-48: half-cosine from BASE_LR to COSINE_END_LR over MAX_EPOCH."""
    s = cfg.SOLVER
    assert s.COSINE_END_LR < s.BASE_LR
    return s.COSINE_END_LR + 0.5 * (s.BASE_LR - s.COSINE_END_LR) * (1.0 + math.cos(math.pi * cur_epoch / s.MAX_EPOCH))


def get_step_index(cfg, cur_epoch):
    """lr_policy.py:65-77 (including its quirk: an epoch beyond MAX_EPOCH maps to the second-to-last entry)."""
    bounds = list(cfg.SOLVER.STEPS) + [cfg.SOLVER.MAX_EPOCH]
    ind = 0
    for ind, b in enumerate(bounds):
        if cur_epoch < b:
            break
    return ind - 1


def lr_func_steps_with_relative_lrs(cfg, cur_epoch):
    """lr_policy.py:51-62."""
    return cfg.SOLVER.LRS[get_step_index(cfg, cur_epoch)] * cfg.SOLVER.BASE_LR


_POLICIES = {"cosine": lr_func_cosine, "steps_with_relative_lrs": lr_func_steps_with_relative_lrs}


def get_lr_func(lr_policy):
    if lr_policy not in _POLICIES:
        raise NotImplementedError("Unknown LR policy: {}".format(lr_policy))
    return _POLICIES[lr_policy]


def get_lr_at_epoch(cfg, cur_epoch):
    """lr_policy.py:8-27: the policy's value, linearly warmed up from WARMUP_START_LR over WARMUP_EPOCHS."""
    f = get_lr_func(cfg.SOLVER.LR_POLICY)
    lr = f(cfg, cur_epoch)
    w = cfg.SOLVER.WARMUP_EPOCHS
    if cur_epoch < w:
        start = cfg.SOLVER.WARMUP_START_LR
        lr = start + cur_epoch * (f(cfg, w) - start) / w
    return lr


def get_epoch_lr(cur_epoch, cfg):
    """optimizer.py:118-127."""
    return get_lr_at_epoch(cfg, cur_epoch)


def set_lr(optimizer, new_lr):
    """optimizer.py:130-140: every group's rate = new_lr x its lr_mult."""
    for g in optimizer.param_groups:
        g["lr"] = new_lr * g["lr_mult"] if "lr_mult" in g else new_lr


# ------------------------------------------------------------------------------------------------ grouping
def parameter_groups(model, cfg):
    """The group lists of optimizer.py:18-84 (same freezing side effects on `requires_grad`)."""
    sol, mult = cfg.SOLVER, cfg.TRAIN.MULT
    named = list(model.named_parameters())
    if mult != 1.0 or cfg.TRAIN.LINEAR:             # fine-tuning: encoder vs head / order transformer (optimizer.py:20-40)
        enc, head = [], []
        for name, p in named:
            if "head" not in name and "order" not in name:
                enc.append(p)
                if cfg.TRAIN.LINEAR:
                    p.requires_grad = False
            else:
                head.append(p)
        if cfg.TRAIN.LINEAR:
            groups = [{"params": head, "weight_decay": sol.WEIGHT_DECAY, "lr": sol.BASE_LR, "lr_mult": 1.0}]
        else:
            groups = [{"params": enc, "weight_decay": cfg.BN.WEIGHT_DECAY, "lr_mult": mult},
                      {"params": head, "weight_decay": sol.WEIGHT_DECAY, "lr_mult": 1.0}]
        assert len(named) == len(enc) + len(head)
        return groups
    bn, text, rest = [], [], []                     # pre-training (optimizer.py:41-88)
    for name, p in named:
        if "bn" in name:
            bn.append(p)
        elif "text_model" in name or "text_module" in name:
            text.append(p)
            if mult == 0:
                p.requires_grad = False
        else:
            rest.append(p)
    assert len(named) == len(bn) + len(text) + len(rest)
    # TRAIN.MULT == 1 here (the branch above took every other value): text parameters join the main group
    return [{"params": bn, "weight_decay": cfg.BN.WEIGHT_DECAY, "lr_mult": 1.0},
            {"params": rest + text, "weight_decay": sol.WEIGHT_DECAY, "lr_mult": 1.0}]


def construct_optimizer(model, cfg):
    """optimizer.py:10-114."""
    sol = cfg.SOLVER
    groups = parameter_groups(model, cfg)
    method = sol.OPTIMIZING_METHOD
    if method == "sgd":
        return FlatOptimizer(groups, "sgd", lr=sol.BASE_LR, momentum=sol.MOMENTUM, dampening=sol.DAMPENING,
                             nesterov=sol.NESTEROV, weight_decay=sol.WEIGHT_DECAY)
    if method in ("adam", "adamw"):
        return FlatOptimizer(groups, method, lr=sol.BASE_LR, betas=(0.9, 0.999), eps=1e-08,
                             weight_decay=sol.WEIGHT_DECAY)
    raise NotImplementedError("Does not support {} optimizer".format(method))


# ------------------------------------------------------------------------------------------------ flat optimizer
def split_runs(runs, off):
    """[(group, start, end)] element runs of the flat buffers -> (the parts at or above `off`, the parts below it): together
    they cover every run exactly once (FlatOptimizer.step(split=...): the upper part is updated while the gradient exchange of
    the lower part is still in flight)."""
    above = [(g, max(s, off), e) for g, s, e in runs if e > off]
    below = [(g, s, min(e, off)) for g, s, e in runs if s < off]
    return above, below


class FlatOptimizer:
    """torch.optim-shaped (param_groups / step / zero_grad / state_dict) optimizer over flat buffers.

    Layout: trainable parameters of group 0, then group 1, ... in the order given; `offsets[i] = (start, end)` of
    TRAINABLE parameter i (`_params[i]`) inside `flat_param` / `flat_grad` / the state buffers.  Frozen parameters
    (`requires_grad == False` at construction) stay in `param_groups` and in torch.optim's integer numbering
    (`_ids[i]` = the torch id of trainable parameter i) exactly as torch.optim keeps them -- the reference's groups always
    contain frozen parameters (`head.*` in every fine-tuning config, the CLIP text tower in pre-training), so an
    `optimizer_state` saved by either side loads on the other -- but they get no slot in the flat buffers and no state.
    One deliberate difference: gradients are views that are cleared, never dropped, so a trainable parameter that the
    backward does not reach is updated with a zero gradient (it still sees weight decay and its moments decay) where
    torch would skip it; setting `p.grad = None` before `step()` restores torch's behaviour for that step."""

    def __init__(self, groups, method="adamw", lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, momentum=0.0,
                 dampening=0.0, nesterov=False, grad_buffer=None):
        """grad_buffer: optional pre-allocated fp32 CUDA tensor (at least as many elements as there are trainable
        parameters) to use as the flat gradient buffer -- e.g. symmetric memory for the peer-to-peer gradient exchange
        (procedurevrl_b200/grad_exchange.py)."""
        assert method in ("sgd", "adam", "adamw")
        if isinstance(groups, (list, tuple)) and groups and not isinstance(groups[0], dict):
            groups = [{"params": list(groups)}]
        self.method = method
        self.defaults = dict(lr=lr, weight_decay=weight_decay)
        if method == "sgd":
            if nesterov and (momentum <= 0 or dampening != 0):
                raise ValueError("Nesterov momentum requires a momentum and zero dampening")
            self.defaults.update(momentum=momentum, dampening=dampening, nesterov=nesterov)
        else:
            self.defaults.update(betas=tuple(betas), eps=eps)
        self.param_groups, self.offsets, self._ranges = [], [], []
        self._params, self._ids = [], []
        total, dev, next_id = 0, None, 0
        for g in groups:
            g = dict(g)
            for k, v in self.defaults.items():
                g.setdefault(k, v)
            g["params"] = list(g["params"])          # frozen parameters included, as torch.optim keeps them
            start = total
            for p in g["params"]:
                if p.requires_grad:
                    assert p.is_cuda and p.dtype == torch.float32, \
                        "FlatOptimizer drives fp32 CUDA parameters (no CPU path)"
                    dev = p.device if dev is None else dev
                    self.offsets.append((total, total + p.numel()))
                    total += p.numel()
                    self._params.append(p)
                    self._ids.append(next_id)
                next_id += 1
            self._ranges.append((start, total))
            self.param_groups.append(g)
        if total == 0:
            raise ValueError("optimizer got an empty parameter list")
        self.flat_param = torch.empty(total, device=dev, dtype=torch.float32)
        if grad_buffer is not None:
            assert grad_buffer.is_cuda and grad_buffer.dtype == torch.float32 and grad_buffer.numel() >= total
            self.flat_grad = grad_buffer[:total]
            self.flat_grad.zero_()
        else:
            self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        n_state = 1 if method == "sgd" else 2
        self._state = [torch.zeros(total, device=dev, dtype=torch.float32) for _ in range(n_state)]
        with torch.no_grad():
            for p, (a, b) in zip(self._params, self.offsets):
                view = self.flat_param[a:b].view_as(p)
                view.copy_(p)
                p.data = view
                old = p.grad
                p.grad = self.flat_grad[a:b].view_as(p)
                if old is not None:
                    p.grad.copy_(old)
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.float32)
        self._lr_dev = torch.zeros(len(self.param_groups), device=dev, dtype=torch.float32)
        self._lr_host = [None] * len(self.param_groups)
        self.grad_scale = 1.0
        self.sync_hyper()

    # -- hyper-parameters that change between steps ---------------------------------------------------------
    def sync_hyper(self):
        """Push the groups' current `lr` to the device scalars the kernels read.  Must run OUTSIDE graph capture
        (PretrainStep calls it before every replay); `step()` calls it itself when no capture is in progress."""
        lrs = [float(g["lr"]) for g in self.param_groups]
        if lrs != self._lr_host:
            self._lr_dev.copy_(torch.tensor(lrs, dtype=torch.float32), non_blocking=False)
            self._lr_host = lrs

    # -- torch.optim surface --------------------------------------------------------------------------------
    def _grad_view(self, i):
        a, b = self.offsets[i]
        return self.flat_grad[a:b].view_as(self._params[i])

    def zero_grad(self, set_to_none=False):
        """Gradients stay views of the flat buffer (the engine's dW kernels accumulate into them), so they are cleared,
        never dropped.  A `.grad` that something else re-pointed (DistributedDataParallel's bucket views) is cleared
        where it lives; a `.grad` that was set to None gets its view back."""
        self.flat_grad.zero_()
        base = self.flat_grad.data_ptr()
        for i, (p, (a, _)) in enumerate(zip(self._params, self.offsets)):
            if p.grad is None:
                p.grad = self._grad_view(i)
            elif p.grad.data_ptr() != base + 4 * a:
                p.grad.zero_()

    def _adopt_foreign_grads(self):
        """Called by the eager `step()`.  Checks that every parameter still IS its slice of the flat buffer (a later
        `model.to(...)` / `.cuda()` would silently detach them), copies gradients that a wrapper re-pointed (e.g.
        DistributedDataParallel with gradient_as_bucket_view) into the flat buffer, and returns the parameters
        without any gradient (torch skips those)."""
        missing = []
        pbase, gbase = self.flat_param.data_ptr(), self.flat_grad.data_ptr()
        for i, (p, (a, b)) in enumerate(zip(self._params, self.offsets)):
            if p.data_ptr() != pbase + 4 * a:
                raise RuntimeError("FlatOptimizer: a parameter no longer lives in the flat buffer (was the model moved or "
                                   "re-cast after the optimizer was built?) -- construct the optimizer last")
            if p.grad is None:
                missing.append(i)
            elif p.grad.data_ptr() != gbase + 4 * a:
                self.flat_grad[a:b].view_as(p).copy_(p.grad)
        return missing

    def step(self, closure=None, zero_grad=False, split=None):
        """One update of every group.  `zero_grad=True` also clears the gradients in the same pass (the
        optimizer.step(); optimizer.zero_grad() pair of train_net.py:191-192 as one read-modify-write).
        `split=(offset, hook)`: update the elements at or above `offset` of the flat buffers first, call `hook()`, then
        update the rest -- the caller's hook waits for a gradient exchange that still covers [0, offset), which thereby
        runs under the bulk of the update (trainer.PretrainStep).  Same result as the unsplit call, element for element."""
        loss = closure() if closure is not None else None
        capturing = torch.cuda.is_current_stream_capturing()
        missing = []
        if not capturing:
            self.sync_hyper()
            missing = self._adopt_foreign_grads()
            torch.autograd.graph.increment_version(self._params)   # the kernels write through raw pointers
        ops.optim_tick(self._step_dev)
        runs = [(gi, g, s, e) for gi, (g, (a, b)) in enumerate(zip(self.param_groups, self._ranges))
                for s, e in _runs(a, b, [self.offsets[i] for i in missing])]
        if split is None:
            for gi, g, s, e in runs:
                self._launch(gi, g, s, e, zero_grad)
            return loss
        off, hook = split
        above, below = split_runs([(gi, s, e) for gi, _, s, e in runs], off)
        for gi, s, e in above:
            self._launch(gi, self.param_groups[gi], s, e, zero_grad)
        hook()
        for gi, s, e in below:
            self._launch(gi, self.param_groups[gi], s, e, zero_grad)
        return loss

    def _launch(self, gi, g, s, e, zero_grad):
        lr, sl = self._lr_dev[gi:gi + 1], slice(s, e)
        if self.method == "sgd":
            ops.sgd_flat(self.flat_param[sl], self.flat_grad[sl], self._state[0][sl], lr, self._step_dev,
                         momentum=g["momentum"], dampening=g["dampening"], nesterov=g["nesterov"],
                         weight_decay=g["weight_decay"], grad_scale=self.grad_scale, zero_grad=zero_grad)
        else:
            b1, b2 = g["betas"]
            ops.adam_flat(self.flat_param[sl], self.flat_grad[sl], self._state[0][sl], self._state[1][sl], lr,
                          self._step_dev, beta1=b1, beta2=b2, eps=g["eps"], weight_decay=g["weight_decay"],
                          decoupled=self.method == "adamw", grad_scale=self.grad_scale, zero_grad=zero_grad)

    # -- checkpoints (torch.optim layout) -------------------------------------------------------------------
    _STATE_KEYS = {"sgd": ("momentum_buffer",), "adam": ("exp_avg", "exp_avg_sq"), "adamw": ("exp_avg", "exp_avg_sq")}

    def state_dict(self):
        step = float(self._step_dev.item())
        state, keys = {}, self._STATE_KEYS[self.method]
        if step > 0:
            for pid, p, (a, b) in zip(self._ids, self._params, self.offsets):
                st = {k: buf[a:b].view_as(p).clone() for k, buf in zip(keys, self._state)}
                if self.method != "sgd":
                    st["step"] = torch.tensor(step)
                state[pid] = st
        groups, i = [], 0
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != "params"}
            d["params"] = list(range(i, i + len(g["params"])))
            i += len(g["params"])
            groups.append(d)
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if [len(g["params"]) for g in groups] != [len(g["params"]) for g in self.param_groups]:
            raise ValueError("loaded state dict contains parameter groups that don't match the optimizer's")
        for mine, theirs in zip(self.param_groups, groups):
            mine.update({k: v for k, v in theirs.items() if k != "params"})
        ids = [i for g in groups for i in g["params"]]      # their id of the parameter at torch position 0, 1, 2, ...
        keys, step = self._STATE_KEYS[self.method], 0.0
        for buf in self._state:
            buf.zero_()
        for pos, tid in enumerate(self._ids):               # trainable parameter `pos` sits at torch position `tid`
            st = sd["state"].get(ids[tid])
            if st is None:
                continue                                    # never stepped (or frozen on the saving side)
            a, b = self.offsets[pos]
            for k, buf in zip(keys, self._state):
                if st.get(k) is not None:
                    buf[a:b].copy_(st[k].reshape(-1))
            if "step" in st:
                step = max(step, float(st["step"]))
        if self.method == "sgd" and sd["state"]:
            step = max(step, 1.0)                      # momentum buffers exist: the next step is not the first
        self._step_dev.fill_(step)
        self._lr_host = [None] * len(self.param_groups)
        self.sync_hyper()


def _runs(a, b, holes):
    """[a, b) minus the (sorted, disjoint) intervals of `holes`, as a list of non-empty (start, end)."""
    out, cur = [], a
    for s, e in holes:
        if e <= a or s >= b:
            continue
        if s > cur:
            out.append((cur, s))
        cur = max(cur, e)
    if cur < b:
        out.append((cur, b))
    return out
