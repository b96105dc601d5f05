"""Flat optimizer on the GPU: `pvrl_adam_flat` / `pvrl_sgd_flat` against torch.optim on the same tensors, and
`FlatOptimizer` built by `construct_optimizer` against parameter trajectories of the unmodified reference's optimizer
(tests/golden/optim.json).  fp32 arithmetic, tolerance rtol 2e-6 per element vs torch, 1e-5 on the golden checksums."""
import json
import os

import pytest
import torch

from test_optim_cpu import Skeleton, make_cfg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from procedurevrl_b200 import ops as O
    O.lib()
    return O


def _flat(n, seed, scale=1.0, offset=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(n + offset, generator=g, device="cuda") * scale)[offset:]


@pytest.mark.parametrize("n,offset", [(1, 0), (7, 1), (1000, 3), (4096, 0), (5_000_003, 2)])
@pytest.mark.parametrize("decoupled", [True, False])
def test_adam_flat_vs_torch(ops, n, offset, decoupled):
    p, g0 = _flat(n, 1, offset=offset), _flat(n, 2, 0.1, offset=offset)
    m, v = torch.zeros(n + offset, device="cuda")[offset:], torch.zeros(n + offset, device="cuda")[offset:]
    q = torch.nn.Parameter(p.clone())
    cls = torch.optim.AdamW if decoupled else torch.optim.Adam
    ref = cls([q], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    lr, step = torch.full((1,), 1e-3, device="cuda"), torch.zeros(1, device="cuda")
    for it in range(5):
        g = g0 * (1.0 + it)
        lr_now = 1e-3 * (1 + it)
        ref.param_groups[0]["lr"] = lr_now
        q.grad = g.clone()
        ref.step()
        lr.fill_(lr_now)
        gg = torch.empty(n + offset, device="cuda")[offset:].copy_(g * 4.0)   # grad_scale 0.25 undoes this (a SUM all-reduce over 4 ranks)
        ops.optim_tick(step)
        ops.adam_flat(p, gg, m, v, lr, step, weight_decay=0.05, decoupled=decoupled, grad_scale=0.25,
                      zero_grad=it % 2 == 0)
        assert (gg == 0).all() if it % 2 == 0 else torch.equal(gg, g * 4.0)
        torch.testing.assert_close(p, q.detach(), rtol=2e-6, atol=1e-7)
    st = ref.state[q]
    torch.testing.assert_close(m, st["exp_avg"], rtol=2e-6, atol=1e-8)      # m ~ 1e-2; cancellation near zero in Adam's L2 mode
    torch.testing.assert_close(v, st["exp_avg_sq"], rtol=2e-6, atol=1e-12)
    assert step.item() == 5.0


@pytest.mark.parametrize("momentum,dampening,nesterov", [(0.9, 0.0, True), (0.9, 0.2, False), (0.0, 0.0, False)])
def test_sgd_flat_vs_torch(ops, momentum, dampening, nesterov):
    n, offset = 300_001, 1
    p, g0 = _flat(n, 3, offset=offset), _flat(n, 4, 0.1, offset=offset)
    buf = torch.full((n + offset,), float("nan"), device="cuda")[offset:]      # step 1 must not read the buffer
    q = torch.nn.Parameter(p.clone())
    ref = torch.optim.SGD([q], lr=0.01, momentum=momentum, dampening=dampening, nesterov=nesterov, weight_decay=1e-2)
    lr, step = torch.full((1,), 0.01, device="cuda"), torch.zeros(1, device="cuda")
    for it in range(4):
        g = g0 * (1.0 - 0.3 * it)
        q.grad = g.clone()
        ref.step()
        gg = torch.empty(n + offset, device="cuda")[offset:].copy_(g)
        ops.optim_tick(step)
        ops.sgd_flat(p, gg, buf, lr, step, momentum=momentum, dampening=dampening, nesterov=nesterov,
                     weight_decay=1e-2, zero_grad=True)
        assert (gg == 0).all()
        torch.testing.assert_close(p, q.detach(), rtol=2e-6, atol=1e-7)


def test_bad_arguments_are_rejected(ops):
    p, lr, step = torch.zeros(16, device="cuda"), torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    with pytest.raises(RuntimeError, match="share their offset"):
        ops.adam_flat(p[1:9], p[2:10].clone()[0:8], torch.zeros(9, device="cuda")[1:], torch.zeros(8, device="cuda"), lr, step)
    with pytest.raises(RuntimeError, match="Nesterov"):
        ops.sgd_flat(p, p.clone(), p.clone(), lr, step, momentum=0.0, nesterov=True)


def test_flat_optimizer_vs_reference_trajectories(ops, gold_dir):
    """construct_optimizer -> FlatOptimizer on the golden skeleton: same groups, same 4-step parameter trajectory as the
    reference's torch optimizer (set_lr before every step, one parameter without a gradient on the last step)."""
    from procedurevrl_b200.lib.models import optimizer as opt
    with open(os.path.join(gold_dir, "optim.json")) as f:
        gold = json.load(f)
    for case, traj in zip(gold["groups"], gold["trajectories"]):
        cfg = make_cfg(case["solver"], case["train"], case["bn_wd"])
        m = Skeleton(gold["names"], gold["shapes"], device="cuda")
        o = opt.construct_optimizer(m, cfg)
        assert [len(g["params"]) for g in o.param_groups] == [len(g["names"]) for g in case["groups"]]   # frozen ones kept
        assert len(o._params) == sum(len([n for n in g["names"] if n not in case["frozen"]]) for g in case["groups"])
        gg = torch.Generator().manual_seed(7)
        for it in range(4):
            opt.set_lr(o, cfg.SOLVER.BASE_LR * (1.0 - 0.2 * it))
            o.zero_grad()
            for n, p in m.named_parameters():
                if p.requires_grad and not (it == 3 and "head_cls" in n):
                    if p.grad is None:                 # dropped on an earlier iteration: hand a fresh tensor, as autograd would
                        p.grad = torch.zeros_like(p)
                    p.grad.copy_(torch.randn(p.shape, generator=gg) * 0.1)
                else:
                    p.grad = None
            o.step()
            for n, p in m.named_parameters():
                s, a = traj[it][n]
                assert p.double().sum().item() == pytest.approx(s, rel=1e-5, abs=1e-6), (case["solver"], it, n)
                assert p.double().abs().sum().item() == pytest.approx(a, rel=1e-5), (case["solver"], it, n)


def test_flat_optimizer_state_dict_round_trip(ops):
    """state_dict uses torch.optim's layout: it loads into torch.optim.AdamW and back, and both continue identically."""
    from procedurevrl_b200.lib.models.optimizer import FlatOptimizer
    torch.manual_seed(0)
    ws = [torch.nn.Parameter(torch.randn(33, 7, device="cuda")), torch.nn.Parameter(torch.randn(5, device="cuda"))]
    o = FlatOptimizer([{"params": [ws[0]], "weight_decay": 0.1}, {"params": [ws[1]], "weight_decay": 0.0}], "adamw", lr=1e-2)
    assert ws[0].data_ptr() == o.flat_param.data_ptr() and ws[0].grad.data_ptr() == o.flat_grad.data_ptr()
    for it in range(3):
        for w in ws:
            w.grad.copy_(torch.randn_like(w))
        o.step(zero_grad=True)
        assert (o.flat_grad == 0).all()
    sd = o.state_dict()
    twins = [torch.nn.Parameter(w.detach().clone()) for w in ws]
    t = torch.optim.AdamW([{"params": [twins[0]], "weight_decay": 0.1}, {"params": [twins[1]], "weight_decay": 0.0}], lr=1e-2)
    t.load_state_dict(sd)
    o2 = FlatOptimizer([{"params": [ws[0]], "weight_decay": 0.1}, {"params": [ws[1]], "weight_decay": 0.0}], "adamw", lr=1e-2)
    o2.load_state_dict(t.state_dict())
    for it in range(2):
        gs = [torch.randn_like(w) for w in ws]
        for w, tw, g in zip(ws, twins, gs):
            w.grad.copy_(g)
            tw.grad = g.clone()
        o2.step()
        t.step()
        for w, tw in zip(ws, twins):
            torch.testing.assert_close(w.detach(), tw.detach(), rtol=2e-6, atol=1e-7)


def test_flat_optimizer_in_cuda_graph(ops):
    """The step replays from a CUDA graph while set_lr changes the rate between replays (device-resident lr / step)."""
    from procedurevrl_b200.lib.models import optimizer as opt
    w = torch.nn.Parameter(torch.randn(1000, device="cuda"))
    tw = torch.nn.Parameter(w.detach().clone())
    o = opt.FlatOptimizer([{"params": [w], "lr_mult": 0.5}], "adamw", lr=1e-2, weight_decay=0.01)
    t = torch.optim.AdamW([tw], lr=1e-2, weight_decay=0.01)
    g_static = torch.zeros(1000, device="cuda")
    scratch = torch.nn.Parameter(torch.zeros(1000, device="cuda"))        # loads the kernels before the capture
    opt.FlatOptimizer([scratch], "adamw").step(zero_grad=True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=s):
            w.grad.copy_(g_static)
            o.step(zero_grad=True)
    torch.cuda.current_stream().wait_stream(s)
    o._step_dev.zero_()                        # (capture does not execute; make the intent explicit)
    for it in range(4):
        lr = 1e-2 / (it + 1)
        opt.set_lr(o, lr)
        o.sync_hyper()
        g = torch.randn(1000, device="cuda")
        g_static.copy_(g)
        graph.replay()
        t.param_groups[0]["lr"] = lr * 0.5
        tw.grad = g.clone()
        t.step()
        torch.testing.assert_close(w.detach(), tw.detach(), rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("off", [0, 12, 231, 236, 10 ** 6])
def test_flat_optimizer_split_step_equals_unsplit(ops, off):
    """step(split=(offset, hook)) -- the update above `offset` first, the hook (the trainer's wait for the last gradient
    exchange), then the rest -- equals the plain step bit for bit, across group boundaries and unaligned offsets."""
    from procedurevrl_b200.lib.models.optimizer import FlatOptimizer
    torch.manual_seed(1)
    shapes = [(33, 7), (5,), (4, 4, 4)]

    def make():
        torch.manual_seed(2)
        ws = [torch.nn.Parameter(torch.randn(*sh, device="cuda")) for sh in shapes]
        o = FlatOptimizer([{"params": ws[:2], "weight_decay": 0.1}, {"params": ws[2:], "weight_decay": 0.0}], "adamw", lr=1e-2)
        return ws, o
    (wa, oa), (wb, ob) = make(), make()
    calls = []
    for it in range(3):
        for a, b in zip(wa, wb):
            g = torch.randn_like(a)
            a.grad.copy_(g), b.grad.copy_(g)
        oa.step(zero_grad=True)
        ob.step(zero_grad=True, split=(off, lambda: calls.append(it)))
        assert torch.equal(oa.flat_param, ob.flat_param) and torch.equal(oa._state[0], ob._state[0])
        assert torch.equal(oa._state[1], ob._state[1]) and (ob.flat_grad == 0).all()
    assert calls == [0, 1, 2]


def test_flat_optimizer_foreign_and_missing_grads(ops):
    """`.grad`s that a wrapper re-pointed (DistributedDataParallel's bucket views) are adopted by step() and cleared by
    zero_grad(); a `.grad` set to None skips that parameter for the step (torch semantics) and gets its view back; a
    parameter moved out of the flat buffer is reported instead of silently ignored."""
    from procedurevrl_b200.lib.models.optimizer import FlatOptimizer
    torch.manual_seed(1)
    ws = [torch.nn.Parameter(torch.randn(40, 3, device="cuda")), torch.nn.Parameter(torch.randn(17, device="cuda"))]
    tw = [torch.nn.Parameter(w.detach().clone()) for w in ws]
    o = FlatOptimizer(ws, "adamw", lr=1e-2, weight_decay=0.1)
    t = torch.optim.AdamW(tw, lr=1e-2, weight_decay=0.1)
    bucket = torch.zeros(40 * 3, device="cuda")
    ws[0].grad = bucket.view(40, 3)                      # what gradient_as_bucket_view does
    for it in range(3):
        g0, g1 = torch.randn(40, 3, device="cuda"), torch.randn(17, device="cuda")
        o.zero_grad()
        t.zero_grad()
        assert (bucket == 0).all()
        ws[0].grad.add_(g0)
        tw[0].grad = g0.clone()
        if it == 1:
            ws[1].grad = None                            # no gradient this step: torch leaves the parameter alone
            tw[1].grad = None
        else:
            ws[1].grad.add_(g1)
            tw[1].grad = g1.clone()
        before = ws[1].detach().clone()
        o.step()
        t.step()
        if it == 1:
            assert torch.equal(ws[1].detach(), before)
        torch.testing.assert_close(ws[0].detach(), tw[0].detach(), rtol=2e-6, atol=1e-7)
    # parameter 1 took 2 steps in torch (its own counter) but the flat optimizer counts 3: compare parameter 0 only above
    o.zero_grad()
    assert ws[1].grad is not None and ws[1].grad.data_ptr() == o.flat_grad.data_ptr() + 4 * 120
    ws[0].data = ws[0].data.clone()
    with pytest.raises(RuntimeError, match="no longer lives in the flat buffer"):
        o.step()


def test_flat_optimizer_loads_reference_grouping_with_frozen_parameters(ops):
    """the reference's groups always contain frozen parameters (`head.*` in every fine-tuning config, the CLIP
    tower in pre-training) and torch.optim keeps them in `param_groups` and in its integer numbering.  A state_dict saved by
    torch.optim.AdamW on that grouping must load into FlatOptimizer (and back), and both must continue identically."""
    from procedurevrl_b200.lib.models.optimizer import FlatOptimizer
    torch.manual_seed(1)
    shapes = [(6, 5), (5,), (4, 6), (4,), (7,)]                 # encoder w, b | frozen head w, b | head_cls b
    frozen = {2, 3}

    def make():
        g = torch.Generator().manual_seed(2)
        ps = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
        for i in frozen:
            ps[i].requires_grad_(False)
        return ps
    tw, fw = make(), make()
    groups = lambda ps: [{"params": ps[:2], "weight_decay": 0.0, "lr_mult": 0.1},     # noqa: E731
                         {"params": ps[2:], "weight_decay": 0.05, "lr_mult": 1.0}]
    t = torch.optim.AdamW(groups(tw), lr=1e-2)
    gg = torch.Generator().manual_seed(3)
    for _ in range(2):
        for p in tw:
            p.grad = torch.randn(p.shape, generator=gg).cuda() if p.requires_grad else None
        t.step()
    sd = t.state_dict()
    assert [len(g["params"]) for g in sd["param_groups"]] == [2, 3] and sorted(sd["state"]) == [0, 1, 4]
    with torch.no_grad():
        for a, b in zip(fw, tw):
            a.copy_(b)
    o = FlatOptimizer(groups(fw), "adamw", lr=1e-2)
    assert [len(g["params"]) for g in o.param_groups] == [2, 3] and o._ids == [0, 1, 4]
    o.load_state_dict(sd)                                        # raised "parameter groups don't match" before
    back = o.state_dict()
    assert sorted(back["state"]) == [0, 1, 4] and [g["params"] for g in back["param_groups"]] == [[0, 1], [2, 3, 4]]
    t2 = torch.optim.AdamW(groups(tw), lr=1e-2)
    t2.load_state_dict(back)                                     # and a checkpoint saved here loads on the reference side
    for _ in range(2):
        gs = [torch.randn(p.shape, generator=gg).cuda() for p in tw]
        for p, q, g in zip(tw, fw, gs):
            if p.requires_grad:
                p.grad = g.clone()
                q.grad.copy_(g)
        t.step()
        o.step()
    for p, q in zip(tw, fw):
        torch.testing.assert_close(q.detach(), p.detach(), rtol=2e-6, atol=1e-7)
