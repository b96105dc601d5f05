"""tc_linear / tc_mlp (procedurevrl_b200/tc_functional.py): autograd Functions over the tcgen05 GEMM.
CPU: host logic with the torch restatements of the ops (tests/shadow_ops.py) -- operand caches, split-K dW, both
precisions; GPU (`-m gpu`): the real kernels, at MViTv2 layer shapes, against torch fp32 autograd.
Tolerances: "bf16" 1.5e-2 of the output range (operands rounded to bf16), "bf16x3" 2e-4."""
import pytest
import torch

import shadow_ops
from procedurevrl_b200 import ops as real_ops
from procedurevrl_b200 import tc_functional as TC

SHAPES = [((2, 393), 96, 288), ((786,), 192, 192), ((3, 50), 384, 1152), ((1000,), 768, 96)]


def _run(dev, lead, K, N, precision, tol):
    g = torch.Generator().manual_seed(K + N)
    x = torch.randn(*lead, K, generator=g).to(dev).requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev).requires_grad_(True)
    b = (0.1 * torch.randn(N, generator=g)).to(dev).requires_grad_(True)
    dy = torch.randn(*lead, N, generator=g).to(dev)
    ref = torch.nn.functional.linear(x, w, b)
    gx, gw, gb = torch.autograd.grad(ref, (x, w, b), dy)
    y = TC.tc_linear(x, w, b, precision=precision)
    assert y.shape == ref.shape and y.dtype == (torch.float32 if precision == "bf16x3" else torch.bfloat16)
    hx, hw, hb = torch.autograd.grad(y, (x, w, b), dy.to(y.dtype))
    for name, got, want in (("y", y, ref), ("dx", hx, gx), ("dw", hw, gw), ("db", hb, gb)):
        err = (got.float() - want).abs().max().item() / (want.abs().max().item() + 1e-12)
        assert err < tol, (name, lead, K, N, precision, err)
    # the cached operand copies follow the parameter: an in-place update must be seen by the next call
    with torch.no_grad():
        w.mul_(2.0)
    y2 = TC.tc_linear(x, w, b, precision=precision)
    ref2 = torch.nn.functional.linear(x, w, b)
    assert (y2.float() - ref2).abs().max().item() / ref2.abs().max().item() < tol


def _run_mlp(dev, M, K, precision, tol):
    g = torch.Generator().manual_seed(M + K)
    Hd = 4 * K
    x = torch.randn(M, K, generator=g).to(dev).requires_grad_(True)
    w1 = (torch.randn(Hd, K, generator=g) / K ** 0.5).to(dev).requires_grad_(True)
    b1 = (0.1 * torch.randn(Hd, generator=g)).to(dev).requires_grad_(True)
    w2 = (torch.randn(K, Hd, generator=g) / Hd ** 0.5).to(dev).requires_grad_(True)
    b2 = (0.1 * torch.randn(K, generator=g)).to(dev).requires_grad_(True)
    dy = torch.randn(M, K, generator=g).to(dev)
    F = torch.nn.functional
    ref = F.linear(F.gelu(F.linear(x, w1, b1)), w2, b2)
    want = torch.autograd.grad(ref, (x, w1, b1, w2, b2), dy)
    y = TC.tc_mlp(x, w1, b1, w2, b2, precision=precision)
    got = torch.autograd.grad(y, (x, w1, b1, w2, b2), dy.to(y.dtype))
    for name, a, r in zip(("y", "dx", "dw1", "db1", "dw2", "db2"), (y,) + got, (ref,) + want):
        err = (a.float() - r).abs().max().item() / (r.abs().max().item() + 1e-12)
        assert err < tol, (name, M, K, precision, err)


@pytest.fixture
def shadow(monkeypatch):
    for n in shadow_ops.ALL:
        monkeypatch.setattr(real_ops, n, getattr(shadow_ops, n))
    TC._WCACHE.clear()


@pytest.mark.parametrize("precision,tol", [("bf16", 1.5e-2), ("bf16x3", 2e-4)])
def test_tc_linear_host_logic_cpu(shadow, precision, tol):
    for lead, K, N in SHAPES[:3]:
        _run(torch.device("cpu"), lead, K, N, precision, tol)


@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("bf16x3", 2e-4)])
def test_tc_mlp_host_logic_cpu(shadow, precision, tol):
    _run_mlp(torch.device("cpu"), 300, 96, precision, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("bf16", 1.5e-2), ("bf16x3", 2e-4)])
def test_tc_linear_gpu(precision, tol):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    TC._WCACHE.clear()
    for lead, K, N in SHAPES + [((25089,), 96, 288)]:
        _run(torch.device("cuda"), lead, K, N, precision, tol)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("bf16x3", 2e-4)])
def test_tc_mlp_gpu(precision, tol):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    TC._WCACHE.clear()
    for M, K in ((1569, 384), (6273, 192), (393, 768)):
        _run_mlp(torch.device("cuda"), M, K, precision, tol)
