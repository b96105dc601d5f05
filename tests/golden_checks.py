"""Shared comparison of a gradient tensor with its golden summary (oracle/make_golden.py::grad_summary): Frobenius norm,
the first 64 elements, evenly spaced 64-element windows through the rest of the tensor, and the plain / absolute sums.
The windows and sums catch what norm + head cannot: a permuted tile or a dropped split-K slice far inside a large dW
(VERDICT r1, weak 9).  Older fixtures without `win` / `abs_sum` are still accepted."""
import torch


def check_grad(name, mine, g, rtol, head=64):
    assert mine is not None, name
    mine = mine.detach().float().cpu()
    n = mine.numel()
    norm = g["norm"]
    atol = rtol * norm / n ** 0.5 + 1e-7
    assert abs(mine.norm().item() - norm) <= rtol * norm + 1e-7, (name, mine.norm().item(), norm)
    flat = mine.flatten()
    h = g["head"]
    torch.testing.assert_close(flat[:h.numel()], h, rtol=rtol, atol=atol, msg=lambda m: f"{name} head: {m}")
    if "win" in g and len(g["win_offsets"]):
        got = torch.stack([flat[o:o + g["win"].shape[1]] for o in g["win_offsets"]])
        torch.testing.assert_close(got, g["win"], rtol=rtol, atol=atol, msg=lambda m: f"{name} windows: {m}")
    # sums: n terms with independent errors of size ~ rtol * |g_i| -> tolerance ~ rtol * norm * (a few); the absolute sum is
    # the sharper of the two (no cancellation)
    if "abs_sum" in g:
        assert abs(flat.double().abs().sum().item() - g["abs_sum"]) <= rtol * g["abs_sum"] + 1e-7, name
    assert abs(flat.double().sum().item() - g["sum"]) <= 4 * rtol * norm * n ** 0.5 / 8 + rtol * abs(g["sum"]) + 1e-6, \
        (name, flat.double().sum().item(), g["sum"])


def check_grads(named, gold, rtol):
    assert gold
    for k, g in gold.items():
        check_grad(k, named[k], g, rtol)
