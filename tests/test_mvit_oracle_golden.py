"""The MViTv2 oracle (oracle/mvit_oracle.py, BASELINE config 5) against golden vectors of the unmodified reference
(tests/golden/mvit_*.pt, written by oracle/make_golden_mvit.py): geometry and parameter schema incl. the shipped
16 x 224 model, logits / encoder feature / per-block cls rows, loss and every parameter gradient.  CPU only; this pins the
oracle; the sm_100a MViT path itself is held to the same goldens in tests/test_mvit_cpu.py / tests/test_mvit_gpu.py."""
import json
import os

import pytest
import torch

import mvit_oracle as MO

torch.set_num_threads(max(1, os.cpu_count() or 1))
CASES = ["d4_t4_c64", "d3_t8_c96"]


@pytest.mark.parametrize("case", CASES)
def test_forward_backward_match_reference(gold_dir, coin_label_emb, case):
    g = torch.load(os.path.join(gold_dir, f"mvit_{case}.pt"))
    c = g["cfg"]
    geo = MO.geometry(c["mvit"], c["frames"], c["crop"])
    enc_shapes = {k: tuple(v) for k, v in g["shapes"].items() if k.startswith(MO.PRE)}
    assert {k: tuple(v) for k, v in MO.param_shapes(geo).items()} == enc_shapes
    p = {k: v.requires_grad_(True) for k, v in MO.seeded_state(g["shapes"], c["seed"]).items()}
    x = MO.synthetic_clips(c["B"], c["frames"], c["crop"], c["seed"] + 1)
    taps = []
    logits = MO.match_lang_forward(p, x, geo, coin_label_emb, taps=taps)
    torch.testing.assert_close(logits.detach(), g["logits"], rtol=1e-4, atol=1e-4)
    assert torch.equal(logits.argmax(1), g["logits"].argmax(1))
    assert len(taps) == len(g["taps"])
    for t, ref in zip(taps, g["taps"]):
        assert tuple(t.shape) == tuple(ref["shape"])
        torch.testing.assert_close(t[:, 0].detach(), ref["cls"], rtol=1e-4, atol=1e-5)
        assert abs(t.norm().item() - ref["norm"]) <= 1e-4 * ref["norm"]
    loss = torch.nn.functional.cross_entropy(logits, g["labels"])
    assert abs(loss.item() - g["loss"]) <= 1e-4 * abs(g["loss"])
    loss.backward()
    got = {k: v.grad for k, v in p.items() if v.grad is not None}
    assert len(got) == g["n_grads"] and set(got) == set(g["grads"])
    for k, ref in g["grads"].items():
        assert abs(got[k].norm().item() - ref["norm"]) <= 2e-4 * ref["norm"] + 1e-7, k
        torch.testing.assert_close(got[k].flatten()[:32], ref["head"], rtol=2e-3, atol=1e-6 + 1e-4 * ref["norm"], msg=lambda m, k=k: f"{k}: {m}")


def test_full_size_geometry_matches_reference(gold_dir):
    """MViTv2-S 16 x 224 as shipped (procedurevrl_mvitv2_adamw.yaml): widths 96 -> 768, heads 1 -> 8, token grids
    8x56x56 -> 8x7x7, 393 pooled keys (1569 in the transition blocks) -- and the complete parameter schema."""
    with open(os.path.join(gold_dir, "mvit_full_geometry.json")) as f:
        g = json.load(f)
    geo = MO.geometry(g["mvit"], g["frames"], g["crop"])
    assert geo["grid"] == g["patch_dims"] == [8, 56, 56] and geo["out_dim"] == g["out_dim"] == 768
    assert len(geo["blocks"]) == len(g["blocks"]) == 16
    grid = list(geo["grid"])
    for i, (b, r) in enumerate(zip(geo["blocks"], g["blocks"])):
        assert (b["dim"], b["dim_out"], b["heads"]) == (r["dim"], r["dim_out"], r["heads"])
        assert (b["rel_sp"], b["rel_t"]) == (r["rel_sp"], r["rel_t"])
        assert ([b["kernel_q"], b["stride_q"]] if b["kernel_q"] else None) == r["pool_q"]
        assert ([b["kernel_kv"], b["stride_kv"]] if b["kernel_kv"] else None) == r["pool_kv"]
        assert b["dim_out"] // b["heads"] == 96                       # every stage runs 96-wide heads
        kv = [s // st for s, st in zip(b["grid"], b["stride_kv"])]
        # the adaptive KV stride keeps 8x7x7 (+cls) = 393 keys, except in the three stage-transition blocks, whose K / V
        # are pooled from the still un-pooled input grid with the NEW stage's stride: 8x14x14 (+cls) = 1569 keys
        assert 1 + kv[0] * kv[1] * kv[2] == (1569 if i in (1, 3, 14) else 393), i
        assert b["grid"] == grid
        if b["stride_q"]:
            grid = [s // st for s, st in zip(grid, b["stride_q"])]
    assert grid == geo["out_grid"] == [8, 7, 7]
    shapes = MO.param_shapes(geo)
    assert {k: list(v) for k, v in shapes.items()} == g["shapes"]
    assert sum(torch.Size(v).numel() for v in shapes.values()) == g["n_params"]


def test_geometry_rejects_unrestated_modes(gold_dir):
    with open(os.path.join(gold_dir, "mvit_full_geometry.json")) as f:
        mv = json.load(f)["mvit"]
    with pytest.raises(NotImplementedError):
        MO.geometry(dict(mv, POOL_FIRST=True), 16, 224)
