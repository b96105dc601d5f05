"""Optimizer boundary, host logic on CPU: the lr policy and the parameter grouping of
`procedurevrl_b200.lib.models.optimizer` against values produced by the unmodified reference
(tests/golden/optim.json, written by oracle/make_golden_optim.py)."""
import json
import os

import pytest
import torch

from procedurevrl_b200.lib.config import get_cfg
from procedurevrl_b200.lib.models import optimizer as opt


@pytest.fixture(scope="module")
def gold(gold_dir):
    with open(os.path.join(gold_dir, "optim.json")) as f:
        return json.load(f)


def make_cfg(solver, train, bn_wd=0.0):
    c = get_cfg()
    ov = []
    for k, v in solver.items():
        ov += ["SOLVER." + k, v]
    for k, v in train.items():
        ov += ["TRAIN." + k, v]
    c.merge_from_list(ov + ["BN.WEIGHT_DECAY", bn_wd])
    return c


class Skeleton(torch.nn.Module):
    """Parameters carrying the golden file's NAMES (dots kept through a named_parameters override)."""

    def __init__(self, names, shapes, device="cpu"):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        for n, s in zip(names, shapes):
            self.register_parameter(n.replace(".", "__"), torch.nn.Parameter(torch.randn(*s, generator=g).to(device)))

    def named_parameters(self, *a, **k):
        for n, p in super().named_parameters(*a, **k):
            yield n.replace("__", "."), p


def test_lr_policy_matches_reference(gold):
    assert len(gold["lr"]) >= 4
    for case in gold["lr"]:
        cfg = make_cfg(case["solver"], {})
        for e, want in zip(case["epochs"], case["lr"]):
            assert opt.get_epoch_lr(e, cfg) == pytest.approx(want, rel=1e-12, abs=1e-18), (case["solver"], e)


def test_unknown_policy_raises():
    cfg = make_cfg({"LR_POLICY": "exp"}, {})
    with pytest.raises(NotImplementedError):
        opt.get_epoch_lr(0.0, cfg)


def test_parameter_groups_match_reference(gold):
    for case in gold["groups"]:
        cfg = make_cfg(case["solver"], case["train"], case["bn_wd"])
        m = Skeleton(gold["names"], gold["shapes"])
        name_of = {id(p): n for n, p in m.named_parameters()}
        groups = opt.parameter_groups(m, cfg)
        assert len(groups) == len(case["groups"])
        for mine, ref in zip(groups, case["groups"]):
            assert [name_of[id(p)] for p in mine["params"]] == ref["names"]
            assert mine["weight_decay"] == ref["weight_decay"]
            assert mine.get("lr_mult") == ref["lr_mult"]
        assert [n for n, p in m.named_parameters() if not p.requires_grad] == case["frozen"]


def test_set_lr_applies_multipliers():
    class O:
        param_groups = [{"lr": 0.0, "lr_mult": 0.1}, {"lr": 0.0, "lr_mult": 1.0}, {"lr": 0.0}]
    opt.set_lr(O, 0.5)
    assert [g["lr"] for g in O.param_groups] == [0.05, 0.5, 0.5]


def test_runs_skip_holes():
    assert opt._runs(0, 10, []) == [(0, 10)]
    assert opt._runs(0, 10, [(0, 2), (5, 6)]) == [(2, 5), (6, 10)]
    assert opt._runs(4, 10, [(0, 4), (8, 10)]) == [(4, 8)]
    assert opt._runs(3, 3, []) == []


def test_flat_optimizer_needs_cuda(gold):
    m = Skeleton(gold["names"], gold["shapes"])
    with pytest.raises(AssertionError, match="no CPU path"):
        opt.FlatOptimizer(list(m.parameters()), "adamw")


def test_load_pretrained_resizes_and_reports(tmp_path, gold_dir):
    """TimeSformer `load_pretrained` nearest-resizes pos_embed / time_embed (helpers.py:199-218), seeds the temporal
    branch from the spatial one (:220-237) and REPORTS what it could not place instead of dropping it silently."""
    import torch
    from procedurevrl_b200.lib.config import get_cfg
    from procedurevrl_b200.lib.models.vit import VisionTransformer, load_pretrained
    cfg = get_cfg()
    cfg.merge_from_list(["DEV.MATCH_LANG_EMB", True, "DEV.TEST_LANG_EMB", os.path.join(gold_dir, "clip_step_emb_coin.pt")])
    m = VisionTransformer(img_size=32, depth=1, num_frames=4, cfg=cfg, qkv_bias=True)           # 2 x 2 patches + cls
    g = torch.Generator().manual_seed(0)
    ck = {"model." + k: torch.randn(v.shape, generator=g) for k, v in m.state_dict().items() if "temporal" not in k}
    ck["model.pos_embed"] = torch.randn(1, 1 + 16, 768, generator=g)          # a 4 x 4 grid checkpoint
    ck["model.time_embed"] = torch.randn(1, 8, 768, generator=g)
    ck["model.head.weight"] = torch.randn(1000, 768, generator=g)             # ImageNet classifier: cannot be placed
    path = str(tmp_path / "ck.pth")
    torch.save({"model_state": ck}, path)
    load_pretrained(m, path, num_frames=4)
    F = torch.nn.functional
    want = torch.cat((ck["model.pos_embed"][:, :1],
                      F.interpolate(ck["model.pos_embed"][:, 1:].transpose(1, 2), size=4, mode="nearest").transpose(1, 2)), 1)
    assert torch.equal(m.pos_embed.detach(), want)
    assert torch.equal(m.time_embed.detach(), F.interpolate(ck["model.time_embed"].transpose(1, 2), size=4, mode="nearest").transpose(1, 2))
    assert torch.equal(m.blocks[0].temporal_attn.qkv.weight.detach(), ck["model.blocks.0.attn.qkv.weight"])
    assert torch.equal(m.blocks[0].temporal_norm1.weight.detach(), ck["model.blocks.0.norm1.weight"])
    assert [k for k, _ in m.pretrained_skipped] == ["head.weight"]


def test_mvit_load_pretrained_converts_image_checkpoint(tmp_path, gold_dir):
    """the released MViTv2 *image* checkpoint is converted as reference helpers.py:127-145 does -- 2-D pooling /
    stem kernels repeated over the temporal extent, rel_pos tables linearly interpolated -- never silently skipped."""
    import json
    import torch
    from procedurevrl_b200.lib.config import get_cfg
    from procedurevrl_b200.lib.models import MODEL_REGISTRY
    from procedurevrl_b200.lib.models.mvit import load_pretrained
    from test_mvit_cpu import mvit_cfg
    c = torch.load(os.path.join(gold_dir, "mvit_d4_t4_c64.pt"))["cfg"]
    m = MODEL_REGISTRY.get("MViT")(mvit_cfg(gold_dir, c["mvit"], c["frames"], c["crop"], "bf16"))
    own = m.model.state_dict()
    g = torch.Generator().manual_seed(1)
    ck, expect = {}, {}
    for k, v in own.items():
        if not k.startswith("video_encoder."):
            continue
        name = k[len("video_encoder."):]
        if ("pool_" in name or name == "patch_embed.proj.weight") and v.dim() == 5:
            w2 = torch.randn(v.shape[0], v.shape[1], v.shape[3], v.shape[4], generator=g)          # a 2-D kernel
            ck[name], expect[k] = w2, w2.unsqueeze(2).repeat(1, 1, v.shape[2], 1, 1)
        elif "rel_pos_" in name:
            t = torch.randn(v.shape[0] + 6, v.shape[1], generator=g)                               # a longer table
            ck[name] = t
            expect[k] = torch.nn.functional.interpolate(t.t().unsqueeze(0), size=v.shape[0], mode="linear")[0].t()
        else:
            ck[name] = expect[k] = torch.randn(v.shape, generator=g)
    ck["head.projection.weight"] = torch.randn(1000, 768, generator=g)
    path = str(tmp_path / "mvit_in1k.pyth")
    torch.save({"model_state": ck}, path)
    load_pretrained(m.model, path)
    got = m.model.state_dict()
    for k, v in expect.items():
        torch.testing.assert_close(got[k], v, rtol=0, atol=1e-6, msg=k)
    assert [k for k, _ in m.model.pretrained_skipped] == ["head.projection.weight"]


def test_split_runs_tile_the_flat_buffer():
    """FlatOptimizer.step(split=(offset, hook)) updates [offset, end) before the hook and [0, offset) after it: the two launch
    lists cover every element of every run exactly once, on the right side of the offset, group by group (the trainer hides
    the last gradient exchange -- embeddings + block 0, the front of the buffer -- under the first list)."""
    import random
    from procedurevrl_b200.lib.models.optimizer import split_runs
    rnd = random.Random(0)
    for _ in range(200):
        cuts = sorted(rnd.sample(range(1, 400), rnd.randint(1, 6)))
        runs, start = [], 0
        for gi, c in enumerate(cuts):                    # contiguous groups, some with a hole (a parameter without gradient)
            if rnd.random() < 0.3 and c - start > 4:
                hole = rnd.randint(start + 1, c - 2)
                runs += [(gi, start, hole), (gi, hole + 1, c)]
            else:
                runs.append((gi, start, c))
            start = c
        off = rnd.choice([0, 1, cuts[0], cuts[-1], cuts[-1] + 5, rnd.randint(0, 400)])
        above, below = split_runs(runs, off)
        assert all(s >= off and e > s for _, s, e in above) and all(e <= off and e > s for _, s, e in below)
        cover = {}
        for g, s, e in above + below:
            for i in range(s, e):
                assert i not in cover
                cover[i] = g
        want = {i: g for g, s, e in runs for i in range(s, e)}
        assert cover == want
