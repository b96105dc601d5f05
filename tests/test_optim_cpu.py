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
