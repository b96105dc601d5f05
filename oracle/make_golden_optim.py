"""Golden vectors for the optimizer boundary, produced by the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY (runs in the build container, where /root/reference exists).  Writes
tests/golden/optim.json:
  * lr: `lib/utils/lr_policy.get_lr_at_epoch` on the SOLVER settings of the shipped HowTo100M / COIN YAMLs and a
    steps_with_relative_lrs setting, at a grid of fractional epochs;
  * groups: `lib/models/optimizer.construct_optimizer` run on a skeleton module whose parameter NAMES cover every
    rule (head / order / bn / text_model / encoder) under pretrain, TRAIN.MULT != 1, TRAIN.MULT == 0 and TRAIN.LINEAR
    settings -- the parameter names per group, each group's weight decay / lr / lr_mult, which parameters were frozen;
  * trajectories: 4 steps of that optimizer (adamw, adam, sgd+nesterov, sgd with dampening) on seeded gradients with
    `set_lr` changing the rate each step: parameter checksums after every step.

    python oracle/make_golden_optim.py
"""
import json
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "optim.json")
NAMES = ["model.cls_token", "model.patch_embed.proj.weight", "model.blocks.0.attn.qkv.weight", "model.blocks.0.norm1.bias",
         "model.bn_like.weight", "model.head.weight", "model.head.bias", "model.head_cls.weight",
         "model.order_tfm.layers.0.weight", "model.text_model.proj"]
SHAPES = [(1, 1, 8), (8, 3, 2, 2), (24, 8), (8,), (8,), (4, 8), (4,), (5, 4), (6, 6), (4, 4)]


def load_ref():
    """lr_policy.py and optimizer.py are plain modules: load them by path under the package names they import."""
    root = ref_shims.REFERENCE_ROOT
    for pkg in ("lib", "lib.utils", "lib.models"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(root, *pkg.split("."))]
            sys.modules[pkg] = m
    mods = {}
    for name, rel in (("lib.utils.lr_policy", "lib/utils/lr_policy.py"), ("lib.models.optimizer", "lib/models/optimizer.py")):
        path = os.path.join(root, rel)
        with open(path) as f:
            src = f.read()
        # optimizer.py:40-41 is a SyntaxError as shipped (an `assert cond, ` whose message sits on the next line without a
        # continuation), so the module cannot be imported at all; join those two lines IN MEMORY -- nothing else changes.
        src = src.replace("+ len(emb), \n", "+ len(emb), \\\n")
        mod = types.ModuleType(name)
        mod.__file__ = path
        sys.modules[name] = mod
        exec(compile(src, path, "exec"), mod.__dict__)
        mods[name] = mod
    return mods["lib.utils.lr_policy"], mods["lib.models.optimizer"]


class Skeleton(torch.nn.Module):
    def __init__(self, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self._names = NAMES
        for n, s in zip(NAMES, SHAPES):
            self.register_parameter(n.replace(".", "__"), torch.nn.Parameter(torch.randn(*s, generator=g)))

    def named_parameters(self, *a, **k):
        for n, p in super().named_parameters(*a, **k):
            yield n.replace("__", "."), p


def cfg_of(solver, train, bn_wd=0.0):
    C = ref_shims._CfgNode
    base = dict(BASE_LR=0.1, LR_POLICY="cosine", COSINE_END_LR=0.0, STEPS=[], LRS=[], MAX_EPOCH=300, MOMENTUM=0.9,
                DAMPENING=0.0, NESTEROV=True, WEIGHT_DECAY=1e-4, WARMUP_EPOCHS=0.0, WARMUP_START_LR=0.01,
                OPTIMIZING_METHOD="sgd")
    base.update(solver)
    t = dict(MULT=1.0, LINEAR=False)
    t.update(train)
    return C({"SOLVER": base, "TRAIN": t, "BN": {"WEIGHT_DECAY": bn_wd}})


def main():
    lrp, opt = load_ref()
    gold = {"names": NAMES, "shapes": SHAPES, "lr": [], "groups": [], "trajectories": []}
    lr_cases = [
        dict(BASE_LR=0.00005, LR_POLICY="cosine", MAX_EPOCH=4, WARMUP_EPOCHS=0.0),                  # procedurevrl_adamw.yaml
        dict(BASE_LR=0.005, LR_POLICY="cosine", MAX_EPOCH=15, WARMUP_EPOCHS=1.0, WARMUP_START_LR=0.0005),
        dict(BASE_LR=0.005, LR_POLICY="steps_with_relative_lrs", STEPS=[0, 11, 14], LRS=[1, 0.1, 0.01], MAX_EPOCH=15),
        dict(BASE_LR=0.1, LR_POLICY="cosine", COSINE_END_LR=0.001, MAX_EPOCH=30, WARMUP_EPOCHS=2.5),
    ]
    for s in lr_cases:
        cfg = cfg_of(s, {})
        epochs = [cfg.SOLVER.MAX_EPOCH * i / 23.0 for i in range(24)] + [0.0, 0.5, 1.0, 2.49, 2.5, 11.0, 13.999, 14.0]
        epochs = [e for e in epochs if e <= cfg.SOLVER.MAX_EPOCH]
        gold["lr"].append({"solver": s, "epochs": epochs, "lr": [lrp.get_lr_at_epoch(cfg, e) for e in epochs]})

    group_cases = [({"OPTIMIZING_METHOD": "adamw", "BASE_LR": 5e-5}, {}, 0.0),
                   ({"OPTIMIZING_METHOD": "sgd", "BASE_LR": 5e-3}, {"MULT": 0.1}, 0.01),
                   ({"OPTIMIZING_METHOD": "adamw", "BASE_LR": 5e-5}, {"MULT": 0.0}, 0.0),
                   ({"OPTIMIZING_METHOD": "sgd", "BASE_LR": 5e-3}, {"LINEAR": True}, 0.0),
                   ({"OPTIMIZING_METHOD": "adam", "BASE_LR": 1e-3}, {}, 0.0),
                   ({"OPTIMIZING_METHOD": "sgd", "BASE_LR": 5e-3, "NESTEROV": False, "DAMPENING": 0.2}, {}, 0.0)]
    for solver, train, bn_wd in group_cases:
        cfg = cfg_of(solver, train, bn_wd)
        m = Skeleton()
        name_of = {id(p): n for n, p in m.named_parameters()}
        stdout = sys.stdout
        sys.stdout = open(os.devnull, "w")            # the reference prints the groups
        try:
            o = opt.construct_optimizer(m, cfg)
        finally:
            sys.stdout = stdout
        groups = [{"names": [name_of[id(p)] for p in g["params"]], "weight_decay": g["weight_decay"], "lr": g["lr"],
                   "lr_mult": g.get("lr_mult")} for g in o.param_groups]
        frozen = [n for n, p in m.named_parameters() if not p.requires_grad]
        gold["groups"].append({"solver": solver, "train": train, "bn_wd": bn_wd, "groups": groups, "frozen": frozen})
        # a short trajectory: seeded gradients, the rate changed through set_lr before every step (train_net.py:123-124)
        gg = torch.Generator().manual_seed(7)
        sums = []
        for it in range(4):
            opt.set_lr(o, cfg.SOLVER.BASE_LR * (1.0 - 0.2 * it))
            o.zero_grad()
            for n, p in m.named_parameters():
                if p.requires_grad and not (it == 3 and "head_cls" in n):
                    p.grad = torch.randn(p.shape, generator=gg) * 0.1
                else:
                    p.grad = None                       # head_cls misses the last step: torch skips parameters without a gradient
            o.step()
            sums.append({n: [p.double().sum().item(), p.double().abs().sum().item()] for n, p in m.named_parameters()})
        gold["trajectories"].append(sums)
    with open(OUT, "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
