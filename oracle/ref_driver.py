"""Imports the reference's OWN driver loops -- `tools/train_net.py::train_epoch` (:56-248) and
`tools/test_net.py::perform_test` (:32-158) -- unmodified, so that tests can run them on the drop-in model.

TEST INFRASTRUCTURE ONLY (needs /root/reference, i.e. the build container).  On top of `ref_shims.install()`:
  * packages absent here and unused by the two functions are stubbed: `lib.datasets` (decoders / ffmpeg / av),
    `lib.visualization.tensorboard_vis`, `timm.loss`, `fvcore.nn.precise_bn`, `lib.utils.logging` (simplejson);
  * `lib.models` is a stub package whose `build_model` is whatever the test registers (the reference's own
    `lib/models/__init__.py` is broken as shipped);
  * `lib/models/optimizer.py` is a SyntaxError as shipped (line 40): its two broken lines are joined in memory, as in
    `make_golden_optim.py` -- nothing else of the reference is altered;
  * everything the loops actually execute -- lr policy, `set_lr`, `construct_optimizer`, meters, metrics,
    `misc.check_nan_losses`, `distributed` -- is the reference's code.
"""
import importlib
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _exec_module(name, rel, patch=None):
    path = os.path.join(ref_shims.REFERENCE_ROOT, rel)
    with open(path) as f:
        src = f.read()
    if patch is not None:
        src = patch(src)
    mod = types.ModuleType(name)
    mod.__file__ = path
    sys.modules[name] = mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def install(build_model=None):
    ref_shims.install()
    root = ref_shims.REFERENCE_ROOT
    if root not in sys.path:
        sys.path.insert(0, root)
    for pkg in ("lib", "lib.utils", "lib.models", "lib.visualization", "lib.config"):
        if pkg not in sys.modules or not hasattr(sys.modules[pkg], "__path__"):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(root, *pkg.split("."))]
            sys.modules[pkg] = m
    import logging as _pylog
    _stub("lib.utils.logging", get_logger=lambda name: _pylog.getLogger(name), log_json_stats=lambda stats: None,
          setup_logging=lambda *a, **k: None)
    # data pipeline: decoders, ffmpeg, av -- the loops only iterate over the loader object handed to them
    ds = _stub("lib.datasets", loader=_stub("lib.datasets.loader"), Mixup=object)
    ds.__path__ = []
    _stub("lib.datasets.utils", pack_pathway_output=lambda cfg, frames: [frames])
    _stub("lib.visualization.tensorboard_vis", TensorboardWriter=object)
    _stub("lib.models.batchnorm_helper", SubBatchNorm3d=type("SubBatchNorm3d", (torch.nn.Module,), {}))
    sys.modules["lib.models"].build_model = build_model
    if "timm" not in sys.modules:
        _stub("timm").__path__ = []
    _stub("timm.loss", LabelSmoothingCrossEntropy=torch.nn.CrossEntropyLoss, SoftTargetCrossEntropy=torch.nn.CrossEntropyLoss)
    pb = sys.modules.get("fvcore.nn.precise_bn") or _stub("fvcore.nn.precise_bn")
    pb.get_bn_modules = lambda model: []
    pb.update_bn_stats = lambda *a, **k: None
    # optimizer.py:40-41: an `assert cond, ` whose message sits on the next line without a continuation
    _exec_module("lib.models.optimizer", "lib/models/optimizer.py", lambda s: s.replace("+ len(emb), \n", "+ len(emb), \\\n"))


def load_train_net(build_model=None):
    install(build_model)
    return _exec_module("ref_tools_train_net", "tools/train_net.py")


def load_test_net(build_model=None):
    install(build_model)
    return _exec_module("ref_tools_test_net", "tools/test_net.py")


def reference_cfg(yaml_rel, overrides=()):
    defaults = importlib.import_module("lib.config.defaults")
    cfg = defaults.get_cfg()
    cfg.merge_from_file(os.path.join(ref_shims.REFERENCE_ROOT, yaml_rel))
    cfg.merge_from_list(list(overrides))

    def coerce(node):        # yacs literal_evals "1e-4" (PyYAML reads a dot-less exponent as a string); the shim CfgNode does not
        for k, v in list(node.items()):
            if hasattr(v, "items"):
                coerce(v)
            elif isinstance(v, str):
                try:
                    node[k] = float(v)
                except ValueError:
                    pass
    coerce(cfg)
    return cfg
