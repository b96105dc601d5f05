"""Import shims that let the UNMODIFIED reference (`/root/reference`) import in this container.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (golden-vector generator) and by the
optional "restatement == live reference" CPU test.  Nothing in the product package imports this
file, and nothing here runs on the GPU box (`/root/reference` does not exist there).

Why shims are needed:
  * `fvcore`, `yacs`, `clip`, `ipdb`, `simplejson`, `matplotlib`, `tkinter` are not installed;
  * `lib/models/__init__.py:5` imports `video_model_builder.py`, whose line 23 imports a symbol
    that does not exist (`vit_base_patch16_224`), so the package init must be bypassed.

None of the stubs carries arithmetic: they are registries, attribute dicts and no-ops.
"""
import ast
import copy
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PVRL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "models", "vit.py"))


class _Registry(dict):
    """fvcore.common.registry.Registry stand-in (name -> object)."""

    def __init__(self, name="REG"):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self[o.__name__] = o
                return o
            return deco
        self[obj.__name__] = obj
        return obj

    def get(self, name):
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]


class _CfgNode(dict):
    """fvcore/yacs CfgNode stand-in: attribute dict with merge_from_file / merge_from_list."""

    def __init__(self, init=None, new_allowed=False):
        super().__init__()
        if init:
            for k, v in init.items():
                self[k] = _CfgNode(v) if isinstance(v, dict) and not isinstance(v, _CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        return copy.deepcopy(self)

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], dict):
                    self[k] = _CfgNode()
                self[k]._merge(v)
            else:
                if isinstance(v, str):
                    try:  # "(3, 7, 7)" style tuples in the MViT YAMLs
                        lit = ast.literal_eval(v)
                        if isinstance(lit, (tuple, list)):
                            v = lit
                    except Exception:
                        pass
                self[k] = v

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f))

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v

    def dump(self, **kw):
        import yaml
        return yaml.safe_dump(_plain(self))

    def freeze(self):
        pass

    def defrost(self):
        pass


def _plain(n):
    return {k: _plain(v) if isinstance(v, dict) else v for k, v in n.items()}


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _StubTextTower:
    """Stands in for the frozen CLIP text tower (`vit.py:257-261`): the benchmark feeds
    pre-extracted text embeddings (north star), so `encode_text` is a seeded table lookup."""


def install(text_table=None):
    """Register the stub modules and put the reference on sys.path.  Idempotent."""
    import torch
    import torch.nn as nn

    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "ipdb" not in sys.modules:
        _mod("ipdb", set_trace=lambda *a, **k: None)
    if "turtle" not in sys.modules:
        _mod("turtle", distance=lambda *a, **k: 0.0)
    if "simplejson" not in sys.modules:
        import json
        _mod("simplejson", dumps=json.dumps, loads=json.loads)
    if "matplotlib" not in sys.modules:
        mpl = _mod("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = _mod("matplotlib.pyplot")
    if "fvcore" not in sys.modules:
        fv = _mod("fvcore")
        fv.common = _mod("fvcore.common")
        fv.common.registry = _mod("fvcore.common.registry", Registry=_Registry)
        fv.common.config = _mod("fvcore.common.config", CfgNode=_CfgNode)

        class _PM:
            open = staticmethod(open)
            exists = staticmethod(os.path.exists)
            isfile = staticmethod(os.path.isfile)
            ls = staticmethod(os.listdir)
            mkdirs = staticmethod(lambda p: os.makedirs(p, exist_ok=True))
        fv.common.file_io = _mod("fvcore.common.file_io", PathManager=_PM)

        class _Timer:
            def __init__(self):
                self.reset()

            def reset(self):
                import time
                self._t = time.perf_counter()

            def pause(self):
                pass

            def seconds(self):
                import time
                return time.perf_counter() - self._t
        fv.common.timer = _mod("fvcore.common.timer", Timer=_Timer)
        fv.nn = _mod("fvcore.nn")
        fv.nn.weight_init = _mod("fvcore.nn.weight_init", c2_msra_fill=lambda m: None,
                                 c2_xavier_fill=lambda m: None)
        for n in ("activation_count", "flop_count", "precise_bn"):
            setattr(fv.nn, n, _mod("fvcore.nn." + n))
        fv.nn.activation_count.activation_count = lambda *a, **k: ({}, None)
        fv.nn.flop_count.flop_count = lambda *a, **k: ({}, None)

    # --- clip stub: `clip.load` returns (model, None); model.encode_text(ids)->[n,512] ---
    class _ClipStub(nn.Module):
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(1234)
            tab = text_table if text_table is not None else 0.4 * torch.randn(512, 512, generator=g)
            self.register_buffer("table", tab)
            self.visual = nn.Identity()

        def float(self):
            return self

        def encode_text(self, ids):
            # ids [n, 77] int64 -> mean of table rows (mod table size): deterministic, no learned math
            return self.table[ids % self.table.shape[0]].mean(dim=1)

    _mod("clip", load=lambda *a, **k: (_ClipStub(), None),
         tokenize=lambda texts, **k: torch.zeros(len(texts), 77, dtype=torch.long))

    # --- bypass the broken `lib/models/__init__.py` ---
    if "lib" not in sys.modules or not hasattr(sys.modules["lib"], "__path__"):
        lib = types.ModuleType("lib")
        lib.__path__ = [os.path.join(REFERENCE_ROOT, "lib")]
        sys.modules["lib"] = lib
    if "lib.models" not in sys.modules:
        lm = types.ModuleType("lib.models")
        lm.__path__ = [os.path.join(REFERENCE_ROOT, "lib", "models")]
        sys.modules["lib.models"] = lm


def load_reference():
    """Returns (ref_vit_module, ref_build_module, get_cfg) of the unmodified reference."""
    install()
    import importlib
    build = importlib.import_module("lib.models.build")
    vit = importlib.import_module("lib.models.vit")
    defaults = importlib.import_module("lib.config.defaults")
    return vit, build, defaults.get_cfg


def reference_cfg(yaml_rel, overrides=()):
    """cfg for a shipped YAML (relative to the reference root) + KEY VAL overrides, CPU/no-pretrain."""
    _, _, get_cfg = load_reference()
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(REFERENCE_ROOT, yaml_rel))
    cfg.merge_from_list(["NUM_GPUS", 0, "MODEL.PRETRAINED", False] + list(overrides))
    return cfg
