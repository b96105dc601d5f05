"""Config tree with the reference's keys and defaults for everything the hot path reads
(reference lib/config/defaults.py:40-65 DEV.*, :73-106 TRAIN.*, :383-440 MODEL.*, :463-466 TIMESFORMER.*, :169-280 MVIT.*,
:504-525 DATA.*, :14-23 BN.*, :576-627 SOLVER.*, :634-659 NUM_GPUS/NUM_SHARDS/RNG_SEED/DIST_BACKEND/GLOBAL_BATCH_SIZE).

The reference uses fvcore/yacs `CfgNode`; neither is a dependency here, so `CfgNode` below is a small
attribute-dict with the same `merge_from_file / merge_from_list / clone / dump` calls.  Unknown keys found in
a reference YAML (data paths, solver, loaders ...) are kept verbatim so the shipped configs load unchanged."""
import ast
import copy

import yaml


class CfgNode(dict):
    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

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
                if not isinstance(self.get(k), CfgNode):
                    self[k] = CfgNode()
                self[k]._merge(v)
            else:
                self[k] = _coerce(v, self.get(k))

    def merge_from_file(self, path):
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, lst):
        assert len(lst) % 2 == 0, "override list must be KEY VALUE pairs"
        for key, val in zip(lst[0::2], lst[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                if not isinstance(node.get(p), CfgNode):
                    node[p] = CfgNode()
                node = node[p]
            if isinstance(val, str):
                try:
                    val = ast.literal_eval(val)
                except (ValueError, SyntaxError):
                    pass
            node[parts[-1]] = _coerce(val, node.get(parts[-1]))

    def dump(self):
        return yaml.safe_dump(_plain(self))


def _plain(n):
    return {k: _plain(v) if isinstance(v, dict) else v for k, v in n.items()}


def _coerce(v, old):
    if isinstance(v, str) and v.startswith("(") and v.endswith(")"):      # "(3, 7, 7)" style tuples in MViT YAMLs
        try:
            v = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            pass
    if isinstance(old, float) and isinstance(v, int) and not isinstance(v, bool):
        v = float(v)
    return v


_DEFAULTS = {
    "DEV": {
        "ENABLE": False, "LOAD_DUMMY_DATA": False, "CLIP_LINKING": False, "CLIP_VIS_FEAT_PATH": "",
        "CLIP_VIS_FEAT_INPUT": False, "MATCH_LANG_EMB": False, "TEST_LANG_EMB": "", "TEMP": 0.02,
        "ZERO_SHOT_ENABLED": False, "ORDER_PRETRAIN_ENABLED": False, "ORDER_PRETRAIN_MAX_LEN": 9,
        "ORDER_FIX_RECOGNITION": False, "ORDER_STRIDE": 2, "ORDER_TFM_LAYERS": 4, "ORDER_RECOG_BATCH": 9,
    },
    "TRAIN": {"ENABLE": True, "DATASET": "kinetics", "LABEL_EMB": "", "LINEAR": False, "TEXT": "", "TOPK": 5,
              "BATCH_SIZE": 64, "MULT": 1.0},
    "BN": {"WEIGHT_DECAY": 0.0},
    "SOLVER": {"BASE_LR": 0.1, "LR_POLICY": "cosine", "COSINE_END_LR": 0.0, "GAMMA": 0.1, "STEP_SIZE": 1, "STEPS": [],
               "LRS": [], "MAX_EPOCH": 300, "MOMENTUM": 0.9, "DAMPENING": 0.0, "NESTEROV": True, "WEIGHT_DECAY": 1e-4,
               "WARMUP_FACTOR": 0.1, "WARMUP_EPOCHS": 0.0, "WARMUP_START_LR": 0.01, "OPTIMIZING_METHOD": "sgd",
               "BASE_LR_SCALE_NUM_SHARDS": False},
    "TEST": {"ENABLE": True, "DATASET": "kinetics", "BATCH_SIZE": 8},
    "MODEL": {"ARCH": "slowfast", "MODEL_NAME": "SlowFast", "NUM_CLASSES": 400, "LOSS_FUNC": "cross_entropy",
              "DROPOUT_RATE": 0.5, "PRETRAINED": True, "MLP": 0, "TEXT_MODEL": "", "TEXT_LP": False, "NUM_SEG": 0,
              "EXTRA_TR": "", "DROP_E": 0.0, "PRE_CLASSES": 0, "DROP_PATH": 0.1},
    "TIMESFORMER": {"ATTENTION_TYPE": "divided_space_time", "PRETRAINED_MODEL": "", "DEPTH": 12},
    # reference lib/config/defaults.py:169-280 (MViTv2 encoder, MODEL.MODEL_NAME: MViT)
    "MVIT": {"MODE": "conv", "POOL_FIRST": False, "CLS_EMBED_ON": True, "PATCH_KERNEL": [3, 7, 7], "PATCH_STRIDE": [2, 4, 4],
             "PATCH_PADDING": [2, 4, 4], "PATCH_2D": False, "EMBED_DIM": 96, "NUM_HEADS": 1, "MLP_RATIO": 4.0,
             "QKV_BIAS": True, "DROPPATH_RATE": 0.1, "LAYER_SCALE_INIT_VALUE": 0.0, "DEPTH": 16, "NORM": "layernorm",
             "DIM_MUL": [], "HEAD_MUL": [], "POOL_KV_STRIDE": [], "POOL_KV_STRIDE_ADAPTIVE": None, "POOL_Q_STRIDE": [],
             "POOL_KVQ_KERNEL": None, "ZERO_DECAY_POS_CLS": True, "NORM_STEM": False, "SEP_POS_EMBED": False,
             "DROPOUT_RATE": 0.0, "USE_ABS_POS": True, "REL_POS_SPATIAL": False, "REL_POS_TEMPORAL": False,
             "REL_POS_ZERO_INIT": False, "RESIDUAL_POOLING": False, "DIM_MUL_IN_ATT": False, "SEPARATE_QKV": False,
             "HEAD_INIT_SCALE": 1.0, "USE_MEAN_POOLING": False, "USE_FIXED_SINCOS_POS": False},
    "DATA": {"NUM_FRAMES": 8, "MEAN": [0.45, 0.45, 0.45], "STD": [0.225, 0.225, 0.225], "INPUT_CHANNEL_NUM": [3, 3],
             "TRAIN_CROP_SIZE": 224, "TEST_CROP_SIZE": 256},
    # extension node of this implementation (not in the reference): arithmetic mode of the sm_100a path
    #   "bf16"   bf16 tensor-core operands, fp32 accumulate / residual stream (throughput mode)
    #   "bf16x3" error-compensated 3-term bf16 split, ~fp32 products (parity mode)
    "B200": {"PRECISION": "bf16", "GRAD_BUCKET_MB": 64},
    "NUM_GPUS": 1, "NUM_SHARDS": 1, "SHARD_ID": 0, "RNG_SEED": 1, "DIST_BACKEND": "nccl", "GLOBAL_BATCH_SIZE": 64,
    "OUTPUT_DIR": ".",
}


def get_cfg():
    """A fresh copy of the defaults (reference lib/config/defaults.py:1073-1077)."""
    return CfgNode(copy.deepcopy(_DEFAULTS))
