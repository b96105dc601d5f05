"""CPU checks of the C-ABI boundary: the shared library builds/loads without a GPU, exports every symbol that
include/pvrl.h declares, rejects bad arguments with an error string instead of crashing, and the product path
fails loudly (no fallback) when the library or a CUDA device is missing.  No compute call is made."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from procedurevrl_b200 import build, ops
    build.build()
    return ops.lib()


def declared_symbols():
    with open(os.path.join(ROOT, "include", "pvrl.h")) as f:
        h = f.read()
    return sorted(set(re.findall(r"\b(pvrl_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_exported(lib):
    from procedurevrl_b200 import ops
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pvrl.h but not exported by libpvrl_sm100.so"
    assert set(ops.exported_symbols()) == set(names), "ops.py binds a different symbol set than the header declares"
    assert lib.pvrl_abi_version() == 1
    assert lib.pvrl_launch_count() >= 0


def test_bad_arguments_return_errors_not_crashes(lib):
    from procedurevrl_b200 import ops
    assert lib.pvrl_gemm_bf16(None, None) == -1
    assert b"null descriptor" in lib.pvrl_last_error()
    d = ops.GemmDesc()
    d.M, d.N, d.K = 128, 100, 64                       # N not a multiple of 32
    d.A = d.B = d.out = 16
    assert lib.pvrl_gemm_bf16(ctypes.byref(d), None) == -1
    assert b"multiple of 32" in lib.pvrl_last_error()
    d.N, d.trans, d.epilogue = 128, 1, ops.EPI_STORE   # TN form is ATOMIC only
    assert lib.pvrl_gemm_bf16(ctypes.byref(d), None) == -1
    assert lib.pvrl_layernorm_fwd(16, None, 16, 16, 16, 0, None, 8, 100, 1e-6, 0, ops.Geom(1, 1), None) == -1
    assert lib.pvrl_attn_tc_fwd(16, 16, None, 1, 300, 12, 0.125, None) == -1          # seq > 256
    assert lib.pvrl_attn_bwd(16, 16, 16, 16, 16, 0, 0, 500, 12, 0.125, None) == -1    # no sequences
    assert lib.pvrl_attn_tc_bwd(16, 16, 16, 16, 16, 1, 500, 12, 0.125, None) == -1     # seq > 256 on the tcgen05 kernel
    assert lib.pvrl_kl_topk_loss(16, 16, None, None, None, 2, 100, 9, 1.0, None) == -1  # topk > 8


def test_mvit_entry_points_validate_arguments(lib):
    """The MViTv2 entry points (include/pvrl.h, csrc/mvit.cu) reject inconsistent geometry with a message, before any launch."""
    import ctypes as C
    from procedurevrl_b200 import ops
    assert lib.pvrl_ln_any_fwd(16, 1, 16, 16, 16, 1, 16, 8, 2048, 1e-6, None) == -1            # width > 1024
    assert b"1 .. 1024" in lib.pvrl_last_error()
    d = ops._pool_desc(2, 2, 96, (4, 8, 8), (3, 3, 3), (1, 2, 2), (1, 1, 1), 576)
    assert list(d.out) == [4, 4, 4]
    d.out[1] = 5                                                                               # wrong output grid
    assert lib.pvrl_pool3d_fwd(16, 16, 16, 1, C.byref(d), None) == -1
    assert b"output grid" in lib.pvrl_last_error()
    d = ops._pool_desc(2, 2, 256, (4, 8, 8), (3, 3, 3), (1, 2, 2), (1, 1, 1), 3 * 2 * 256)       # more than 128 channels per head
    assert lib.pvrl_pool3d_fwd(16, 16, 16, 1, C.byref(d), None) == -1
    d = ops._pool_desc(2, 2, 96, (4, 8, 8), (3, 3, 3), (1, 2, 2), (1, 1, 1), 576)
    assert lib.pvrl_pool3d_fwd(16, None, 16, 1, C.byref(d), None) == -1                        # no weights but a pooled grid
    assert b"re-layout only" in lib.pvrl_last_error()
    d = ops._pool_desc(1, 2, 96, (2, 8, 8), (1, 3, 3), (1, 2, 2), (0, 1, 1), 96)
    assert lib.pvrl_maxpool3d_fwd(16, 16, 16, 1, C.byref(d), None) == -1                       # heads must be 1
    a = ops.PooledAttnDesc(1, 2, 129, 33, 64, 2, 4, 4, 0.125, 1)
    assert lib.pvrl_pooled_attn_fwd(16, 16, 16, 16, 16, 16, 1, C.byref(a), None) == -1         # head width 64
    assert b"96-wide heads" in lib.pvrl_last_error()
    a = ops.PooledAttnDesc(1, 2, 129, 34, 96, 2, 4, 4, 0.1, 1)
    assert lib.pvrl_pooled_attn_bwd(16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 1, C.byref(a), None) == -1   # key grid != Nk - 1
    assert b"key grid" in lib.pvrl_last_error()
    d = ops._pool_desc(1, 1, 3, (4, 32, 32), (3, 7, 7), (2, 4, 4), (1, 3, 3), 0)
    assert lib.pvrl_im2col3d(16, 16, 0, 3, 440, C.byref(d), None) == -1                        # Kpad shorter than a window (441)


def test_no_cpu_fallback(monkeypatch):
    from procedurevrl_b200 import build, ops
    x = torch.zeros(4, 8)
    with pytest.raises(AssertionError, match="CUDA tensors only"):
        ops.colsum(x, torch.zeros(8), 4, 8)
    # a missing library is an error, never a silent eager path
    monkeypatch.setattr(ops, "_lib", None)
    monkeypatch.setattr(build, "LIB_PATH", "/nonexistent/libpvrl_sm100.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        ops.lib()


def test_model_requires_cuda(gold_dir):
    from procedurevrl_b200.lib.config import get_cfg
    from procedurevrl_b200.lib.models import MODEL_REGISTRY, build_model
    cfg = get_cfg()
    cfg.merge_from_list(["DEV.MATCH_LANG_EMB", True, "DEV.TEST_LANG_EMB", os.path.join(gold_dir, "clip_step_emb_coin.pt"),
                         "MODEL.MODEL_NAME", "vit_base_patch16_224_develop", "MODEL.PRETRAINED", False,
                         "TIMESFORMER.DEPTH", 1, "MODEL.NUM_CLASSES", 778])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            build_model(cfg)
    m = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(cfg)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(torch.zeros(1, 3, 8, 224, 224))
    with pytest.raises(KeyError):
        MODEL_REGISTRY.get("SlowFast")


def test_config_mirror_keys():
    """lib/config keys the hot path reads keep the reference's names and defaults (defaults.py:40-65,380-466)."""
    from procedurevrl_b200.lib.config import get_cfg
    c = get_cfg()
    assert c.DEV.TEMP == 0.02 and c.DEV.ORDER_PRETRAIN_MAX_LEN == 9 and c.DEV.ORDER_TFM_LAYERS == 4
    assert c.TRAIN.TOPK == 5 and c.MODEL.DROP_PATH == 0.1 and c.MODEL.PRETRAINED is True
    assert c.TIMESFORMER.ATTENTION_TYPE == "divided_space_time" and c.TIMESFORMER.DEPTH == 12
    assert c.DATA.NUM_FRAMES == 8 and c.DATA.TRAIN_CROP_SIZE == 224 and c.DIST_BACKEND == "nccl"
    c.merge_from_list(["MODEL.DROP_PATH", "0.0", "NUM_GPUS", 8, "MVIT.POOL_KV_STRIDE_ADAPTIVE", "(1, 8, 8)"])
    assert c.MODEL.DROP_PATH == 0.0 and c.NUM_GPUS == 8 and c.MVIT.POOL_KV_STRIDE_ADAPTIVE == (1, 8, 8)
    c2 = c.clone()
    c2.DEV.TEMP = 0.5
    assert c.DEV.TEMP == 0.02 and "TEMP" in c.dump()
