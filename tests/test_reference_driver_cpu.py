"""The north star's "tools/train_net.py and tools/test_net.py drive it unchanged", executed: the reference's OWN
`train_epoch` (tools/train_net.py:56-248) and `perform_test` (tools/test_net.py:32-158) -- imported unmodified through
oracle/ref_driver.py -- run on the drop-in model with the reference's `construct_optimizer`, lr policy and meters, next
to the same loops on the reference's own model with the same seeded state and the same random draws.  Per-iteration
losses (as logged by the reference's TrainMeter), updated parameters and the ensembled test predictions must agree.

CPU only, and only where /root/reference exists (the build container): the C-ABI ops are replaced by their torch
restatements (tests/shadow_ops.py), so this pins the host logic of the boundary -- call signatures, meta handling,
parameter naming / grouping, train / eval switching, return types -- not the kernels (tests/test_model_gpu.py does).
The drivers call `.cuda()` unconditionally (train_net.py:105-120); for this CPU run `Tensor.cuda` is the identity."""
import contextlib
import os
import sys

import pytest
import torch

import shadow_ops
import timesformer_oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the unmodified reference is only present in the build container")

torch.set_num_threads(max(1, os.cpu_count() or 1))
T, CROP, DEPTH, BV = 2, 32, 1, 2


@pytest.fixture
def env(monkeypatch, gold_dir):
    from procedurevrl_b200 import ops as real_ops
    from procedurevrl_b200.lib.models.vit import VisionTransformer
    for n in shadow_ops.ALL:
        monkeypatch.setattr(real_ops, n, getattr(shadow_ops, n))
    monkeypatch.setattr(VisionTransformer, "_require_cuda", False)
    monkeypatch.setenv("PVRL_PRECISION", "bf16x3")
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_driver
    return ref_driver


def _pretrain_cfg(ref_driver, gold_dir, accumulate=False):
    bank = os.path.join(gold_dir, "clip_step_emb_coin.pt")
    return ref_driver.reference_cfg("configs/HowTo100M/procedurevrl_adamw.yaml", [
        "NUM_GPUS", 1, "NUM_SHARDS", 1, "MODEL.PRETRAINED", False, "MODEL.DROP_PATH", 0.0, "TIMESFORMER.DEPTH", DEPTH,
        "DATA.NUM_FRAMES", T, "DATA.TRAIN_CROP_SIZE", CROP, "DATA.TEST_CROP_SIZE", CROP, "TRAIN.LABEL_EMB", bank,
        "MODEL.NUM_CLASSES", 778, "TRAIN.TEXT", "preextracted", "TRAIN.BATCH_SIZE", BV,
        "GLOBAL_BATCH_SIZE", 2 * BV if accumulate else BV, "LOG_PERIOD", 1, "SOLVER.BASE_LR", 1e-3, "SOLVER.WARMUP_EPOCHS", 0.0])


def _state(model, seed):
    """Seeded values for every parameter / buffer of `model` (reference and mirror share the state_dict schema)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in model.state_dict().items():
        if k.startswith("model.text_model."):
            sd[k] = v
        elif v.dtype.is_floating_point:
            sd[k] = (1.0 if k.endswith("norm.weight") or ".ln_" in k and k.endswith("weight") or "norm1.weight" in k
                     or "norm2.weight" in k else 0.0) + 0.05 * torch.randn(v.shape, generator=g)
        else:
            sd[k] = v
    return sd


class Loader:
    """A train / test loader as the drivers see it: len(), iteration over (inputs, labels, index, meta).  Before handing out
    batch i it runs `before(i)` (the mirror gets the reference's recorded random draws of that iteration)."""

    def __init__(self, batches, before=None):
        self.batches, self.before = batches, before
        self.dataset = type("D", (), {"_path_to_videos": [f"v{i}" for i in range(64)]})()

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        for i, b in enumerate(self.batches):
            if self.before is not None:
                self.before(i)
            inputs, labels, idx, meta = b
            yield inputs.clone(), labels.clone(), idx.clone(), {k: v.clone() for k, v in meta.items()}


def _batches(n):
    g = torch.Generator().manual_seed(99)
    out = []
    for i in range(n):
        frames = O.synthetic_clips(BV, 9, 3, T, CROP, CROP, seed=100 + i)
        meta = {"clip_text_emb": 0.4 * torch.randn(BV, 9, 512, generator=g), "clip_vis_feat": 0.4 * torch.randn(BV, 9, 512, generator=g)}
        out.append((frames, torch.ones(BV, 1, dtype=torch.long), torch.arange(BV) + BV * i, meta))
    return out


@contextlib.contextmanager
def _record(log):
    names = ["randint", "randn_like", "randperm"]
    orig = {n: getattr(torch, n) for n in names}
    for n in names:
        setattr(torch, n, (lambda n: lambda *a, **k: log.append((n, (lambda o: o)(orig[n](*a, **k)))) or log[-1][1])(n))
    try:
        yield
    finally:
        for n in names:
            setattr(torch, n, orig[n])


def _draws_per_iteration(log, n_iter):
    """Split the recorded RNG calls of `n_iter` reference forwards into per-iteration (mask_inds, pad_start, noise, rand_inds)."""
    out, i = [], 0
    per = []
    for name, t in log:
        per.append((name, t))
        if name == "randperm":
            out.append(per)
            per = []
    assert len(out) == n_iter
    res = []
    for per in out:
        ints = [t for n, t in per if n == "randint"]
        mask = ints[0]
        pad = torch.full((BV,), 9, dtype=torch.long)
        j = 1
        for b in range(BV):
            if int(mask[b]) + 1 != 9:
                pad[b] = int(ints[j])
                j += 1
        noise = torch.stack([t for n, t in per if n == "randn_like"])
        res.append((mask, pad, noise, [t for n, t in per if n == "randperm"][0]))
    return res


@pytest.mark.parametrize("accumulate", [False, True])
def test_train_epoch_unchanged_driver(env, gold_dir, accumulate):
    ref_driver = env
    import ref_shims
    _, ref_build, _ = ref_shims.load_reference()
    tn = ref_driver.load_train_net()
    optim = sys.modules["lib.models.optimizer"]
    meters = sys.modules["lib.utils.meters"]
    cfg = _pretrain_cfg(ref_driver, gold_dir, accumulate)
    n_iter = 2
    batches = _batches(n_iter)

    # ---- the reference's model under the reference's driver: records losses and random draws
    torch.manual_seed(0)
    cfg_cpu = _pretrain_cfg(ref_driver, gold_dir, accumulate)
    cfg_cpu.NUM_GPUS = 0                                   # build_model would call .cuda(device) for NUM_GPUS > 0
    ref_model = ref_build.build_model(cfg_cpu)
    state = _state(ref_model, seed=5)
    ref_model.load_state_dict(state, strict=True)
    ref_model.model.label_emb = ref_model.model.label_emb / ref_model.model.label_emb.norm(dim=1, keepdim=True)  # GPU semantics

    def ref_before(i):                                     # the frozen tower's output = the pre-extracted embedding of batch i
        ref_model.model.text_model.encode_text = lambda ids, i=i: batches[i][3]["clip_text_emb"].reshape(-1, 512)
    ref_batches = [(f, l, x, {"clip_text_ids": torch.zeros(BV, 9, 77, dtype=torch.long), "clip_vis_feat": m["clip_vis_feat"]})
                   for f, l, x, m in batches]
    ref_opt = optim.construct_optimizer(ref_model, cfg)
    ref_meter = meters.TrainMeter(n_iter, cfg)
    log, ref_losses = [], []
    orig_update = ref_meter.update_stats
    ref_meter.update_stats = lambda t1, t5, loss, lr, mb: (ref_losses.append(loss), orig_update(t1, t5, loss, lr, mb))[1]
    with _record(log):
        tn.train_epoch(Loader(ref_batches, ref_before), ref_model, ref_opt, ref_meter, 0, cfg)
    draws = _draws_per_iteration(log, n_iter)

    # ---- the drop-in model under the SAME driver, optimizer constructor, lr policy and meter
    from procedurevrl_b200.lib.models import MODEL_REGISTRY
    mirror = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(cfg)
    own = mirror.state_dict()
    mirror.load_state_dict({k: v for k, v in state.items() if k in own}, strict=True)
    assert set(k for k in own if not k.startswith("model.text_model.")) == \
        set(k for k in state if not k.startswith("model.text_model."))

    def before(i):
        mask, pad, noise, perm = draws[i]
        mirror.model.order_tfm.fixed_draws = (mask, pad, noise)
        mirror.model.fixed_rand_inds = perm
    opt = optim.construct_optimizer(mirror, cfg)
    meter = meters.TrainMeter(n_iter, cfg)
    losses = []
    orig_update2 = meter.update_stats
    meter.update_stats = lambda t1, t5, loss, lr, mb: (losses.append(loss), orig_update2(t1, t5, loss, lr, mb))[1]
    tn.train_epoch(Loader(batches, before), mirror, opt, meter, 0, cfg)

    assert len(losses) == len(ref_losses) == n_iter
    print("[train_epoch] reference", ref_losses, "drop-in", losses)
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-3 * abs(b) + 1e-4
    # same groups -> same update: parameters after the epoch
    ref_params = dict(ref_model.named_parameters())
    moved = 0
    for k, p in mirror.named_parameters():
        r = ref_params[k]
        assert p.requires_grad == r.requires_grad, k
        bad = (p.detach() - r.detach()).abs() > 2e-4 + 1e-3 * r.detach().abs()
        if k.endswith("qkv.bias") or k.endswith("in_proj_bias"):
            # softmax is invariant to the key bias: its true gradient is ZERO, what arrives is rounding noise, and Adam
            # turns noise into +-lr steps -- that third of the vector is a random walk in both runs
            n = bad.numel() // 3
            bad = torch.cat((bad[:n], bad[2 * n:]))
        assert bad.float().mean().item() <= max(5e-3, 2.0 / bad.numel()), (k, bad.float().mean().item())
        assert (p.detach() - r.detach()).abs().max().item() <= 2 * n_iter * cfg.SOLVER.BASE_LR + 1e-6, k
        moved += int(not torch.equal(r.detach(), state[k]))
    assert moved > 20                                      # the optimizer really stepped
    assert mirror.training and not any(m.training for m in [mirror.model.text_model] if isinstance(m, torch.nn.Module) and list(m.children()))


def test_perform_test_unchanged_driver(env, gold_dir):
    """tools/test_net.py::perform_test on the zero-shot step-classification configuration (DEV.MATCH_LANG_EMB, eval-mode
    softmax, 3-view sum ensemble in the reference's TestMeter): same video-level predictions and top-k stats."""
    ref_driver = env
    import ref_shims
    _, ref_build, _ = ref_shims.load_reference()
    tt = ref_driver.load_test_net()
    meters = sys.modules["lib.utils.meters"]
    bank = os.path.join(gold_dir, "clip_step_emb_coin.pt")
    ov = ["NUM_GPUS", 0, "MODEL.PRETRAINED", False, "MODEL.DROP_PATH", 0.0, "TIMESFORMER.DEPTH", DEPTH, "DATA.NUM_FRAMES", T,
          "DATA.TRAIN_CROP_SIZE", CROP, "DATA.TEST_CROP_SIZE", CROP, "DEV.MATCH_LANG_EMB", True, "DEV.TEST_LANG_EMB", bank,
          "MODEL.NUM_CLASSES", 778, "TEST.NUM_ENSEMBLE_VIEWS", 3, "TEST.NUM_SPATIAL_CROPS", 1, "LOG_PERIOD", 1,
          "TEST.DATASET", "howto100m_develop", "TEST.SAVE_RESULTS_PATH", ""]
    cfg = ref_driver.reference_cfg("configs/COIN/step_classification.yaml", ov)
    n_videos, views = 4, 3
    g = torch.Generator().manual_seed(3)
    batches = []
    clip_ids = torch.randperm(n_videos * views, generator=g)
    labels_v = torch.randint(1, 778, (n_videos,), generator=g)
    for a in range(0, n_videos * views, 4):
        ids = clip_ids[a:a + 4]
        x = torch.stack([O.synthetic_clips(3, T, CROP, CROP, seed=500 + int(i)) for i in ids])
        batches.append((x, labels_v[ids // views], ids, {}))

    ref_model = ref_build.build_model(cfg)
    state = _state(ref_model, seed=8)
    ref_model.load_state_dict(state, strict=True)
    ref_model.model.label_emb = ref_model.model.label_emb / ref_model.model.label_emb.norm(dim=1, keepdim=True)
    from procedurevrl_b200.lib.models import MODEL_REGISTRY
    mirror = MODEL_REGISTRY.get("vit_base_patch16_224_develop")(cfg)
    mirror.load_state_dict({k: v for k, v in state.items() if k in mirror.state_dict()}, strict=True)

    out = {}
    cwd = os.getcwd()
    for name, model in (("ref", ref_model), ("mirror", mirror)):
        meter = meters.TestMeter(n_videos, views, 778, len(batches), False, "sum")
        os.chdir(os.environ.get("PYTEST_TMP", "/tmp"))       # perform_test drops vis_pred_zeroshot_step_cls.pth into the cwd
        try:
            tt.perform_test(Loader(batches), model, meter, cfg)
        finally:
            os.chdir(cwd)
        assert not model.training
        out[name] = (meter.video_preds.clone(), meter.video_labels.clone(), dict(meter.stats))
    torch.testing.assert_close(out["mirror"][0], out["ref"][0], rtol=5e-3, atol=1e-6)
    assert torch.equal(out["mirror"][1], out["ref"][1])
    assert out["mirror"][2] == out["ref"][2], (out["mirror"][2], out["ref"][2])
    assert torch.equal(out["mirror"][0].argmax(1), out["ref"][0].argmax(1))
