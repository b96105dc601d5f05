"""Golden vectors for the evaluation edge (multi-view sum-ensemble, reference lib/utils/meters.py:20-200 +
lib/utils/metrics.py:10-41), produced by the UNMODIFIED reference `TestMeter`.

TEST INFRASTRUCTURE ONLY (runs where /root/reference exists).  `lib.utils.meters` pulls in the data pipeline through
`lib.utils.misc` (ffmpeg, av, ...), which the meter itself never touches: `lib.utils.misc` and `lib.utils.logging` are
stubbed in sys.modules, the meter / metrics files are the reference's own.  Writes tests/golden/test_meter.pt: clip-level
predictions in shuffled order, cut into uneven batches, and the meter's state + final stats for the "sum" and "max" ensembles.

    python oracle/make_golden_meter.py
"""
import importlib
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

GOLD = os.path.join(HERE, "..", "tests", "golden")


def load_meters():
    ref_shims.install()
    if ref_shims.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, ref_shims.REFERENCE_ROOT)
    log = types.ModuleType("lib.utils.logging")
    log.get_logger = lambda name: __import__("logging").getLogger(name)
    log.log_json_stats = lambda stats: None
    sys.modules["lib.utils.logging"] = log
    sys.modules["lib.utils.misc"] = types.ModuleType("lib.utils.misc")
    for pkg in ("lib", "lib.utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(ref_shims.REFERENCE_ROOT, *pkg.split("."))]
            sys.modules[pkg] = m
    return importlib.import_module("lib.utils.meters")


def main():
    meters = load_meters()
    g = torch.Generator().manual_seed(2024)
    num_videos, num_clips, num_cls = 37, 3, 50
    n = num_videos * num_clips
    labels_v = torch.randint(1, num_cls, (num_videos,), generator=g)          # > 0: the reference's consistency assert fires on sum() > 0
    logits = torch.randn(n, num_cls, generator=g)
    logits[torch.arange(n), labels_v.repeat_interleave(num_clips)] += 1.5      # a learnable signal: top-1 well above chance
    preds = logits.softmax(1)                                                # what the model returns in eval mode (vit.py:355-356)
    order = torch.randperm(n, generator=g)
    cuts = [0, 5, 6, 30, 31, 64, 100, n]
    out = {"num_videos": num_videos, "num_clips": num_clips, "num_cls": num_cls, "preds": preds, "clip_ids": order,
           "labels": labels_v.repeat_interleave(num_clips)[order], "cuts": cuts}
    for method in ("sum", "max"):
        m = meters.TestMeter(num_videos, num_clips, num_cls, len(cuts) - 1, False, method)
        for a, b in zip(cuts[:-1], cuts[1:]):
            ids = order[a:b]
            m.update_stats(preds[ids], out["labels"][a:b], ids)
        m.finalize_metrics(ks=(1, 5))
        out[method] = {"video_preds": m.video_preds.clone(), "video_labels": m.video_labels.clone(),
                       "clip_count": m.clip_count.clone(), "stats": dict(m.stats)}
        print(method, m.stats)
    torch.save(out, os.path.join(GOLD, "test_meter.pt"))


if __name__ == "__main__":
    main()
