"""Evaluation edge on the GPU: the meter's accumulators live on the device of the predictions (no per-iteration .cpu()
as in tools/test_net.py:115-118) and reach the reference's numbers (tests/golden/test_meter.pt); the eval-mode softmax
that feeds it is the pvrl_softmax_rows kernel (vit.py:355-356)."""
import os

import pytest
import torch

from procedurevrl_b200 import functional as PF
from procedurevrl_b200.lib.utils.meters import TestMeter

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", ["sum", "max"])
def test_meter_on_device(gold_dir, method):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    g = torch.load(os.path.join(gold_dir, "test_meter.pt"))
    logits = g["preds"].log().cuda()                      # softmax(log p) = p: run the eval softmax kernel on the way in
    m = TestMeter(g["num_videos"], g["num_clips"], g["num_cls"], len(g["cuts"]) - 1, False, method)
    for a, b in zip(g["cuts"][:-1], g["cuts"][1:]):
        ids = g["clip_ids"][a:b].cuda()
        m.update_stats(PF.softmax_rows(logits[ids]), g["labels"][a:b].cuda(), ids)
    assert m.video_preds.is_cuda
    ref = g[method]
    torch.testing.assert_close(m.video_preds.cpu(), ref["video_preds"], rtol=1e-5, atol=1e-6)
    assert torch.equal(m.clip_count.cpu(), ref["clip_count"]) and torch.equal(m.video_labels.cpu(), ref["video_labels"])
    assert m.finalize_metrics(ks=(1, 5)) == ref["stats"]
