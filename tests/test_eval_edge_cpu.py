"""Evaluation edge: `procedurevrl_b200.lib.utils.meters.TestMeter` (multi-view ensemble, vectorised and
device-resident) and `lib.utils.distributed.all_gather` against the UNMODIFIED reference meter (tests/golden/test_meter.pt,
written by oracle/make_golden_meter.py): accumulators, labels, clip counts and the final top-k strings for the "sum" and
"max" ensembles; then the N > 1 path: two gloo ranks each score half of every batch, all-gather, and both reach the
reference's numbers (tools/test_net.py:105-121)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from procedurevrl_b200.lib.utils import distributed as du
from procedurevrl_b200.lib.utils.meters import TestMeter
from procedurevrl_b200.lib.utils.metrics import topks_correct


@pytest.mark.parametrize("method", ["sum", "max"])
def test_meter_matches_reference(gold_dir, method):
    g = torch.load(os.path.join(gold_dir, "test_meter.pt"))
    m = TestMeter(g["num_videos"], g["num_clips"], g["num_cls"], len(g["cuts"]) - 1, False, method)
    for a, b in zip(g["cuts"][:-1], g["cuts"][1:]):
        ids = g["clip_ids"][a:b]
        m.update_stats(g["preds"][ids], g["labels"][a:b], ids)
    ref = g[method]
    torch.testing.assert_close(m.video_preds, ref["video_preds"], rtol=1e-6, atol=1e-7)
    assert torch.equal(m.video_labels, ref["video_labels"]) and torch.equal(m.clip_count, ref["clip_count"])
    assert m.finalize_metrics(ks=(1, 5)) == ref["stats"]
    m.reset()
    assert m.video_preds.abs().sum() == 0 and m.clip_count.sum() == 0


def test_topks_and_unsupported_modes():
    preds = torch.tensor([[0.1, 0.7, 0.2], [0.5, 0.3, 0.2], [0.2, 0.3, 0.5]])
    assert [int(x) for x in topks_correct(preds, torch.tensor([1, 1, 0]), (1, 2))] == [1, 2]
    with pytest.raises(NotImplementedError):
        TestMeter(2, 2, 3, 1, multi_label=True)
    with pytest.raises(NotImplementedError):
        TestMeter(2, 2, 3, 1, ensemble_method="mean")
    assert du.all_gather([preds])[0] is preds                     # no process group: identity


def _worker(rank, world, port, gold, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.load(gold)
    m = TestMeter(g["num_videos"], g["num_clips"], g["num_cls"], 1, False, "sum")
    n = g["clip_ids"].numel()
    for a in range(0, n - n % 6, 6):                              # batches of 6 clips: 3 per rank, as a DistributedSampler deals them
        ids = g["clip_ids"][a + 3 * rank:a + 3 * rank + 3]
        preds, labels, idx = du.all_gather([g["preds"][ids], g["labels"][a + 3 * rank:a + 3 * rank + 3], ids])
        assert preds.shape[0] == 6
        m.update_stats(preds, labels, idx)
    tail = g["clip_ids"][n - n % 6:]                               # the ragged tail is scored by every rank (no gather)
    m.update_stats(g["preds"][tail], g["labels"][n - n % 6:], tail)
    if rank == 1:
        torch.save({"video_preds": m.video_preds, "stats": m.finalize_metrics()}, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_all_gather_ensemble(gold_dir, tmp_path):
    gold = os.path.join(gold_dir, "test_meter.pt")
    out_path = str(tmp_path / "meter.pt")
    mp.spawn(_worker, args=(2, 33500 + os.getpid() % 2000, gold, out_path), nprocs=2, join=True)
    got, ref = torch.load(out_path), torch.load(gold)["sum"]
    torch.testing.assert_close(got["video_preds"], ref["video_preds"], rtol=1e-6, atol=1e-7)
    assert got["stats"] == ref["stats"]
