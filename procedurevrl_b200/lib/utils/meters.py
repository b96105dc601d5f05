"""Evaluation edge: the multi-view ensemble of reference lib/utils/meters.py:20-200 (`TestMeter`), kept on
the device the predictions live on.

The reference moves every batch of softmax probabilities to the host (tools/test_net.py:115-118: three `.cpu()` round
trips per iteration) and adds them to the per-video accumulators one clip at a time in Python.  Here `update_stats` is one
`index_add_` (sum) / `scatter_reduce_` (max) on the accumulators, which stay wherever `preds` is (the GPU in test_net's
loop: nothing synchronises until `finalize_metrics`).  Same constructor, attributes (`video_preds`, `video_labels`,
`clip_count`, `stats`, `video_names`) and final numbers as the reference (tests/golden/test_meter.pt)."""
import time

import torch

from . import metrics


class _Timer:
    """The three calls of fvcore.common.timer.Timer the meter uses."""

    def __init__(self):
        self.reset()

    def reset(self):
        self._start, self._paused = time.perf_counter(), None

    def pause(self):
        self._paused = time.perf_counter()

    def seconds(self):
        return (self._paused if self._paused is not None else time.perf_counter()) - self._start


class TestMeter:
    __test__ = False        # not a pytest class

    def __init__(self, num_videos, num_clips, num_cls, overall_iters, multi_label=False, ensemble_method="sum", device="cpu"):
        if multi_label:
            raise NotImplementedError("multi_label (mAP over AVA / Charades) is not on the ProcedureVRL evaluation path")
        if ensemble_method not in ("sum", "max"):
            raise NotImplementedError("Ensemble Method {} is not supported".format(ensemble_method))
        self.iter_timer, self.data_timer, self.net_timer = _Timer(), _Timer(), _Timer()
        self.num_clips, self.overall_iters = num_clips, overall_iters
        self.multi_label, self.ensemble_method = multi_label, ensemble_method
        self.video_preds = torch.zeros((num_videos, num_cls), device=device)
        self.video_labels = torch.zeros((num_videos,), dtype=torch.long, device=device)
        self.clip_count = torch.zeros((num_videos,), dtype=torch.long, device=device)
        self.topk_accs, self.stats = [], {}
        self.video_names = [""] * num_videos

    def _to(self, device):
        if self.video_preds.device != device:
            self.video_preds, self.video_labels = self.video_preds.to(device), self.video_labels.to(device)
            self.clip_count = self.clip_count.to(device)

    def reset(self):
        self.clip_count.zero_()
        self.video_preds.zero_()
        self.video_labels.zero_()

    def update_stats(self, preds, labels, clip_ids, video_name_list=None):
        """preds [N, classes] (eval-mode probabilities), labels [N], clip_ids [N] -> accumulate into video clip_id // num_clips
        (meters.py:83-123).  The accumulators follow `preds` to its device on first use."""
        self._to(preds.device)
        vid = torch.div(clip_ids.to(preds.device).long(), self.num_clips, rounding_mode="floor")
        self.video_labels[vid] = labels.to(preds.device).long()
        if self.ensemble_method == "sum":
            self.video_preds.index_add_(0, vid, preds.to(self.video_preds.dtype))
        else:
            self.video_preds.scatter_reduce_(0, vid.view(-1, 1).expand_as(preds), preds.to(self.video_preds.dtype), reduce="amax",
                                             include_self=True)
        self.clip_count.index_add_(0, vid, torch.ones_like(vid))
        if video_name_list is not None:                           # visualisation aid of the reference; host-side by nature
            for c in clip_ids.tolist():
                if self.video_names[int(c) // self.num_clips] == "":
                    self.video_names[int(c) // self.num_clips] = video_name_list[int(c)].split("/")[-1]

    def iter_tic(self):
        self.iter_timer.reset()
        self.data_timer.reset()

    def iter_toc(self):
        self.iter_timer.pause()
        self.net_timer.pause()

    def data_toc(self):
        self.data_timer.pause()
        self.net_timer.reset()

    def log_iter_stats(self, cur_iter):
        return {"split": "test_iter", "cur_iter": "{}".format(cur_iter + 1), "time_diff": self.iter_timer.seconds()}

    def finalize_metrics(self, ks=(1, 5)):
        """meters.py:163-200: top-k accuracies of the ensembled predictions (the one host synchronisation of the meter)."""
        self.stats = {"split": "test_final"}
        topks = metrics.topk_accuracies(self.video_preds, self.video_labels, ks)
        for k, topk in zip(ks, topks):
            self.stats["top{}_acc".format(k)] = "{:.{prec}f}".format(float(topk), prec=2)
        return self.stats
