"""reference lib/utils/metrics.py:10-41: top-k hit counts (device-agnostic torch; no host synchronisation)."""
import torch


def topks_correct(preds, labels, ks):
    """preds [N, classes], labels [N] -> list of 0-d tensors: how many rows have their label among the top-k predictions."""
    assert preds.size(0) == labels.size(0), "Batch dim of predictions and labels must match"
    top = torch.topk(preds, max(ks), dim=1, largest=True, sorted=True)[1]
    hit = top.eq(labels.view(-1, 1))
    return [hit[:, :k].float().sum() for k in ks]


def topk_accuracies(preds, labels, ks):
    """metrics.py:60-72: percentages."""
    return [(x / preds.size(0)) * 100.0 for x in topks_correct(preds, labels, ks)]
