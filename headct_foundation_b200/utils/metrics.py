"""Epoch-level downstream metrics as plain torch ops on whatever device the predictions live on.

The reference collects them with torchmetrics (`MetricCollection([MulticlassAccuracy(average=None),
MulticlassAUROC(average=None)])`, engine_downstream.py:299-308; fed once per epoch with all softmax outputs and targets,
:138 / :246, then `.compute()` / `.reset()`, :337-346).  torchmetrics is a third-party dependency that is not part of
this image; the two quantities are restated here from their definitions: per-class accuracy = recall of each class,
per-class AUROC = one-vs-rest area under the exact ROC curve (Mann-Whitney statistic with average ranks for ties,
which equals the trapezoidal area torchmetrics / scikit-learn compute).  Classes without positives or without negatives
score 0, as torchmetrics does.  Not on the hot path: one sort of n scores per class and epoch.
"""
from __future__ import annotations

from typing import Dict

import torch


def multiclass_accuracy(preds: torch.Tensor, target: torch.Tensor, num_classes: int) -> torch.Tensor:
    """preds [n, C] (probabilities or logits) or [n] (labels); target [n] -> per-class recall [C]."""
    labels = preds.argmax(dim=1) if preds.dim() == 2 else preds
    target = target.long()
    hit = (labels == target).to(torch.float64)
    tp = torch.zeros(num_classes, dtype=torch.float64, device=target.device).scatter_add_(0, target, hit)
    cnt = torch.zeros(num_classes, dtype=torch.float64, device=target.device).scatter_add_(0, target, torch.ones_like(hit))
    return torch.where(cnt > 0, tp / cnt.clamp_min(1), torch.zeros_like(tp)).to(torch.float32)


def _average_ranks(scores: torch.Tensor) -> torch.Tensor:
    """1-based ranks of `scores` in ascending order, tied values sharing the mean of their positions."""
    order = torch.argsort(scores, stable=True)
    s = scores[order]
    _, inverse, counts = torch.unique_consecutive(s, return_inverse=True, return_counts=True)
    ends = torch.cumsum(counts, 0).to(torch.float64)                # last 1-based position of each tie group
    mean_rank = ends - (counts.to(torch.float64) - 1) / 2
    ranks = torch.empty_like(mean_rank[inverse])
    ranks[order] = mean_rank[inverse]
    return ranks


def multiclass_auroc(probs: torch.Tensor, target: torch.Tensor, num_classes: int) -> torch.Tensor:
    """probs [n, C]; target [n] -> one-vs-rest AUROC per class [C]."""
    target = target.long()
    out = torch.zeros(num_classes, dtype=torch.float32, device=probs.device)
    n = target.numel()
    for c in range(num_classes):
        pos = target == c
        n_pos = int(pos.sum())
        n_neg = n - n_pos
        if n_pos == 0 or n_neg == 0:
            continue
        ranks = _average_ranks(probs[:, c].to(torch.float64))
        u = ranks[pos].sum() - n_pos * (n_pos + 1) / 2
        out[c] = (u / (n_pos * n_neg)).to(torch.float32)
    return out


class DownstreamMetrics:
    """Stand-in for the reference's MetricCollection: call / update with (softmax outputs, targets), then `compute()`
    returns {"MulticlassAccuracy": [C], "MulticlassAUROC": [C]} over everything seen since `reset()`."""

    def __init__(self, num_classes: int):
        self.num_classes = num_classes
        self.reset()

    def reset(self) -> None:
        self._probs, self._targets = [], []

    def to(self, device):          # the reference moves its collection to the device; state follows the inputs here
        return self

    def update(self, probs: torch.Tensor, target: torch.Tensor) -> None:
        self._probs.append(probs.detach())
        self._targets.append(target.detach().long())

    __call__ = update

    def compute(self) -> Dict[str, torch.Tensor]:
        if not self._probs:
            z = torch.zeros(self.num_classes)
            return {"MulticlassAccuracy": z, "MulticlassAUROC": z.clone()}
        probs, target = torch.cat(self._probs), torch.cat(self._targets)
        return {"MulticlassAccuracy": multiclass_accuracy(probs, target, self.num_classes),
                "MulticlassAUROC": multiclass_auroc(probs, target, self.num_classes)}
