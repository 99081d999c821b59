"""Downstream metrics (engine_downstream.py:299-308) against scikit-learn's definitions."""
import numpy as np
import pytest
import torch

from headct_foundation_b200.utils.metrics import DownstreamMetrics, multiclass_accuracy, multiclass_auroc


@pytest.mark.parametrize("n,C,ties", [(200, 2, False), (500, 3, False), (300, 2, True), (64, 4, True)])
def test_metrics_match_sklearn(n, C, ties):
    from sklearn.metrics import recall_score, roc_auc_score
    rng = np.random.default_rng(n + C)
    target = rng.integers(0, C, n)
    logits = rng.standard_normal((n, C)) + 1.5 * np.eye(C)[target]
    if ties:
        logits = np.round(logits * 2) / 2                       # many exactly equal scores
    probs = torch.softmax(torch.from_numpy(logits), dim=1).float()
    t = torch.from_numpy(target)
    acc = multiclass_accuracy(probs, t, C).numpy()
    ref_acc = recall_score(target, probs.argmax(1).numpy(), labels=list(range(C)), average=None, zero_division=0)
    assert np.allclose(acc, ref_acc, atol=1e-6)
    auc = multiclass_auroc(probs, t, C).numpy()
    ref_auc = [roc_auc_score((target == c).astype(int), probs[:, c].numpy()) for c in range(C)]
    assert np.allclose(auc, ref_auc, atol=1e-6)


def test_collection_accumulates_and_handles_missing_classes():
    m = DownstreamMetrics(3).to("cpu")
    p1 = torch.tensor([[0.7, 0.2, 0.1], [0.1, 0.8, 0.1]])
    p2 = torch.tensor([[0.6, 0.3, 0.1], [0.2, 0.7, 0.1]])
    m(p1, torch.tensor([0, 1]))
    m.update(p2, torch.tensor([1, 1]))
    out = m.compute()
    assert torch.allclose(out["MulticlassAccuracy"], torch.tensor([1.0, 2 / 3, 0.0]))
    assert out["MulticlassAUROC"][2] == 0                       # class 2 never occurs: 0, as torchmetrics reports it
    assert abs(out["MulticlassAUROC"][0].item() - 1.0) < 1e-6
    m.reset()
    assert m.compute()["MulticlassAUROC"].abs().sum() == 0
