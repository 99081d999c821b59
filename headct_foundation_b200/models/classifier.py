"""Downstream linear probe head.  Mirror of `src/models/classifier.py:7-33`.

768 -> 2 logits on a [B, 768] CLS matrix: negligible work, so this stays host PyTorch;
the attentive-pooling classifier (classifier.py:35-100) is outside every benchmark config and not provided.
"""
import torch
from torch import nn


class LinearClassifier(nn.Module):
    def __init__(self, dim: int, num_classes: int):
        super().__init__()
        self.bn = nn.BatchNorm1d(dim, affine=False, eps=1e-6)
        self.linear = nn.Linear(dim, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.linear(self.bn(x))
