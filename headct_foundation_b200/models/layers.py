"""RMSNorm on the B200 row kernels.  Drop-in for `src/models/layers.py:11-53` (the NORM_LAYER: 'rmsnorm' option,
main_downstream.py:111-116, main_pretrain_mae.py:102): `x * rsqrt(mean(x^2) + eps) * weight`, one `weight` parameter."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as HF


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    @torch.compiler.disable          # opaque to torch.compile: the body enqueues C-ABI launches, nothing to trace
    def forward(self, x):
        y = HF.LayerNormFn.apply(x.float(), self.weight, None, self.eps, False)      # layers.py:52: normalise in fp32 ...
        return y.type_as(x)                                                          # ... and hand back x's dtype
