"""Downstream probe heads.  Mirror of `src/models/classifier.py`.

`LinearClassifier` (classifier.py:7-33) works on a [B, 768] CLS matrix: negligible work, so it stays host PyTorch on
the device.  `AttentionClassifier` (classifier.py:35-100) reads every token of every sample:
its BatchNorm over the token matrix, the wkv Linear and the attentive pooling run on the B200 kernels
(`hct_colnorm_*`, the tcgen05 GEMM, `hct_pool_attention_*`); the [B, 768] tail (bn2, mean over queries, 768 -> classes)
is host PyTorch like the linear probe.
"""
from typing import Optional

import torch
from torch import nn

from .. import functional as HF


class LinearClassifier(nn.Module):
    def __init__(self, dim: int, num_classes: int):
        super().__init__()
        self.bn = nn.BatchNorm1d(dim, affine=False, eps=1e-6)
        self.linear = nn.Linear(dim, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.linear(self.bn(x))


class AttentionClassifier(nn.Module):
    def __init__(self, dim: int, num_classes: int, num_heads: int = 12, qkv_bias: bool = False,
                 qk_scale: Optional[float] = None, num_queries: int = 1):
        super().__init__()
        self.num_heads = num_heads
        self.num_queries = num_queries
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.bn1 = nn.BatchNorm1d(dim, affine=False, eps=1e-6)
        self.bn2 = nn.BatchNorm1d(dim, affine=False, eps=1e-6)
        self.wkv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.linear = nn.Linear(dim, num_classes)
        self.cls_token = nn.Parameter(torch.zeros(1, num_queries, dim))
        nn.init.trunc_normal_(self.cls_token, std=.02)

    @torch.compiler.disable          # opaque to torch.compile: the body enqueues C-ABI launches, nothing to trace
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, N, C = x.shape
        with torch.autocast(device_type="cuda", enabled=False):
            bn = self.bn1
            use_batch_stats = self.training or bn.running_mean is None
            momentum = bn.momentum
            if self.training and bn.track_running_stats and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
                if momentum is None:
                    momentum = 1.0 / float(bn.num_batches_tracked)
            xh = HF.ColNormFn.apply(x.float(), bn.running_mean, bn.running_var, use_batch_stats, bn.eps,
                                    momentum if momentum is not None else 0.0, True)                      # :89
            kv = HF.LinearFn.apply(xh, self.wkv.weight, self.wkv.bias, False, False)                       # :90, bf16 [B,N,2C]
            # q is scaled by self.scale (:87) and SDPA applies 1/sqrt(head_dim) once more (:93)
            scale_total = self.scale * (C // self.num_heads) ** -0.5
            pooled = HF.PoolAttentionFn.apply(self.cls_token.view(self.num_queries, C), kv, self.num_heads, scale_total)
            if self.num_queries > 1:
                # :95 reshapes SDPA's [B, H, nq, hd] output to [B, nq, C] WITHOUT a transpose; mirror that mixing of
                # heads and queries (for one query the two layouts coincide)
                hd = C // self.num_heads
                pooled = pooled.view(B, self.num_queries, self.num_heads, hd).permute(0, 2, 1, 3).reshape(B, self.num_queries, C)
            x_cls = self.bn2(pooled.transpose(-2, -1)).transpose(-2, -1)                                   # :95-96
            return self.linear(x_cls.mean(dim=1))                                                          # :97-99
