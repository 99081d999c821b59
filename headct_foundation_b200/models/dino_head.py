"""DINO projection head.  Drop-in for `src/models/dino_head.py:7-41`."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as HF


class WeightNormLinear(nn.Module):
    """`nn.utils.weight_norm(nn.Linear(in, out, bias=False))` parameter layout (legacy names weight_g, weight_v,
    registered in that order) with the normalisation fused in front of the prototype GEMM."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        lin = nn.Linear(in_features, out_features, bias=False)       # same default init / RNG use as the reference
        w = lin.weight.detach()
        self.weight_g = nn.Parameter(w.norm(2, dim=1, keepdim=True))
        self.weight_v = nn.Parameter(w.clone())

    @torch.compiler.disable          # opaque to torch.compile: the body enqueues C-ABI launches, nothing to trace
    def forward(self, x16):
        return HF.WeightNormLinearFn.apply(x16, self.weight_g, self.weight_v)


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
                 bottleneck_dim=256):
        super().__init__()
        if use_bn:
            raise NotImplementedError("USE_BN is False in the shipped DINO config; BatchNorm heads are not accelerated")
        nlayers = max(nlayers, 1)
        if nlayers == 1:
            self.mlp = nn.Linear(in_dim, bottleneck_dim)
        else:
            layers = [nn.Linear(in_dim, hidden_dim), nn.GELU()]
            for _ in range(nlayers - 2):
                layers += [nn.Linear(hidden_dim, hidden_dim), nn.GELU()]
            layers.append(nn.Linear(hidden_dim, bottleneck_dim))
            self.mlp = nn.Sequential(*layers)
        self.apply(self._init_weights)
        self.last_layer = WeightNormLinear(bottleneck_dim, out_dim)
        self.last_layer.weight_g.data.fill_(1)
        if norm_last_layer:
            self.last_layer.weight_g.requires_grad = False

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    @torch.compiler.disable          # opaque to torch.compile: the body enqueues C-ABI launches, nothing to trace
    def forward(self, x):
        with torch.autocast(device_type="cuda", enabled=False):
            linears = [self.mlp] if isinstance(self.mlp, nn.Linear) else [m for m in self.mlp if isinstance(m, nn.Linear)]
            h = x
            for i, lin in enumerate(linears):
                h = HF.LinearFn.apply(h, lin.weight, lin.bias, i + 1 < len(linears), False)
            h = HF.L2NormFn.apply(h)
            return self.last_layer(h)
