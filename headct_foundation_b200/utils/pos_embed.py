"""3-D / 2-D sin-cos position tables (host side, init time only).

Mirror of `src/utils/pos_embed.py:10-85` (build_sincos_position_embedding).  Checkpoint-time
interpolation (`interpolate_pos_embed*`, pos_embed.py:102-217) is out of the hot-path scope
 and is provided as a plain torch helper for API completeness only.
"""
from __future__ import annotations

from itertools import repeat
from typing import Sequence, Union

import torch
import torch.nn as nn


def _ntuple(x, n):
    if isinstance(x, (tuple, list)):
        return tuple(x)
    return tuple(repeat(x, n))


def _axis_terms(coords: torch.Tensor, n_freq: int, temperature: float) -> torch.Tensor:
    omega = 1.0 / (temperature ** (torch.arange(n_freq, dtype=torch.float32) / n_freq))
    return coords.reshape(-1, 1) * omega.reshape(1, -1)


def build_sincos_position_embedding(grid_size: Union[int, Sequence[int]], embed_dim: int, spatial_dims: int = 3,
                                    temperature: float = 10000.0) -> nn.Parameter:
    """[1, prod(grid), embed_dim] table, frozen Parameter (same return type as the reference).

    3-D quirk kept on purpose (pos_embed.py:54-58, :68-77): the first two mesh axes are built from
    arange(w) and arange(h) respectively, and the table concatenates sin/cos of the SECOND axis first.
    """
    if spatial_dims == 2:
        h, w = _ntuple(grid_size, 2)
        if embed_dim % 4 != 0:
            raise AssertionError("Embed dimension must be divisible by 4 for 2D sin-cos position embedding")
        gh, gw = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                                indexing="ij")
        th, tw = _axis_terms(gh, embed_dim // 4, temperature), _axis_terms(gw, embed_dim // 4, temperature)
        table = torch.cat([th.sin(), th.cos(), tw.sin(), tw.cos()], dim=1)[None]
    elif spatial_dims == 3:
        h, w, d = _ntuple(grid_size, 3)
        if embed_dim % 6 != 0:
            raise AssertionError("Embed dimension must be divisible by 6 for 3D sin-cos position embedding")
        g0, g1, g2 = torch.meshgrid(torch.arange(w, dtype=torch.float32), torch.arange(h, dtype=torch.float32),
                                    torch.arange(d, dtype=torch.float32), indexing="ij")
        n = embed_dim // 6
        t0, t1, t2 = (_axis_terms(g, n, temperature) for g in (g0, g1, g2))
        table = torch.cat([t1.sin(), t1.cos(), t0.sin(), t0.cos(), t2.sin(), t2.cos()], dim=1)[None]
    else:
        raise NotImplementedError(f"Spatial Dimension Size {spatial_dims} Not Implemented!")
    p = nn.Parameter(table)
    p.requires_grad = False
    return p
