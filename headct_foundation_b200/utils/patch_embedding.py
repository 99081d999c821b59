"""PatchEmbeddingBlock: Conv3d(k = s = patch) + position embedding, as a TMA/tcgen05 GEMM.

Drop-in for `src/utils/patch_embedding.py:32-161`.  The Conv3d module only HOLDS the parameters
(same names, shapes and default initialisation as the reference); the forward pass is
patchify (im2col in the conv-weight K order) -> GEMM with bias + position-embedding epilogue.
"""
from __future__ import annotations

from typing import Optional, Sequence, Union

import numpy as np
import torch
import torch.nn as nn

from .. import functional as HF
from .pos_embed import build_sincos_position_embedding

SUPPORTED_EMBEDDING_TYPES = {"conv", "perceptron"}


def _tuple_rep(v, n):
    if isinstance(v, (tuple, list)):
        if len(v) != n:
            raise ValueError(f"Sequence must have length {n}, got {len(v)}.")
        return tuple(int(i) for i in v)
    return (int(v),) * n


class PatchEmbeddingBlock(nn.Module):
    def __init__(self, in_channels: int, img_size: Union[Sequence[int], int], patch_size: Union[Sequence[int], int],
                 hidden_size: int, num_heads: int, patch_embed: str = "conv", pos_embed: str = "learnable",
                 dropout_rate: float = 0.0, spatial_dims: int = 3) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError(f"dropout_rate {dropout_rate} should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError(f"hidden size {hidden_size} should be divisible by num_heads {num_heads}.")
        if patch_embed not in SUPPORTED_EMBEDDING_TYPES:
            raise ValueError(f"Unsupported option '{patch_embed}', Available options are {SUPPORTED_EMBEDDING_TYPES}.")
        self.patch_embed = patch_embed
        img_size = _tuple_rep(img_size, spatial_dims)
        patch_size = _tuple_rep(patch_size, spatial_dims)
        self.img_size, self.patch_size, self.spatial_dims = img_size, patch_size, spatial_dims
        for m, p in zip(img_size, patch_size):
            if m < p:
                raise ValueError("patch_size should be smaller than img_size.")
            if patch_embed == "perceptron" and m % p != 0:
                raise ValueError("patch_size should be divisible by img_size for perceptron.")
        self.n_patches = np.prod([im // p for im, p in zip(img_size, patch_size)])
        self.patch_dim = int(in_channels * np.prod(patch_size))
        grid = []
        for im, p in zip(img_size, patch_size):
            assert im % p == 0, "input size and patch size are not proper"
            grid.append(im // p)
        if patch_embed != "conv":
            raise ValueError(f"patch_embed type {patch_embed} not supported.")
        conv = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[spatial_dims]
        self.patch_embeddings = conv(in_channels=in_channels, out_channels=hidden_size, kernel_size=patch_size,
                                     stride=patch_size)
        self.position_embeddings: Optional[nn.Parameter] = nn.Parameter(torch.zeros(1, int(self.n_patches), hidden_size))
        self.dropout = nn.Dropout(dropout_rate)
        self.dropout_rate = dropout_rate
        if pos_embed == "none":
            self.position_embeddings = None
        elif pos_embed == "learnable":
            nn.init.trunc_normal_(self.position_embeddings, mean=0.0, std=0.02, a=-2.0, b=2.0)
        elif pos_embed == "sincos":
            with torch.no_grad():
                self.position_embeddings.data.copy_(
                    build_sincos_position_embedding(grid, hidden_size, spatial_dims).float())
        else:
            raise ValueError(f"pos_embed type {pos_embed} not supported.")
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, mean=0.0, std=0.02, a=-2.0, b=2.0)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _check(self, x: torch.Tensor) -> None:
        if self.spatial_dims != 3 or len(set(self.patch_size)) != 1:
            raise NotImplementedError("the B200 patch-embed path supports cubic 3-D patches only")
        if tuple(x.shape[2:]) != self.img_size:
            # the reference interpolates the table on the fly (patch_embedding.py:137-144); never triggered by
            # the shipped configs (all inputs are 96^3) and out of the hot-path scope.
            raise NotImplementedError(f"input size {tuple(x.shape[2:])} != constructor size {self.img_size}")
        if self.dropout_rate > 0 and self.training:
            raise NotImplementedError("dropout_rate > 0 is not supported by the fused path (all shipped configs use 0.)")

    def embed(self, x: torch.Tensor, prefix: Optional[torch.Tensor] = None,
              ids_keep: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B, P + n, hidden] fp32: optional prefix tokens (cls / registers) then the (selected) patch tokens."""
        self._check(x)
        pe = self.patch_embeddings
        return HF.embed(x, pe.weight, pe.bias, self.position_embeddings, prefix, ids_keep, self.patch_size[0])

    def forward(self, x: torch.Tensor) -> torch.Tensor:      # traceable (headct::embed under torch.compile)
        return self.embed(x)
