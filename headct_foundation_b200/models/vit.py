"""ViT-B 3-D backbone (DINO / fine-tune / feature extraction).  Drop-in for `src/models/vit.py`."""
from __future__ import annotations

from typing import Sequence, Union

import torch
import torch.nn as nn

from .. import functional as HF
from ..utils.patch_embedding import PatchEmbeddingBlock
from .attentionblock import AttentionBlock

__all__ = ["ViT"]


class ViT(nn.Module):
    def __init__(self, in_chans: int, img_size: Union[Sequence[int], int], patch_size: Union[Sequence[int], int],
                 hidden_size: int = 768, mlp_dim: int = 3072, num_layers: int = 12, num_heads: int = 12,
                 patch_embed: str = "conv", pos_embed: str = "learnable", classification: bool = False,
                 num_classes: int = 2, dropout_rate: float = 0.0, spatial_dims: int = 3,
                 num_register_tokens: int = 0, post_activation: str = "Tanh", qkv_bias: bool = False,
                 lora: bool = False, norm_layer=nn.LayerNorm) -> None:
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        self.classification = classification
        self.patch_embedding = PatchEmbeddingBlock(img_size=img_size, patch_size=patch_size, in_channels=in_chans,
                                                   hidden_size=hidden_size, num_heads=num_heads,
                                                   patch_embed=patch_embed, pos_embed=pos_embed,
                                                   dropout_rate=dropout_rate, spatial_dims=spatial_dims)
        self.blocks = nn.ModuleList([
            AttentionBlock(hidden_size, mlp_dim, num_heads, dropout_rate, qkv_bias=qkv_bias, save_attn=False,
                           lora=lora, norm_layer=norm_layer) for _ in range(num_layers)])
        self.cls_token = nn.Parameter(torch.zeros(1, 1, hidden_size))
        self.norm = norm_layer(hidden_size, eps=1e-6)
        self.num_register_tokens = num_register_tokens
        assert num_register_tokens >= 0
        self.register_tokens = (nn.Parameter(torch.zeros(1, num_register_tokens, hidden_size))
                                if num_register_tokens else None)
        if self.classification:
            if post_activation == "Tanh":
                self.classification_head = nn.Sequential(nn.Linear(hidden_size, num_classes), nn.Tanh())
            else:
                self.classification_head = nn.Linear(hidden_size, num_classes)

    def init_weights(self):
        nn.init.normal_(self.cls_token, std=1e-6)
        if self.register_tokens is not None:
            nn.init.normal_(self.register_tokens, std=1e-6)

    def forward(self, x):
        # traceable: under torch.compile (main_downstream.py:162) the kernels are reached through torch.library ops
        # (ops.py: headct::embed, headct::block, headct::layernorm) -- no graph breaks
        with torch.autocast(device_type="cuda", enabled=False):
            prefix = self.cls_token
            if self.register_tokens is not None:
                prefix = torch.cat((self.cls_token, self.register_tokens), dim=1)   # cls, then registers (vit.py:152-160)
            x = self.patch_embedding.embed(x, prefix=prefix)
            hidden_states_out = []
            residual = None
            for blk in self.blocks:
                x, residual = blk(x, residual)
                hidden_states_out.append(x)
            x = HF.layernorm(x, self.norm.weight, getattr(self.norm, "bias", None), self.norm.eps, False)
            if hasattr(self, "classification_head"):
                x = self.classification_head(x[:, 0])       # tiny host-torch head (CLASSIFICATION: False in all configs)
        return x, hidden_states_out
