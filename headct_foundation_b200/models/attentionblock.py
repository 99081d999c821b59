"""Pre-LN transformer block on B200 kernels.  Drop-in for `src/models/attentionblock.py`.

nn.Linear / nn.LayerNorm modules hold the parameters (identical names, order and default init as
the reference + monai MLPBlock); compute goes through `functional.BlockFn` -- fused LayerNorm ->
tcgen05 GEMMs with bias / GELU / residual epilogues -> flash attention.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as HF
from .layers import RMSNorm


class LoraLinear(nn.Module):
    """LoRA adapter (attentionblock.py:6-22): y = x (B A)^T, evaluated low-rank as (x A^T) B^T on the tcgen05 GEMM.
    Inside an `AttentionBlock` the adapters are folded into `BlockFn`; this forward serves stand-alone use."""

    def __init__(self, in_features: int, out_features: int, r: int = 8):
        super().__init__()
        self.lora_matrix_B = nn.Parameter(torch.zeros(out_features, r))
        self.lora_matrix_A = nn.Parameter(torch.randn(r, in_features))

    @torch.compiler.disable          # opaque to torch.compile: the body enqueues C-ABI launches, nothing to trace
    def forward(self, x):
        t = HF.LinearFn.apply(x, self.lora_matrix_A, None, False, False)
        y = HF.LinearFn.apply(t, self.lora_matrix_B, None, False, False)
        return y.float() if x.dtype == torch.float32 else y


class MLPBlock(nn.Module):
    """linear2(GELU_erf(linear1(x))) -- monai.networks.blocks.mlp.MLPBlock as used at attentionblock.py:91."""

    def __init__(self, hidden_size: int, mlp_dim: int, dropout_rate: float = 0.0):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        mlp_dim = mlp_dim or hidden_size
        self.linear1 = nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.fn = nn.GELU()
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = self.drop1

    @torch.compiler.disable          # opaque to torch.compile: the body enqueues C-ABI launches, nothing to trace
    def forward(self, x):
        h = HF.LinearFn.apply(x, self.linear1.weight, self.linear1.bias, True, False)
        y = HF.LinearFn.apply(h, self.linear2.weight, self.linear2.bias, False, False)
        return y.float() if x.dtype == torch.float32 else y


class SelfAttention(nn.Module):
    def __init__(self, hidden_size: int, num_heads: int = 12, dropout: float = 0.0, qkv_proj_bias: bool = False,
                 lora: bool = False):
        super().__init__()
        self.num_heads = num_heads
        self.hidden_size = hidden_size
        self.lora = lora
        self.qkv = nn.Linear(hidden_size, hidden_size * 3, bias=qkv_proj_bias)
        self.proj = nn.Linear(hidden_size, hidden_size)
        self.proj_drop = nn.Dropout(dropout)
        if self.lora:
            self.lora_q = LoraLinear(hidden_size, hidden_size, r=128)
            self.lora_v = LoraLinear(hidden_size, hidden_size, r=128)
        self.dropout = dropout

    @torch.compiler.disable          # opaque to torch.compile: the body enqueues C-ABI launches, nothing to trace
    def forward(self, x, attn_mask=None):
        if attn_mask is not None or (self.dropout > 0 and self.training):
            raise NotImplementedError("attn_mask / dropout are outside the accelerated path "
                                      "(unused by every shipped config)")
        B, N, C = x.shape
        qkv = HF.LinearFn.apply(x, self.qkv.weight, self.qkv.bias, False, False)          # bf16 [B,N,3C]
        if self.lora:
            lq = HF.LinearFn.apply(HF.LinearFn.apply(x, self.lora_q.lora_matrix_A, None, False, False),
                                   self.lora_q.lora_matrix_B, None, False, False)
            lv = HF.LinearFn.apply(HF.LinearFn.apply(x, self.lora_v.lora_matrix_A, None, False, False),
                                   self.lora_v.lora_matrix_B, None, False, False)
            qkv = HF.LoraAddFn.apply(qkv, lq, lv, self.num_heads)                         # the reshape quirk, :57-59
        y = HF.AttentionFn.apply(qkv, self.num_heads)                                     # bf16 [B,N,C]
        y = HF.LinearFn.apply(y, self.proj.weight, self.proj.bias, False, False)
        return y.float() if x.dtype == torch.float32 else y


class AttentionBlock(nn.Module):
    def __init__(self, hidden_size: int, mlp_dim: int, num_heads: int, dropout_rate: float = 0.0,
                 qkv_bias: bool = False, save_attn: bool = False, lora: bool = False, norm_layer=nn.LayerNorm):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        if norm_layer not in (nn.LayerNorm, RMSNorm):
            raise NotImplementedError("norm_layer must be nn.LayerNorm or RMSNorm (config NORM_LAYER: layernorm | rmsnorm)")
        self.dropout_rate = dropout_rate
        self.num_heads = num_heads
        self.mlp = MLPBlock(hidden_size, mlp_dim, dropout_rate)
        self.att_norm = norm_layer(hidden_size)
        self.ffn_norm = norm_layer(hidden_size)
        self.attn = SelfAttention(hidden_size, num_heads, dropout=dropout_rate, qkv_proj_bias=qkv_bias, lora=lora)

    def forward(self, hidden_states, residual=None):
        # traceable: under torch.compile the body runs as the torch.library op headct::block (ops.py)
        if self.dropout_rate > 0 and self.training:
            raise NotImplementedError("dropout is outside the accelerated path")
        if self.att_norm.eps != self.ffn_norm.eps:
            raise NotImplementedError("att_norm and ffn_norm must share eps")
        a, m = self.attn, self.mlp
        lora = ((a.lora_q.lora_matrix_A, a.lora_q.lora_matrix_B, a.lora_v.lora_matrix_A, a.lora_v.lora_matrix_B)
                if a.lora else (None, None, None, None))
        out = HF.block(hidden_states, self.att_norm.weight, getattr(self.att_norm, "bias", None),
                       a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias,
                       self.ffn_norm.weight, getattr(self.ffn_norm, "bias", None),
                       m.linear1.weight, m.linear1.bias, m.linear2.weight, m.linear2.bias, lora,
                       self.num_heads, self.att_norm.eps)
        return out, residual
