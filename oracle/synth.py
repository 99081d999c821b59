"""Deterministic synthetic weights / inputs shared by the golden generator, the tests and bench.

TEST INFRASTRUCTURE.  numpy PCG64 streams (bit-stable across machines and torch versions), so the
GPU box can rebuild exactly the tensors the golden vectors were produced from without shipping
600 MB of weights.  Key names / shapes / order follow SURVEY.md Appendix A (the reference
`state_dict()` layout); `oracle/gen_golden.py` asserts that against the real reference modules.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Sequence

import numpy as np
import torch

from . import headct_oracle as O

# yaml values: configs/mae/mae_HeadCT.yaml:31-50, configs/dino/dino_HeadCT.yaml:33-68,
# configs/downstream/vit_HeadCT_cq500.yaml:34-51
MAE_FULL = dict(input_size=96, patch_size=12, mask_ratio=0.75, in_chans=3, pos_embed="sincos",
                encoder_depth=12, encoder_embed_dim=768, encoder_mlp_dim=3072, encoder_num_heads=12,
                decoder_depth=8, decoder_embed_dim=768, decoder_mlp_dim=3072, decoder_num_heads=16,
                norm_pix_loss=False, use_bias=True)
# small shapes that still exercise hd=64 (encoder) and hd=48 (decoder), ragged S (17 / 65)
MAE_SMALL = dict(input_size=48, patch_size=12, mask_ratio=0.75, in_chans=3, pos_embed="sincos",
                 encoder_depth=2, encoder_embed_dim=192, encoder_mlp_dim=384, encoder_num_heads=3,
                 decoder_depth=2, decoder_embed_dim=96, decoder_mlp_dim=256, decoder_num_heads=2,
                 norm_pix_loss=True, use_bias=True)
VIT_FULL_DINO = dict(in_chans=3, img_size=96, patch_size=12, hidden_size=768, mlp_dim=3072,
                     num_layers=12, num_heads=12, pos_embed="sincos", num_register_tokens=4,
                     qkv_bias=True)
VIT_FULL_EXTRACT = dict(in_chans=3, img_size=96, patch_size=12, hidden_size=768, mlp_dim=3072,
                        num_layers=12, num_heads=12, pos_embed="sincos", num_register_tokens=0,
                        qkv_bias=False)
VIT_SMALL = dict(in_chans=3, img_size=48, patch_size=12, hidden_size=192, mlp_dim=384,
                 num_layers=2, num_heads=3, pos_embed="sincos", num_register_tokens=4, qkv_bias=True)
# downstream variants: LoRA adapters on q/v (TRAIN.LORA) and NORM_LAYER: 'rmsnorm'.
# "norm_layer" is kept as a string here (json-able); `resolve_norm` turns it into the class for a constructor.
VIT_SMALL_LORA = dict(VIT_SMALL, lora=True)
VIT_SMALL_RMS = dict(VIT_SMALL, norm_layer="rmsnorm", num_register_tokens=0, qkv_bias=False)
DINO_HEAD_FULL = dict(in_dim=768, out_dim=65536, nlayers=3, hidden_dim=2048, bottleneck_dim=256)
DINO_HEAD_SMALL = dict(in_dim=192, out_dim=1024, nlayers=3, hidden_dim=256, bottleneck_dim=64)


class _Gen:
    def __init__(self, seed: int):
        self.rng = np.random.default_rng(seed)

    def normal(self, shape, std=1.0, mean=0.0) -> torch.Tensor:
        a = self.rng.standard_normal(size=tuple(shape), dtype=np.float32)
        return torch.from_numpy(a * np.float32(std) + np.float32(mean))


def _linear(sd, g: _Gen, name: str, out_f: int, in_f: int, bias: bool = True):
    sd[name + ".weight"] = g.normal((out_f, in_f), std=(2.0 / (in_f + out_f)) ** 0.5)
    if bias:
        sd[name + ".bias"] = g.normal((out_f,), std=0.02)


def resolve_norm(cfg: Dict, rmsnorm_cls, layernorm_cls=torch.nn.LayerNorm) -> Dict:
    """Constructor kwargs from a synth config: the "norm_layer" string becomes the class."""
    out = dict(cfg)
    if "norm_layer" in out:
        out["norm_layer"] = rmsnorm_cls if out["norm_layer"] == "rmsnorm" else layernorm_cls
    return out


def _block(sd, g: _Gen, pre: str, dim: int, mlp: int, qkv_bias: bool, lora: bool = False, rms: bool = False):
    # child order mlp, att_norm, ffn_norm, attn (attentionblock.py:91-94)
    _linear(sd, g, pre + "mlp.linear1", mlp, dim)
    _linear(sd, g, pre + "mlp.linear2", dim, mlp)
    for n in ("att_norm", "ffn_norm"):
        sd[pre + n + ".weight"] = g.normal((dim,), std=0.1, mean=1.0)
        if not rms:
            sd[pre + n + ".bias"] = g.normal((dim,), std=0.05)
    _linear(sd, g, pre + "attn.qkv", 3 * dim, dim, bias=qkv_bias)
    _linear(sd, g, pre + "attn.proj", dim, dim)
    if lora:      # registration order B then A, r = 128 (attentionblock.py:18-19, :45-47); B made non-zero on purpose
        for n in ("lora_q", "lora_v"):
            sd[pre + f"attn.{n}.lora_matrix_B"] = g.normal((dim, 128), std=0.05)
            sd[pre + f"attn.{n}.lora_matrix_A"] = g.normal((128, dim), std=dim ** -0.5)


def _patch_embed(sd, g: _Gen, in_chans: int, size: int, patch: int, dim: int, pos_embed: str):
    grid = size // patch
    n = grid ** 3
    if pos_embed == "sincos":
        pos = O.sincos_pos_embed_3d((grid,) * 3, dim)
    else:
        pos = g.normal((1, n, dim), std=0.02)
    sd["patch_embedding.position_embeddings"] = pos.clone()
    fan_in = in_chans * patch ** 3
    sd["patch_embedding.patch_embeddings.weight"] = g.normal((dim, in_chans, patch, patch, patch),
                                                             std=fan_in ** -0.5)
    sd["patch_embedding.patch_embeddings.bias"] = g.normal((dim,), std=0.02)


def mae_state_dict(cfg: Dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    g = _Gen(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    E, Dd = cfg["encoder_embed_dim"], cfg["decoder_embed_dim"]
    grid = cfg["input_size"] // cfg["patch_size"]
    n = grid ** 3
    sd["cls_token"] = g.normal((1, 1, E), std=0.02)
    sd["decoder_cls_token"] = g.normal((1, 1, Dd), std=0.02)
    if cfg["pos_embed"] == "sincos":
        sd["decoder_pos_embed"] = O.sincos_pos_embed_3d((grid,) * 3, Dd).clone()
    else:
        sd["decoder_pos_embed"] = g.normal((1, n, Dd), std=0.02)
    sd["mask_token"] = g.normal((1, 1, Dd), std=0.02)
    _patch_embed(sd, g, cfg["in_chans"], cfg["input_size"], cfg["patch_size"], E, cfg["pos_embed"])
    for i in range(cfg["encoder_depth"]):
        _block(sd, g, f"blocks.{i}.", E, cfg["encoder_mlp_dim"], cfg["use_bias"])
    for i in range(cfg["decoder_depth"]):
        _block(sd, g, f"decoder_blocks.{i}.", Dd, cfg["decoder_mlp_dim"], cfg["use_bias"])
    for nm, d in (("norm", E), ("decoder_norm", Dd)):
        sd[nm + ".weight"] = g.normal((d,), std=0.1, mean=1.0)
        sd[nm + ".bias"] = g.normal((d,), std=0.05)
    _linear(sd, g, "decoder_embed", Dd, E, bias=cfg["use_bias"])
    _linear(sd, g, "decoder_pred", cfg["patch_size"] ** 3 * cfg["in_chans"], Dd, bias=cfg["use_bias"])
    return sd


def vit_state_dict(cfg: Dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    g = _Gen(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    H = cfg["hidden_size"]
    sd["cls_token"] = g.normal((1, 1, H), std=0.02)
    if cfg.get("num_register_tokens", 0):
        sd["register_tokens"] = g.normal((1, cfg["num_register_tokens"], H), std=0.02)
    _patch_embed(sd, g, cfg["in_chans"], cfg["img_size"], cfg["patch_size"], H, cfg["pos_embed"])
    rms = cfg.get("norm_layer") == "rmsnorm"
    for i in range(cfg["num_layers"]):
        _block(sd, g, f"blocks.{i}.", H, cfg["mlp_dim"], cfg["qkv_bias"], lora=cfg.get("lora", False), rms=rms)
    sd["norm.weight"] = g.normal((H,), std=0.1, mean=1.0)
    if not rms:
        sd["norm.bias"] = g.normal((H,), std=0.05)
    return sd


def dino_head_state_dict(cfg: Dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    g = _Gen(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    dims = [cfg["in_dim"]] + [cfg["hidden_dim"]] * (cfg["nlayers"] - 1) + [cfg["bottleneck_dim"]]
    for li in range(cfg["nlayers"]):
        _linear(sd, g, f"mlp.{2 * li}", dims[li + 1], dims[li])
    sd["last_layer.weight_g"] = torch.ones(cfg["out_dim"], 1)
    sd["last_layer.weight_v"] = g.normal((cfg["out_dim"], cfg["bottleneck_dim"]), std=0.02)
    return sd


def linear_classifier_state_dict(dim: int, num_classes: int, seed: int = 0):
    g = _Gen(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["bn.running_mean"] = torch.zeros(dim)
    sd["bn.running_var"] = torch.ones(dim)
    sd["bn.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    _linear(sd, g, "linear", num_classes, dim)
    return sd


def attention_classifier_state_dict(dim: int, num_classes: int, num_queries: int = 1, qkv_bias: bool = False, seed: int = 0):
    """Registration order of src/models/classifier.py:64-71: bn1, bn2, wkv, linear modules; cls_token parameter first."""
    g = _Gen(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["cls_token"] = g.normal((1, num_queries, dim), std=0.5)
    for n in ("bn1", "bn2"):
        sd[n + ".running_mean"] = torch.zeros(dim)
        sd[n + ".running_var"] = torch.ones(dim)
        sd[n + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    _linear(sd, g, "wkv", 2 * dim, dim, bias=qkv_bias)
    _linear(sd, g, "linear", num_classes, dim)
    return sd


def volume(batch: int, chans: int, size: int, seed: int) -> torch.Tensor:
    """Uniform [0,1) like post-window CT intensities (transforms.py:26-28)."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.random((batch, chans, size, size, size), dtype=np.float32))


def hu_volume(batch: int, size: int, seed: int) -> torch.Tensor:
    """Integer HU in [-1024, 3072): covers below / inside / above all three windows."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.integers(-1024, 3072, (batch, 1, size, size, size)).astype(np.float32))


def noise(batch: int, length: int, seed: int, ties: bool = False) -> torch.Tensor:
    """Mask noise on the fp32 2^-24 lattice torch.rand uses; `ties=True` quantises to 1/64 so
    that every row contains many equal keys (exercises the stable tie rule)."""
    rng = np.random.default_rng(seed)
    if ties:
        a = rng.integers(0, 64, (batch, length)).astype(np.float32) / np.float32(64)
    else:
        a = rng.integers(0, 1 << 24, (batch, length)).astype(np.float32) * np.float32(2.0 ** -24)
    return torch.from_numpy(a)


def to_device(sd, device, dtype=None):
    return OrderedDict((k, v.to(device=device, dtype=dtype if v.is_floating_point() else None))
                       for k, v in sd.items())
