"""Import the UNMODIFIED reference modules from /root/reference (this container only).

TEST INFRASTRUCTURE -- never imported by the product package.

The reference (`src/models/*.py`, `src/utils/*.py`, `src/losses/losses.py`) imports a few
third-party symbols that are not installed here (monai, timm, matplotlib).  We register
minimal in-memory stand-ins for exactly the symbols the hot path touches:

  timm.models.layers.{to_2tuple,to_3tuple}          mae.py:15, pos_embed.py:7
  monai.networks.layers.{Conv,trunc_normal_}         mae.py:17, patch_embedding.py:22
  monai.utils.{ensure_tuple_rep,optional_import}     patch_embedding.py:23
  monai.utils.module.look_up_option                  patch_embedding.py:24
  monai.networks.blocks.mlp.MLPBlock                 attentionblock.py:4
  matplotlib.pyplot                                  src/utils/misc.py:6 (star import chain)

MLPBlock restates monai 1.3.x `linear2(drop2(GELU(linear1(x))))` (structure pinned by the
reference notebook's module repr, notebooks/extract_feature_sample.ipynb cell 2).

/root/reference does not exist on the GPU box; callers must gate on `available()`.
"""
from __future__ import annotations

import collections.abc
import importlib
import os
import sys
import types
from itertools import repeat

REF_ROOT = os.environ.get("HCT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "models", "mae.py"))


def _ntuple(n):
    def parse(x):
        if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
            return tuple(x)
        return tuple(repeat(x, n))
    return parse


def _install_stubs() -> None:
    import torch
    import torch.nn as nn

    if "monai" in sys.modules and getattr(sys.modules["monai"], "_hct_stub", False):
        return

    def mod(name):
        m = types.ModuleType(name)
        m.__path__ = []  # mark as package
        sys.modules[name] = m
        return m

    # ---- timm ----
    timm = mod("timm"); timm_models = mod("timm.models"); timm_layers = mod("timm.models.layers")
    timm_layers.to_2tuple = _ntuple(2)
    timm_layers.to_3tuple = _ntuple(3)
    timm.models = timm_models; timm_models.layers = timm_layers

    # ---- monai ----
    monai = mod("monai"); monai._hct_stub = True
    networks = mod("monai.networks"); layers = mod("monai.networks.layers")
    blocks = mod("monai.networks.blocks"); mlp = mod("monai.networks.blocks.mlp")
    utils = mod("monai.utils"); utils_module = mod("monai.utils.module")
    mtransforms = mod("monai.transforms")

    class _ConvFactory:
        CONV = "conv"

        def __getitem__(self, key):
            kind, dims = key
            assert kind == "conv"
            return {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}[dims]

    layers.Conv = _ConvFactory()
    layers.trunc_normal_ = torch.nn.init.trunc_normal_

    def ensure_tuple_rep(tup, dim):
        if isinstance(tup, torch.Tensor):
            tup = tup.detach().cpu().numpy()
        if not isinstance(tup, collections.abc.Iterable) or isinstance(tup, str):
            return (tup,) * dim
        if len(tup) == dim:
            return tuple(tup)
        raise ValueError(f"Sequence must have length {dim}, got {len(tup)}.")

    def optional_import(module, name="", **_):
        try:
            m = importlib.import_module(module)
            return (getattr(m, name) if name else m), True
        except Exception:  # pragma: no cover
            return None, False

    def look_up_option(opt, supported, default="no_default", print_all_options=True):
        if opt in supported:
            return opt
        if default != "no_default":
            return default
        raise ValueError(f"Unsupported option '{opt}', Available options are {set(supported)}.")

    utils.ensure_tuple_rep = ensure_tuple_rep
    utils.optional_import = optional_import
    utils_module.look_up_option = look_up_option
    utils.module = utils_module

    class MLPBlock(nn.Module):
        def __init__(self, hidden_size, mlp_dim, dropout_rate=0.0, act="GELU", dropout_mode="vit"):
            super().__init__()
            if not (0 <= dropout_rate <= 1):
                raise ValueError("dropout_rate should be between 0 and 1.")
            mlp_dim = mlp_dim or hidden_size
            self.linear1 = nn.Linear(hidden_size, mlp_dim)
            self.linear2 = nn.Linear(mlp_dim, hidden_size)
            self.fn = nn.GELU()
            self.drop1 = nn.Dropout(dropout_rate)
            self.drop2 = self.drop1

        def forward(self, x):
            x = self.fn(self.linear1(x))
            x = self.drop1(x)
            x = self.linear2(x)
            x = self.drop2(x)
            return x

    mlp.MLPBlock = MLPBlock
    blocks.mlp = mlp
    networks.layers = layers; networks.blocks = blocks
    monai.networks = networks; monai.utils = utils; monai.transforms = mtransforms

    # ---- matplotlib (only imported, never used on the hot path) ----
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = mod("matplotlib"); plt = mod("matplotlib.pyplot"); mpl.pyplot = plt


def load():
    """Return a namespace with the reference classes (unmodified code)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    # transformers probes `timm` with find_spec at import time: import it BEFORE the stubs exist
    # (src/utils/lr_sched.py:10-12 imports it, but only for names the hot path never uses).
    try:
        import transformers.trainer_utils  # noqa: F401
        import transformers.optimization  # noqa: F401
    except Exception:  # pragma: no cover
        pass
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    ns.mae = importlib.import_module("src.models.mae")
    ns.vit = importlib.import_module("src.models.vit")
    ns.attentionblock = importlib.import_module("src.models.attentionblock")
    ns.dino_head = importlib.import_module("src.models.dino_head")
    ns.classifier = importlib.import_module("src.models.classifier")
    ns.layers = importlib.import_module("src.models.layers")
    ns.patch_embedding = importlib.import_module("src.utils.patch_embedding")
    ns.pos_embed = importlib.import_module("src.utils.pos_embed")
    ns.misc = importlib.import_module("src.utils.misc")
    ns.losses = importlib.import_module("src.losses.losses")
    return ns
