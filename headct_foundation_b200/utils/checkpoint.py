"""Checkpoint wire format of the reference training scripts.

Host-side only; mirrors `save_checkpoint` / `load_optimizer` / `load_model` (src/utils/misc.py:35-96) and
`interpolate_pos_embed` (src/utils/pos_embed.py:102-153) so that files written by either code base load in the other:

    {"epoch", "best_loss", "state_dict", "momentum_model_state_dict", "optimizer", "scheduler"}

Keys may carry the wrappers' prefixes (`module.` from DDP, `backbone.` from MultiCropWrapper, `_orig_mod.` from
torch.compile); the loader removes them the way the reference does (plain substring removal, misc.py:78-79) and loads
with strict=False, so a released backbone checkpoint drops into ViT / MaskedAutoencoderViT / MultiCropWrapper.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Mapping, Optional

import torch

_PREFIXES = ("module.", "backbone.", "_orig_mod.")


def _say(logger, msg: str) -> None:
    if logger is not None:
        logger.info(msg)


def strip_wrapper_prefixes(state_dict: Mapping[str, Any]) -> Dict[str, Any]:
    """misc.py:78-79: every occurrence of the wrapper prefixes is removed from every key."""
    out = {}
    for k, v in state_dict.items():
        for pre in _PREFIXES:
            k = k.replace(pre, "")
        out[k] = v
    return out


def save_checkpoint(model, momentum_model, epoch, optimizer, scheduler, filename: str = "model.pt", best_loss=0,
                    dir_add: Optional[str] = None, logger=None) -> str:
    """misc.py:35-51.  Returns the path written."""
    payload = {
        "epoch": epoch,
        "best_loss": best_loss,
        "state_dict": model.state_dict(),
        "momentum_model_state_dict": None if momentum_model is None else momentum_model.state_dict(),
        "optimizer": optimizer.state_dict(),
        "scheduler": scheduler.state_dict(),
    }
    path = filename if dir_add is None else os.path.join(dir_add, filename)
    torch.save(payload, path)
    _say(logger, f"Saving checkpoint {path}")
    return path


def load_optimizer(optimizer, scheduler, loaded_state_dict: Mapping[str, Any], logger=None):
    """misc.py:54-69: restores whatever of optimizer / scheduler / epoch the file holds; epoch defaults to 0."""
    epoch = 0
    if "optimizer" in loaded_state_dict:
        optimizer.load_state_dict(loaded_state_dict["optimizer"])
        _say(logger, "Loaded optimizer state")
    if "scheduler" in loaded_state_dict:
        scheduler.load_state_dict(loaded_state_dict["scheduler"])
        _say(logger, "Loaded scheduler state")
    if "epoch" in loaded_state_dict:
        epoch = loaded_state_dict["epoch"]
        _say(logger, f"Loaded epoch: {epoch}")
    return optimizer, scheduler, epoch


def _pretrained_path(config_or_path) -> Optional[str]:
    if config_or_path is None or isinstance(config_or_path, (str, os.PathLike)):
        return config_or_path
    return config_or_path.MODEL.PRETRAINED            # yacs-style config, misc.py:74


def _read_checkpoint(path, trust_pickle: bool = False):
    """misc.py:73-76: the reference allow-lists numpy's scalar constructor (checkpoints carry `best_loss` as a numpy
    scalar) and loads with torch's safe unpickler.  Same here: `weights_only=True` with numpy scalar / dtype globals
    allow-listed, so a crafted .pt cannot execute code on load.  `trust_pickle=True` is the explicit opt-in for legacy
    files that hold arbitrary Python objects."""
    if trust_pickle:
        return torch.load(path, map_location=torch.device("cpu"), weights_only=False)
    import numpy as np
    safe = [np.dtype, np.ndarray]
    for mod in ("numpy.core.multiarray", "numpy._core.multiarray"):
        try:
            m = __import__(mod, fromlist=["scalar", "_reconstruct"])
            safe += [m.scalar, m._reconstruct]
        except Exception:          # module path differs between numpy 1.x and 2.x
            pass
    safe += [type(np.dtype(t)) for t in ("float64", "float32", "int64", "int32", "bool")]
    with torch.serialization.safe_globals(safe):
        return torch.load(path, map_location=torch.device("cpu"), weights_only=True)


def load_model(config_or_path, model, momentum_model=None, logger=None, model_name: str = "dino",
               interpolate_position_embeddings: bool = False, trust_pickle: bool = False):
    """misc.py:72-95.  `config_or_path`: the reference's config object (uses .MODEL.PRETRAINED) or a file path.
    Returns the whole checkpoint dict (for `load_optimizer`) or None when no checkpoint is configured.
    The reference keeps position-embedding interpolation commented out (misc.py:80-81); it is available here behind
    a flag because downstream runs at another resolution need it.  Files are read with torch's safe unpickler
    (`_read_checkpoint`); pass trust_pickle=True only for trusted legacy pickles."""
    path = _pretrained_path(config_or_path)
    if path is None:
        return None
    ckpt = _read_checkpoint(path, trust_pickle)
    sd = strip_wrapper_prefixes(ckpt["state_dict"])
    if interpolate_position_embeddings:
        interpolate_pos_embed(_unwrap(model), sd)
    msg = model.load_state_dict(sd, strict=False)
    _say(logger, f"Load Pretrained Model: {msg}")
    if momentum_model is not None:
        msd = strip_wrapper_prefixes(ckpt["momentum_model_state_dict"])
        if interpolate_position_embeddings:
            interpolate_pos_embed(_unwrap(momentum_model), msd)
        msg = momentum_model.load_state_dict(msd, strict=False)
        _say(logger, f"Load Pretrained Momentum Model: {msg}")
    return ckpt


def _unwrap(model):
    for attr in ("module", "backbone"):
        while hasattr(model, attr) and isinstance(getattr(model, attr), torch.nn.Module):
            model = getattr(model, attr)
    return model


def interpolate_pos_embed(model: torch.nn.Module, checkpoint_model: Dict[str, torch.Tensor], spatial_dims: int = 3) -> None:
    """pos_embed.py:102-153: resample `patch_embedding.position_embeddings` of a checkpoint (in place in the dict) to
    the grid of `model`.  Cubic grids (integer spatial_dims-th root of the patch count, as the reference assumes);
    leading extra tokens are kept; trilinear (3-D) / bicubic (2-D), align_corners=False."""
    key = "patch_embedding.position_embeddings"
    if key not in checkpoint_model:
        return
    if spatial_dims not in (2, 3):
        raise NotImplementedError(f"Spatial Dimension Size {spatial_dims} Not Implemented!")
    src = checkpoint_model[key]
    dim = src.shape[-1]
    n_new = model.patch_embedding.n_patches
    n_extra = model.patch_embedding.position_embeddings.shape[-2] - n_new
    side_old = int(round((src.shape[-2] - n_extra) ** (1.0 / spatial_dims)))
    side_new = int(round(n_new ** (1.0 / spatial_dims)))
    if side_old == side_new:
        return
    extra, grid = src[:, :n_extra], src[:, n_extra:]
    shape_old = (side_old,) * spatial_dims
    grid = grid.reshape(-1, *shape_old, dim).movedim(-1, 1)                     # [1, dim, *grid]
    grid = torch.nn.functional.interpolate(grid.float(), size=(side_new,) * spatial_dims,
                                           mode="trilinear" if spatial_dims == 3 else "bicubic", align_corners=False)
    grid = grid.movedim(1, -1).reshape(grid.shape[0], -1, dim).to(src.dtype)
    checkpoint_model[key] = torch.cat((extra, grid), dim=1)
