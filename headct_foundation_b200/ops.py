"""torch.library registration of the hot-path ops (`torch.ops.headct.*`), so that `torch.compile` -- which the reference
wraps around the downstream model, main_downstream.py:162 -- traces THROUGH the drop-in modules instead of breaking the
graph around every one of them (SURVEY.md 8(b) "What calls it").

Each op pair is `torch.library.custom_op` + `register_fake` (shapes / dtypes for the tracer) + `register_autograd`:

    headct::embed / embed_bwd            EmbedFn        patch embedding + cls / register prefix
    headct::block / block_bwd            BlockFn        one pre-LN transformer block
    headct::layernorm / layernorm_bwd    LayerNormFn    final norms
    headct::linear / linear_bwd          LinearFn       decoder_embed, decoder_pred, head MLPs
    headct::mae_loss / mae_loss_bwd      MaeLossFn      masked MSE

The op bodies are the SAME code eager mode runs -- the static forward / backward of the autograd.Functions in
functional.py, driven with a stand-in context object -- so compiled and eager results are bit-identical and the C ABI
stays the only compute path.  What a Function saves for its backward does not travel through the traced graph: the
forward op parks the context in a process-local table and returns an int64 CPU scalar naming it, the backward op pops it.
(The ops are opaque and non-recomputable for the partitioner, each runs once per step, in order; contexts of forwards
that never see a backward are evicted oldest-first beyond a fixed cap.)  Eager mode keeps calling `Fn.apply` directly.
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import functional as HF

_CTX: Dict[int, "_Ctx"] = {}
_IDS = itertools.count(1)
_CTX_CAP = 16384


class _Ctx:
    """What an autograd.Function's static forward / backward expect of `ctx`."""

    def __init__(self, needs: Sequence[bool]):
        self.needs_input_grad = tuple(needs)
        self.saved_tensors: Tuple = ()

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors


def _park(ctx: _Ctx) -> Tensor:
    tok = next(_IDS)
    _CTX[tok] = ctx
    if len(_CTX) > _CTX_CAP:                 # forwards whose backward never ran (e.g. validation under enable_grad)
        for k in list(_CTX)[: len(_CTX) - _CTX_CAP]:
            del _CTX[k]
    return torch.tensor(tok, dtype=torch.int64)


def _take(token: Tensor) -> _Ctx:
    tok = int(token)
    ctx = _CTX.pop(tok, None)
    if ctx is None:
        raise RuntimeError(f"headct op context {tok} is gone: backward ran twice for one forward (retain_graph is not "
                           "supported through torch.compile) or more than %d forwards are pending" % _CTX_CAP)
    return ctx


def _fake_token() -> Tensor:
    return torch.empty((), dtype=torch.int64, device="cpu")


def _mask(tensors: Sequence[Optional[Tensor]]) -> int:
    return sum(1 << i for i, t in enumerate(tensors) if t is not None and t.requires_grad)


def _needs(mask: int, n: int, extra: int) -> List[bool]:
    return [bool(mask >> i & 1) for i in range(n)] + [False] * extra


def _pack_grads(grads: Sequence[Optional[Tensor]], mask: int, n: int) -> List[Tensor]:
    """Gradients of the inputs selected by `mask`, in input order (an op returns tensors only)."""
    out = []
    for i in range(n):
        if mask >> i & 1:
            g = grads[i]
            if g is None:
                raise RuntimeError(f"headct op backward produced no gradient for input {i}")
            out.append(g)
    return out


def _unpack_grads(packed: Sequence[Tensor], mask: int, n: int, extra: int) -> Tuple[Optional[Tensor], ...]:
    it = iter(packed)
    return tuple(next(it) if mask >> i & 1 else None for i in range(n)) + (None,) * extra


def _fake_grads(like: Sequence[Optional[Tensor]], mask: int) -> List[Tensor]:
    return [torch.empty_like(t) for i, t in enumerate(like) if mask >> i & 1]


# --------------------------------------------------------------------------------------------------------------------
# block
# --------------------------------------------------------------------------------------------------------------------
_NB = 13      # tensor inputs of a block without LoRA


@torch.library.custom_op("headct::block", mutates_args=())
def _block(x: Tensor, n1w: Tensor, n1b: Optional[Tensor], qkv_w: Tensor, qkv_b: Optional[Tensor], proj_w: Tensor,
           proj_b: Tensor, n2w: Tensor, n2b: Optional[Tensor], fc1_w: Tensor, fc1_b: Tensor, fc2_w: Tensor, fc2_b: Tensor,
           heads: int, eps: float, mask: int) -> Tuple[Tensor, Tensor]:
    ctx = _Ctx(_needs(mask, _NB, 6))
    out = HF.BlockFn.forward(ctx, x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b,
                             None, None, None, None, heads, eps)
    return out, (_park(ctx) if mask else torch.tensor(0, dtype=torch.int64))


@_block.register_fake
def _(x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b, heads, eps, mask):
    return torch.empty(x.shape, dtype=torch.float32, device=x.device), _fake_token()


@torch.library.custom_op("headct::block_bwd", mutates_args=())
def _block_bwd(dout: Tensor, token: Tensor, like: List[Tensor], mask: int) -> List[Tensor]:
    HF.ZeroArena.separate = True          # op outputs may not alias each other: no shared gradient arena here
    try:
        grads = HF.BlockFn.backward(_take(token), dout)
    finally:
        HF.ZeroArena.separate = False
    return _pack_grads(grads, mask, _NB)


@_block_bwd.register_fake
def _(dout, token, like, mask):
    return [torch.empty_like(t) for t in like]


def _block_setup(ctx, inputs, output):
    tensors = inputs[:_NB]
    ctx.mask = inputs[-1]
    ctx.save_for_backward(output[1], *[t for i, t in enumerate(tensors) if ctx.mask >> i & 1])
    ctx.set_materialize_grads(False)


def _block_backward(ctx, dout, _dtoken):
    token, *like = ctx.saved_tensors
    if dout is None:
        raise RuntimeError("headct::block: no gradient reached the block output")
    packed = torch.ops.headct.block_bwd(dout, token, like, ctx.mask)
    return _unpack_grads(packed, ctx.mask, _NB, 3)


_block.register_autograd(_block_backward, setup_context=_block_setup)


def block(x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b, heads: int, eps: float) -> Tensor:
    t = (x, n1w, n1b, qkv_w, qkv_b, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b)
    return torch.ops.headct.block(*t, heads, float(eps), _mask(t) if torch.is_grad_enabled() else 0)[0]


# --------------------------------------------------------------------------------------------------------------------
# patch embedding (+ prefix tokens)
# --------------------------------------------------------------------------------------------------------------------
_NE = 5       # vol, conv_w, conv_b, pos, prefix (ids_keep is an index tensor)


@torch.library.custom_op("headct::embed", mutates_args=())
def _embed(vol: Tensor, conv_w: Tensor, conv_b: Optional[Tensor], pos: Optional[Tensor], prefix: Optional[Tensor],
           ids_keep: Optional[Tensor], patch: int, mask: int) -> Tuple[Tensor, Tensor]:
    ctx = _Ctx(_needs(mask, _NE, 2))
    out = HF.EmbedFn.forward(ctx, vol, conv_w, conv_b, pos, prefix, ids_keep, patch)
    return out, (_park(ctx) if mask else torch.tensor(0, dtype=torch.int64))


@_embed.register_fake
def _(vol, conv_w, conv_b, pos, prefix, ids_keep, patch, mask):
    B = vol.shape[0]
    L = (vol.shape[2] // patch) * (vol.shape[3] // patch) * (vol.shape[4] // patch)
    n = L if ids_keep is None else ids_keep.shape[1]
    P = 0 if prefix is None else prefix.shape[-2]
    return torch.empty((B, P + n, conv_w.shape[0]), dtype=torch.float32, device=vol.device), _fake_token()


@torch.library.custom_op("headct::embed_bwd", mutates_args=())
def _embed_bwd(dout: Tensor, token: Tensor, like: List[Tensor], mask: int) -> List[Tensor]:
    grads = HF.EmbedFn.backward(_take(token), dout)
    return _pack_grads(grads, mask, _NE)


@_embed_bwd.register_fake
def _(dout, token, like, mask):
    return [torch.empty_like(t) for t in like]


def _embed_setup(ctx, inputs, output):
    ctx.mask = inputs[-1] & ~1                      # no gradient flows into the volume
    ctx.save_for_backward(output[1], *[t for i, t in enumerate(inputs[:_NE]) if ctx.mask >> i & 1])
    ctx.set_materialize_grads(False)


def _embed_backward(ctx, dout, _dtoken):
    token, *like = ctx.saved_tensors
    packed = torch.ops.headct.embed_bwd(dout, token, like, ctx.mask)
    return _unpack_grads(packed, ctx.mask, _NE, 3)


_embed.register_autograd(_embed_backward, setup_context=_embed_setup)


def embed(vol, conv_w, conv_b, pos, prefix, ids_keep, patch: int) -> Tensor:
    t = (None, conv_w, conv_b, pos, prefix)
    return torch.ops.headct.embed(vol, conv_w, conv_b, pos, prefix, ids_keep, int(patch),
                                  _mask(t) if torch.is_grad_enabled() else 0)[0]


# --------------------------------------------------------------------------------------------------------------------
# LayerNorm / RMSNorm
# --------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("headct::layernorm", mutates_args=())
def _layernorm(x: Tensor, w: Tensor, b: Optional[Tensor], eps: float, out_bf16: bool, mask: int) -> Tuple[Tensor, Tensor]:
    ctx = _Ctx(_needs(mask, 3, 2))
    out = HF.LayerNormFn.forward(ctx, x, w, b, eps, out_bf16)
    return out, (_park(ctx) if mask else torch.tensor(0, dtype=torch.int64))


@_layernorm.register_fake
def _(x, w, b, eps, out_bf16, mask):
    dt = torch.bfloat16 if (out_bf16 and HF.get_precision() != "fp32") else torch.float32
    return torch.empty(x.shape, dtype=dt, device=x.device), _fake_token()


@torch.library.custom_op("headct::layernorm_bwd", mutates_args=())
def _layernorm_bwd(dout: Tensor, token: Tensor, like: List[Tensor], mask: int) -> List[Tensor]:
    grads = HF.LayerNormFn.backward(_take(token), dout)
    return _pack_grads(grads, mask, 3)


@_layernorm_bwd.register_fake
def _(dout, token, like, mask):
    return [torch.empty_like(t) for t in like]


def _layernorm_setup(ctx, inputs, output):
    ctx.mask = inputs[-1]
    ctx.save_for_backward(output[1], *[t for i, t in enumerate(inputs[:3]) if ctx.mask >> i & 1])
    ctx.set_materialize_grads(False)


def _layernorm_backward(ctx, dout, _dtoken):
    token, *like = ctx.saved_tensors
    packed = torch.ops.headct.layernorm_bwd(dout, token, like, ctx.mask)
    return _unpack_grads(packed, ctx.mask, 3, 3)


_layernorm.register_autograd(_layernorm_backward, setup_context=_layernorm_setup)


def layernorm(x, w, b, eps: float, out_bf16: bool) -> Tensor:
    t = (x, w, b)
    return torch.ops.headct.layernorm(x, w, b, float(eps), bool(out_bf16), _mask(t) if torch.is_grad_enabled() else 0)[0]


# --------------------------------------------------------------------------------------------------------------------
# Linear
# --------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("headct::linear", mutates_args=())
def _linear(x: Tensor, w: Tensor, b: Optional[Tensor], gelu: bool, out_f32: bool, mask: int) -> Tuple[Tensor, Tensor]:
    ctx = _Ctx(_needs(mask, 3, 2))
    out = HF.LinearFn.forward(ctx, x, w, b, gelu, out_f32)
    return out, (_park(ctx) if mask else torch.tensor(0, dtype=torch.int64))


@_linear.register_fake
def _(x, w, b, gelu, out_f32, mask):
    fp32 = HF.get_precision() == "fp32" or (out_f32 and not gelu)
    return (torch.empty((*x.shape[:-1], w.shape[0]), dtype=torch.float32 if fp32 else torch.bfloat16, device=x.device),
            _fake_token())


@torch.library.custom_op("headct::linear_bwd", mutates_args=())
def _linear_bwd(dout: Tensor, token: Tensor, like: List[Tensor], mask: int) -> List[Tensor]:
    grads = HF.LinearFn.backward(_take(token), dout)
    return _pack_grads(grads, mask, 3)


@_linear_bwd.register_fake
def _(dout, token, like, mask):
    return [torch.empty_like(t) for t in like]


def _linear_setup(ctx, inputs, output):
    ctx.mask = inputs[-1]
    ctx.save_for_backward(output[1], *[t for i, t in enumerate(inputs[:3]) if ctx.mask >> i & 1])
    ctx.set_materialize_grads(False)


def _linear_backward(ctx, dout, _dtoken):
    token, *like = ctx.saved_tensors
    packed = torch.ops.headct.linear_bwd(dout, token, like, ctx.mask)
    return _unpack_grads(packed, ctx.mask, 3, 3)


_linear.register_autograd(_linear_backward, setup_context=_linear_setup)


def linear(x, w, b, gelu: bool, out_f32: bool) -> Tensor:
    t = (x, w, b)
    return torch.ops.headct.linear(x, w, b, bool(gelu), bool(out_f32), _mask(t) if torch.is_grad_enabled() else 0)[0]


# --------------------------------------------------------------------------------------------------------------------
# masked-MSE loss
# --------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("headct::mae_loss", mutates_args=())
def _mae_loss(pred: Tensor, imgs: Tensor, mask_t: Tensor, patch: int, norm_pix: bool, prefix: int, mask: int) -> Tuple[Tensor, Tensor]:
    ctx = _Ctx(_needs(mask, 1, 6))
    # not in place: under the tracer the prediction tensor belongs to the graph
    out = HF.MaeLossFn.forward(ctx, pred, imgs, mask_t, patch, norm_pix, False, prefix)
    return out.clone(), (_park(ctx) if mask else torch.tensor(0, dtype=torch.int64))


@_mae_loss.register_fake
def _(pred, imgs, mask_t, patch, norm_pix, prefix, mask):
    return torch.empty((), dtype=torch.float32, device=pred.device), _fake_token()


@torch.library.custom_op("headct::mae_loss_bwd", mutates_args=())
def _mae_loss_bwd(dloss: Tensor, token: Tensor, like: List[Tensor], mask: int) -> List[Tensor]:
    grads = HF.MaeLossFn.backward(_take(token), dloss)
    return _pack_grads(grads, mask, 1)


@_mae_loss_bwd.register_fake
def _(dloss, token, like, mask):
    return [torch.empty_like(t) for t in like]


def _mae_loss_setup(ctx, inputs, output):
    ctx.mask = inputs[-1]
    ctx.save_for_backward(output[1], *[t for i, t in enumerate(inputs[:1]) if ctx.mask >> i & 1])
    ctx.set_materialize_grads(False)


def _mae_loss_backward(ctx, dloss, _dtoken):
    token, *like = ctx.saved_tensors
    packed = torch.ops.headct.mae_loss_bwd(dloss, token, like, ctx.mask)
    return _unpack_grads(packed, ctx.mask, 1, 6)


_mae_loss.register_autograd(_mae_loss_backward, setup_context=_mae_loss_setup)


def mae_loss(pred, imgs, mask_t, patch: int, norm_pix: bool, prefix: int) -> Tensor:
    return torch.ops.headct.mae_loss(pred, imgs, mask_t, int(patch), bool(norm_pix), int(prefix),
                                     _mask((pred,)) if torch.is_grad_enabled() else 0)[0]


def pending_contexts() -> int:
    """Number of forward contexts waiting for their backward (0 after a completed training step)."""
    return len(_CTX)
