"""CPU oracle: a plain fp32 restatement of the HeadCT-Foundation 3D-ViT hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under `headct_foundation_b200/` may import this file; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs do,
and there only as the checker / the CPU arm -- never as the product path.

Parity status: PINNED.  Every function below is checked in `tests/test_oracle_golden.py`
against golden vectors produced by executing the UNMODIFIED reference modules from
/root/reference (`oracle/gen_golden.py`, vectors committed under `tests/golden/`).  The
reference ships no tests or fixtures of its own, so those generated
vectors are the only pin available.

The oracle is functional: it works on a flat `state_dict`-style mapping of fp32 tensors (the
layout in SURVEY.md Appendix A) and never touches nn.Module code from the reference.
Citations are `file:line` relative to the reference tree.
"""
from __future__ import annotations

import math
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Mapping[str, Tensor]

# ----------------------------------------------------------------------------------------
# a1  HU windowing  (src/data/transforms.py:13-36, windows at :130)
# ----------------------------------------------------------------------------------------
HU_WINDOWS: Tuple[Tuple[int, int], ...] = ((40, 80), (80, 200), (600, 2800))


def window_bounds(windows: Sequence[Tuple[int, int]] = HU_WINDOWS) -> List[Tuple[float, float]]:
    """(a_min, a_max) per window: `l - w // 2`, `l + w // 2` (transforms.py:23-24)."""
    return [(float(l - w // 2), float(l + w // 2)) for l, w in windows]


def window_scale_stack(hu: Tensor, windows: Sequence[Tuple[int, int]] = HU_WINDOWS) -> Tensor:
    """[..., 1, D, H, W] HU -> [..., 3, D, H, W] in [0, 1].

    MONAI ScaleIntensityRange with b_min=0, b_max=1, clip=True:
    `(x - a_min) / (a_max - a_min)` then clamp to [0, 1]  (transforms.py:22-30, :34).
    """
    x = hu.to(torch.float32)
    chans = []
    for a_min, a_max in window_bounds(windows):
        y = (x - a_min) / (a_max - a_min)
        chans.append(y.clamp(0.0, 1.0))
    return torch.cat(chans, dim=-4)


# ----------------------------------------------------------------------------------------
# a3  3-D sin-cos table  (src/utils/pos_embed.py:51-78)
# ----------------------------------------------------------------------------------------
def sincos_pos_embed_3d(grid: Sequence[int], dim: int, temperature: float = 10000.0) -> Tensor:
    """[1, h*w*d, dim].  Token index = (i*w' + j)*d + k over a meshgrid built from
    (arange(w), arange(h), arange(d)) -- the reference swaps the h/w names (pos_embed.py:54-58)
    and concatenates sin/cos of the *second* axis first (:68-77)."""
    h, w, d = grid
    assert dim % 6 == 0
    nfreq = dim // 6
    omega = 1.0 / (temperature ** (torch.arange(nfreq, dtype=torch.float32) / nfreq))
    a0 = torch.arange(w, dtype=torch.float32)  # called grid_h in the reference
    a1 = torch.arange(h, dtype=torch.float32)  # called grid_w
    a2 = torch.arange(d, dtype=torch.float32)
    g0, g1, g2 = torch.meshgrid(a0, a1, a2, indexing="ij")
    o0 = g0.reshape(-1, 1) * omega[None, :]
    o1 = g1.reshape(-1, 1) * omega[None, :]
    o2 = g2.reshape(-1, 1) * omega[None, :]
    table = torch.cat([o1.sin(), o1.cos(), o0.sin(), o0.cos(), o2.sin(), o2.cos()], dim=1)
    return table[None]


# ----------------------------------------------------------------------------------------
# a2  patch embedding  (src/utils/patch_embedding.py:135-161)
# ----------------------------------------------------------------------------------------
def im2col_patches(x: Tensor, patch: Sequence[int]) -> Tensor:
    """[B,C,H,W,D] -> [B, gh*gw*gd, C*ph*pw*pd]; K order (c, ph, pw, pd) = Conv3d weight order."""
    B, C, H, W, D = x.shape
    ph, pw, pd = patch
    gh, gw, gd = H // ph, W // pw, D // pd
    x = x.reshape(B, C, gh, ph, gw, pw, gd, pd)
    x = x.permute(0, 2, 4, 6, 1, 3, 5, 7)
    return x.reshape(B, gh * gw * gd, C * ph * pw * pd)


def patch_embed(x: Tensor, weight: Tensor, bias: Tensor, pos: Optional[Tensor]) -> Tensor:
    """Conv3d(k = s = patch) restated as a GEMM over non-overlapping patches, then
    `flatten(2).transpose(-1,-2)` token order (gh, gw, gd) and `+ position_embeddings`
    (patch_embedding.py:149-156)."""
    patch = weight.shape[2:]
    cols = im2col_patches(x.float(), patch)
    out = cols @ weight.reshape(weight.shape[0], -1).t() + bias
    if pos is not None:
        out = out + pos
    return out


# ----------------------------------------------------------------------------------------
# a4  random masking  (src/models/mae.py:194-218)
# ----------------------------------------------------------------------------------------
def masking_indices(noise: Tensor, mask_ratio: float) -> Tuple[Tensor, Tensor, Tensor, int]:
    """noise [N,L] -> (ids_keep [N,len_keep], ids_restore [N,L], mask [N,L] f32, len_keep).

    `len_keep = int(L * (1 - mask_ratio))` (mae.py:205).  The reference calls
    `torch.argsort(noise)` (unstable by default, mae.py:208); the tie rule adopted by this
    project is STABLE ascending (equal keys keep index order), which is what the CUDA radix
    path of torch produces.  `ids_restore` is the inverse permutation (mae.py:209)."""
    N, L = noise.shape
    len_keep = int(L * (1 - mask_ratio))
    ids_shuffle = torch.argsort(noise, dim=1, stable=True)
    ids_restore = torch.empty_like(ids_shuffle)
    ar = torch.arange(L, dtype=ids_shuffle.dtype, device=noise.device).expand(N, L)
    ids_restore.scatter_(1, ids_shuffle, ar)
    ids_keep = ids_shuffle[:, :len_keep]
    mask = (ids_restore >= len_keep).to(torch.float32)  # == gather([0]*keep+[1]*rest, ids_restore)
    return ids_keep, ids_restore, mask, len_keep


def random_masking(x: Tensor, noise: Tensor, mask_ratio: float):
    ids_keep, ids_restore, mask, _ = masking_indices(noise, mask_ratio)
    x_masked = torch.gather(x, 1, ids_keep[:, :, None].expand(-1, -1, x.shape[2]))
    return x_masked, mask, ids_restore, ids_keep


# ----------------------------------------------------------------------------------------
# a6  transformer block  (src/models/attentionblock.py:51-66, :96-99; monai MLPBlock)
# ----------------------------------------------------------------------------------------
def _rms(x: Tensor, w: Tensor, eps: float) -> Tensor:
    """RMSNorm, src/models/layers.py:29-53 (NORM_LAYER: 'rmsnorm', main_downstream.py:111-116)."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def _norm(x: Tensor, sd: SD, name: str, eps: float) -> Tensor:
    """nn.LayerNorm, or RMSNorm when the state_dict holds no bias for this norm."""
    if (name + ".bias") in sd:
        return _ln(x, sd[name + ".weight"], sd[name + ".bias"], eps)
    return _rms(x, sd[name + ".weight"], eps)


def _ln(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def _gelu_erf(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def self_attention(x: Tensor, sd: SD, pre: str, heads: int) -> Tensor:
    B, S, C = x.shape
    hd = C // heads
    qkv = x @ sd[pre + "qkv.weight"].t()
    if (pre + "qkv.bias") in sd:
        qkv = qkv + sd[pre + "qkv.bias"]
    # output channels ordered [3][heads][hd]  (attentionblock.py:54)
    qkv = qkv.reshape(B, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if (pre + "lora_q.lora_matrix_A") in sd:
        # LoRA residuals (attentionblock.py:19-22, :57-59): x (B A)^T, RESHAPED (not permuted) to [B, H, S, hd]
        lq = x @ (sd[pre + "lora_q.lora_matrix_B"] @ sd[pre + "lora_q.lora_matrix_A"]).t()
        lv = x @ (sd[pre + "lora_v.lora_matrix_B"] @ sd[pre + "lora_v.lora_matrix_A"]).t()
        q = q + lq.reshape(B, heads, S, hd)
        v = v + lv.reshape(B, heads, S, hd)
    att = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))  # SDPA default scale (:61)
    att = att.softmax(-1)
    y = (att @ v).transpose(1, 2).reshape(B, S, C)
    return y @ sd[pre + "proj.weight"].t() + sd[pre + "proj.bias"]


def attention_block(x: Tensor, sd: SD, pre: str, heads: int, eps: Optional[float] = None) -> Tensor:
    if eps is None:      # constructor defaults: nn.LayerNorm 1e-5, RMSNorm 1e-6 (attentionblock.py:92-93, layers.py:12)
        eps = 1e-5 if (pre + "att_norm.bias") in sd else 1e-6
    h = _norm(x, sd, pre + "att_norm", eps)
    x = x + self_attention(h, sd, pre + "attn.", heads)
    h = _norm(x, sd, pre + "ffn_norm", eps)
    h = _gelu_erf(h @ sd[pre + "mlp.linear1.weight"].t() + sd[pre + "mlp.linear1.bias"])
    h = h @ sd[pre + "mlp.linear2.weight"].t() + sd[pre + "mlp.linear2.bias"]
    return x + h


def _depth(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}{n}.att_norm.weight" in sd:
        n += 1
    return n


def _linear(x: Tensor, sd: SD, name: str) -> Tensor:
    y = x @ sd[name + ".weight"].t()
    if (name + ".bias") in sd:
        y = y + sd[name + ".bias"]
    return y


# ----------------------------------------------------------------------------------------
# a5 / a7 / a8 / a9  MAE  (src/models/mae.py:220-317)
# ----------------------------------------------------------------------------------------
def mae_forward_encoder(sd: SD, x: Tensor, noise: Tensor, mask_ratio: float, enc_heads: int):
    t = patch_embed(x, sd["patch_embedding.patch_embeddings.weight"],
                    sd["patch_embedding.patch_embeddings.bias"],
                    sd.get("patch_embedding.position_embeddings"))
    t, mask, ids_restore, ids_keep = random_masking(t, noise, mask_ratio)
    t = torch.cat([sd["cls_token"].expand(t.shape[0], -1, -1), t], dim=1)
    for i in range(_depth(sd, "blocks.")):
        t = attention_block(t, sd, f"blocks.{i}.", enc_heads)
    t = _norm(t, sd, "norm", 1e-5 if "norm.bias" in sd else 1e-6)         # LayerNorm / RMSNorm default eps (mae.py:116)
    return t, mask, ids_restore, ids_keep


def mae_forward_decoder(sd: SD, latent: Tensor, ids_restore: Tensor, dec_heads: int) -> Tensor:
    B, L = ids_restore.shape
    y = _linear(latent, sd, "decoder_embed")
    C = y.shape[2]
    n_mask = L + 1 - y.shape[1]
    pool = torch.cat([y[:, 1:], sd["mask_token"].expand(B, n_mask, C)], dim=1)
    body = torch.gather(pool, 1, ids_restore[:, :, None].expand(-1, -1, C))   # un-shuffle (mae.py:259)
    y = torch.cat([y[:, :1], body], dim=1)
    pos = torch.cat([sd["decoder_cls_token"], sd["decoder_pos_embed"]], dim=1)  # (mae.py:262-264)
    y = y + pos
    for i in range(_depth(sd, "decoder_blocks.")):
        y = attention_block(y, sd, f"decoder_blocks.{i}.", dec_heads)
    y = _norm(y, sd, "decoder_norm", 1e-5 if "decoder_norm.bias" in sd else 1e-6)
    y = _linear(y, sd, "decoder_pred")
    return y[:, 1:]


def patchify_target(x: Tensor, patch: Sequence[int]) -> Tensor:
    """Loss-target layout: per-patch order (ph, pw, pd, c) -- channel LAST (mae.py:167-168)."""
    B, C, H, W, D = x.shape
    ph, pw, pd = patch
    gh, gw, gd = H // ph, W // pw, D // pd
    x = x.reshape(B, C, gh, ph, gw, pw, gd, pd).permute(0, 2, 4, 6, 3, 5, 7, 1)
    return x.reshape(B, gh * gw * gd, ph * pw * pd * C)


def mae_loss(x: Tensor, pred: Tensor, mask: Tensor, patch: Sequence[int], norm_pix: bool) -> Tensor:
    tgt = patchify_target(x.float(), patch)
    if norm_pix:
        mu = tgt.mean(-1, keepdim=True)
        var = tgt.var(-1, keepdim=True)            # unbiased (mae.py:292)
        tgt = (tgt - mu) / (var + 1e-6) ** 0.5
    per_patch = ((pred - tgt) ** 2).mean(-1)
    return (per_patch * mask).sum() / mask.sum()


def mae_forward(sd: SD, x: Tensor, noise: Tensor, *, patch: Sequence[int], mask_ratio: float,
                enc_heads: int, dec_heads: int, norm_pix: bool = False):
    latent, mask, ids_restore, ids_keep = mae_forward_encoder(sd, x, noise, mask_ratio, enc_heads)
    pred = mae_forward_decoder(sd, latent, ids_restore, dec_heads)
    loss = mae_loss(x, pred, mask, patch, norm_pix)
    return dict(loss=loss, latent=latent, pred=pred, mask=mask, ids_restore=ids_restore,
                ids_keep=ids_keep)


# ----------------------------------------------------------------------------------------
# a10  ViT  (src/models/vit.py:144-173)
# ----------------------------------------------------------------------------------------
def vit_forward(sd: SD, x: Tensor, heads: int):
    t = patch_embed(x, sd["patch_embedding.patch_embeddings.weight"],
                    sd["patch_embedding.patch_embeddings.bias"],
                    sd.get("patch_embedding.position_embeddings"))
    B = t.shape[0]
    parts = [sd["cls_token"].expand(B, -1, -1)]
    if "register_tokens" in sd:
        parts.append(sd["register_tokens"].expand(B, -1, -1))   # after cls (vit.py:152-160)
    t = torch.cat(parts + [t], dim=1)
    hidden = []
    for i in range(_depth(sd, "blocks.")):
        t = attention_block(t, sd, f"blocks.{i}.", heads)
        hidden.append(t)
    t = _norm(t, sd, "norm", 1e-6)                                # eps 1e-6 (vit.py:124)
    return t, hidden


# ----------------------------------------------------------------------------------------
# a12 / a11  DINO head and multi-crop wrapper  (dino_head.py:37-41, misc.py:463-484)
# ----------------------------------------------------------------------------------------
def dino_head(sd: SD, x: Tensor, pre: str = "") -> Tensor:
    i = 0
    keys = sorted({int(k[len(pre) + 4:].split(".")[0]) for k in sd
                   if k.startswith(pre + "mlp.") and k.endswith(".weight")})
    for n, li in enumerate(keys):
        x = x @ sd[f"{pre}mlp.{li}.weight"].t() + sd[f"{pre}mlp.{li}.bias"]
        if n + 1 < len(keys):
            x = _gelu_erf(x)
    x = x / x.norm(dim=-1, keepdim=True).clamp_min(1e-12)          # F.normalize(p=2)
    v = sd[pre + "last_layer.weight_v"]
    g = sd[pre + "last_layer.weight_g"]
    w = v * (g / v.norm(dim=1, keepdim=True))                      # weight_norm, dim=0
    return x @ w.t()


def multicrop_forward(backbone_sd: SD, head_sd: SD, crops: Sequence[Tensor], heads: int) -> Tensor:
    """All shipped crops share one size, so this is one backbone pass on the concatenation."""
    groups: List[List[Tensor]] = []
    for c in crops:
        if groups and groups[-1][0].shape[-1] == c.shape[-1]:
            groups[-1].append(c)
        else:
            groups.append([c])
    outs = [vit_forward(backbone_sd, torch.cat(g), heads)[0] for g in groups]
    cls = torch.cat(outs)[:, 0]
    return dino_head(head_sd, cls)


# ----------------------------------------------------------------------------------------
# a13  DINO loss + center update  (src/losses/losses.py:63-102)
# ----------------------------------------------------------------------------------------
def dino_loss(student: Tensor, teacher: Tensor, center: Tensor, *, ncrops: int,
              teacher_temp: float, student_temp: float = 0.1) -> Tensor:
    s = (student / student_temp).chunk(ncrops)
    q = F.softmax((teacher - center) / teacher_temp, dim=-1).detach().chunk(2)
    total, n = 0.0, 0
    for iq, qq in enumerate(q):
        for v in range(len(s)):
            if v == iq:
                continue
            total = total + (-(qq * F.log_softmax(s[v], dim=-1)).sum(-1)).mean()
            n += 1
    return total / n


def dino_center_update(center: Tensor, teacher: Tensor, momentum: float = 0.9,
                       world_size: int = 1, all_reduce=None) -> Tensor:
    bc = teacher.sum(0, keepdim=True)
    if all_reduce is not None:
        bc = all_reduce(bc)
    bc = bc / (teacher.shape[0] * world_size)
    return center * momentum + bc * (1 - momentum)


def teacher_temp_schedule(warmup_temp: float, temp: float, warmup_epochs: int, nepochs: int):
    import numpy as np
    return np.concatenate((np.linspace(warmup_temp, temp, warmup_epochs),
                           np.ones(nepochs - warmup_epochs) * temp))


# ----------------------------------------------------------------------------------------
# a14  EMA teacher  (src/utils/misc.py:386-397)
# ----------------------------------------------------------------------------------------
def ema_update(teacher: Sequence[Tensor], student: Sequence[Tensor], m: float) -> None:
    for pk, pq in zip(teacher, student):
        pk.mul_(m).add_((1 - m) * pq)


# ----------------------------------------------------------------------------------------
# a16  linear classifier  (src/models/classifier.py:21-33)
# ----------------------------------------------------------------------------------------
def linear_classifier(sd: SD, x: Tensor, training: bool = True) -> Tensor:
    if training:
        mu = x.mean(0)
        var = x.var(0, unbiased=False)
    else:
        mu, var = sd["bn.running_mean"], sd["bn.running_var"]
    x = (x - mu) / torch.sqrt(var + 1e-6)
    return x @ sd["linear.weight"].t() + sd["linear.bias"]


def attention_classifier(sd: SD, x: Tensor, heads: int, training: bool = True, qk_scale: Optional[float] = None) -> Tensor:
    """AttentionClassifier.forward, src/models/classifier.py:74-100.  x [B, N, C] tokens -> [B, classes]."""
    B, N, C = x.shape
    hd = C // heads
    scale = qk_scale or hd ** -0.5
    cls = sd["cls_token"]                                            # [1, nq, C]
    nq = cls.shape[1]
    q = cls.expand(B, -1, -1).reshape(B, nq, heads, hd).permute(0, 2, 1, 3) * scale          # :86-87
    if training:                                                     # bn1 over (batch, tokens) per channel (:89)
        mu, var = x.mean((0, 1)), x.var((0, 1), unbiased=False)
    else:
        mu, var = sd["bn1.running_mean"], sd["bn1.running_var"]
    xh = (x - mu) / torch.sqrt(var + 1e-6)
    kv = xh @ sd["wkv.weight"].t()
    if "wkv.bias" in sd:
        kv = kv + sd["wkv.bias"]
    kv = kv.reshape(B, N, 2, heads, hd).permute(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]
    att = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))          # SDPA scales AGAIN by 1/sqrt(hd) (:93)
    out = att.softmax(-1) @ v                                        # [B, H, nq, hd]
    x_cls = out.reshape(B, nq, C)                                    # :95 -- reshape of [B,H,nq,hd], no transpose
    if training:                                                     # bn2 over (batch, queries) per channel (:96)
        mu2, var2 = x_cls.mean((0, 1)), x_cls.var((0, 1), unbiased=False)
    else:
        mu2, var2 = sd["bn2.running_mean"], sd["bn2.running_var"]
    x_cls = ((x_cls - mu2) / torch.sqrt(var2 + 1e-6)).mean(1)
    return x_cls @ sd["linear.weight"].t() + sd["linear.bias"]


# ----------------------------------------------------------------------------------------
# next-row helpers (8(f) rank 1): per-parameter clip and AdamW, restated for the fused kernels
# (src/utils/misc.py:374-383; torch.optim.AdamW as called at src/utils/optimizers.py:354-360)
# ----------------------------------------------------------------------------------------
def clip_per_param(grads: Sequence[Tensor], clip: float) -> List[float]:
    norms = []
    for g in grads:
        n = float(g.norm(2))
        norms.append(n)
        coef = clip / (n + 1e-6)
        if coef < 1:
            g.mul_(coef)
    return norms


def adamw_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, *, lr: float, beta1: float,
               beta2: float, eps: float, weight_decay: float) -> None:
    p.mul_(1 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------------
# 8(f) rank 3: the random part of mae3d_transforms(mode='train') (src/data/transforms.py:195-236), CPU restatement.
# MONAI is not installed here (monai==1.3.2, setup.py:13): the published algorithms are restated -- RandFlip = numpy
# flip of a spatial axis, RandShiftIntensity = img + offset, GaussianFilter = separable zero-padded convolution with
# gaussian_1d(sigma, truncated=4, approx="erf").  PARITY UNPINNED against MONAI itself; the draws are inputs.
# --------------------------------------------------------------------------------------------
def flip_shift(vol: Tensor, flip_bits: Tensor, offsets: Tensor) -> Tensor:
    out = vol.float().clone()
    for b in range(vol.shape[0]):
        dims = [k + 1 for k in range(3) if (int(flip_bits[b]) >> k) & 1]
        if dims:
            out[b] = torch.flip(out[b], dims)
        out[b] += float(offsets[b])
    return out


def gaussian_kernel_1d(sigma: float, truncated: float = 4.0) -> Tensor:
    tail = max(int(sigma * truncated + 0.5), 1)
    x = torch.arange(-tail, tail + 1, dtype=torch.float64)
    t = 0.70710678 / abs(sigma)
    return (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0).float()


def gaussian_smooth(vol: Tensor, sigmas: Tensor) -> Tensor:
    """vol fp32 [B, C, D0, D1, D2]; sigmas [B, 3] (0 = sample not smoothed)."""
    out = vol.clone()
    for b in range(vol.shape[0]):
        if not bool((sigmas[b] > 0).all()):
            continue
        x = out[b][:, None]                                  # [C, 1, D0, D1, D2]
        for axis in range(3):
            k = gaussian_kernel_1d(float(sigmas[b, axis]))
            shape = [1, 1, 1, 1, 1]
            shape[2 + axis] = k.numel()
            pad = [0, 0, 0]
            pad[axis] = k.numel() // 2
            x = torch.nn.functional.conv3d(x, k.view(shape), padding=pad)
        out[b] = x[:, 0]
    return out


def crop_resize_area(src: Tensor, boxes: Tensor, out_size) -> Tensor:
    """DINO crop chain restated with torch: zero-pad/crop by indexing, then F.interpolate(mode='area') -- the op MONAI's
    Resize(mode='area') calls.  boxes int [n, 7] = (sample, start0..2, size0..2) in source coordinates."""
    outs = []
    B, C, S0, S1, S2 = src.shape
    for b, s0, s1, s2, n0, n1, n2 in boxes.tolist():
        crop = torch.zeros(C, n0, n1, n2)
        lo = [max(0, -s0), max(0, -s1), max(0, -s2)]
        hi = [min(n0, S0 - s0), min(n1, S1 - s1), min(n2, S2 - s2)]
        if all(h > l for l, h in zip(lo, hi)):
            crop[:, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = src[b, :, s0 + lo[0]:s0 + hi[0], s1 + lo[1]:s1 + hi[1],
                                                                 s2 + lo[2]:s2 + hi[2]].float()
        outs.append(torch.nn.functional.interpolate(crop[None], size=tuple(out_size), mode="area")[0])
    return torch.stack(outs)


def adjust_contrast(vol: Tensor, gamma: Tensor) -> Tensor:
    """MONAI AdjustContrast per sample where gamma > 0 (epsilon 1e-7, min / range over the whole sample)."""
    out = vol.clone()
    for b in range(vol.shape[0]):
        g = float(gamma[b])
        if g > 0:
            lo = out[b].min()
            rng = out[b].max() - lo
            out[b] = ((out[b] - lo) / (rng + 1e-7)) ** g * rng + lo
    return out
