"""On-GPU HU windowing.  Replaces the CPU/MONAI `MultipleWindowScaleStack` (`src/data/transforms.py:8-36`) for
volumes that already live on the device; the rest of the MONAI chain (I/O, resampling, crops) is out of scope."""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import torch

from .. import functional as HF

DEFAULT_WINDOWS: Tuple[Tuple[int, int], ...] = ((40, 80), (80, 200), (600, 2800))     # transforms.py:130


class MultipleWindowScaleStack:
    def __init__(self, keys: List[str], window_sizes: List[Tuple[int, int]] = list(DEFAULT_WINDOWS),
                 out_dtype: torch.dtype = torch.float32) -> None:
        self.keys = keys
        self.window_sizes = [tuple(w) for w in window_sizes]
        self.out_dtype = out_dtype

    def __call__(self, data: Dict[str, Any]) -> Dict[str, Any]:
        d = dict(data)
        d["image"] = HF.window_scale_stack(d["image"], self.window_sizes, self.out_dtype)
        return d


def gaussian_taps(sigma: float, truncated: float = 4.0, radius: int | None = None) -> torch.Tensor:
    """MONAI `gaussian_1d(sigma, truncated, approx="erf")`: tail = max(int(sigma * truncated + 0.5), 1), weights
    0.5 (erf((x + 0.5) / (sigma sqrt 2)) - erf((x - 0.5) / (sigma sqrt 2))) clamped at 0 (not renormalised).
    `radius` pads the kernel with zeros to a common length."""
    import math
    tail = max(int(float(sigma) * truncated + 0.5), 1)
    x = torch.arange(-tail, tail + 1, dtype=torch.float64)
    t = 0.70710678 / abs(float(sigma))
    w = (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0).float()
    if radius is not None and radius > tail:
        w = torch.nn.functional.pad(w, (radius - tail, radius - tail))
    return w


class MAE3DTrainAugment:
    """The random part of `mae3d_transforms(mode='train')` (src/data/transforms.py:195-236) for a batch of cached
    volumes that already sits in HBM: RandFlipd(prob) on each spatial axis, RandShiftIntensityd(offsets, prob_shift),
    optionally RandGaussianSmoothd(sigma in [0.5, 1], prob_smooth).  The draws are made on the host with a numpy
    RandomState (as MONAI does, one set per sample); two launches (+3 for a batch with smoothed samples) apply them.
    Input fp16 (the cache format) or fp32 [B, C, D0, D1, D2]; output fp32."""

    def __init__(self, prob_flip: float = 0.1, offsets: float = 0.1, prob_shift: float = 0.5, smooth: bool = True,
                 sigma=(0.5, 1.0), prob_smooth: float = 0.2, seed: int | None = None) -> None:
        import numpy as np
        self.prob_flip, self.offsets, self.prob_shift = prob_flip, offsets, prob_shift
        self.smooth, self.sigma, self.prob_smooth = smooth, tuple(sigma), prob_smooth
        self.R = np.random.RandomState(seed)

    def randomize(self, batch: int):
        """Per-sample draws: flip bits (uint8), offsets (fp32), sigmas (fp32 [B, 3], 0 = not smoothed)."""
        import numpy as np
        flips = np.zeros(batch, dtype=np.uint8)
        for k in range(3):
            flips |= ((self.R.rand(batch) < self.prob_flip).astype(np.uint8) << k)
        do_shift = self.R.rand(batch) < self.prob_shift
        offs = np.where(do_shift, self.R.uniform(-self.offsets, self.offsets, batch), 0.0).astype(np.float32)
        sig = np.zeros((batch, 3), dtype=np.float32)
        if self.smooth:
            on = self.R.rand(batch) < self.prob_smooth
            sig[on] = self.R.uniform(self.sigma[0], self.sigma[1], (int(on.sum()), 3)).astype(np.float32)
        return torch.from_numpy(flips), torch.from_numpy(offs), torch.from_numpy(sig)

    def apply(self, vol: torch.Tensor, flips: torch.Tensor, offs: torch.Tensor, sig: torch.Tensor) -> torch.Tensor:
        out = HF.flip_shift(vol, flips, offs)
        sig = sig.cpu()
        on = torch.nonzero(sig[:, 0] > 0).flatten()
        if on.numel() > 0:                                  # only the smoothed samples are filtered (in place)
            s_on = sig[on].double()                                           # [n, 3]
            tails = (s_on * 4.0 + 0.5).floor().clamp(min=1)                   # gaussian_1d: max(int(sigma * 4 + 0.5), 1)
            radius = int(tails.max())
            x = torch.arange(-radius, radius + 1, dtype=torch.float64)
            taps = []
            for k in range(3):                                                # all samples of an axis at once
                t = (0.70710678 / s_on[:, k]).unsqueeze(1)
                wk = (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0)
                wk = wk * (x.abs().unsqueeze(0) <= tails[:, k].unsqueeze(1))   # each sample's own truncation
                taps.append(wk.float())
            HF.gaussian_smooth(out, taps, on.to(torch.int32))
        return out

    def __call__(self, data: Dict[str, Any]) -> Dict[str, Any]:
        d = dict(data)
        d["image"] = self.apply(d["image"], *self.randomize(d["image"].shape[0]))
        return d


class ViTTrainAugment(MAE3DTrainAugment):
    """`vit_transforms(mode='train')` (src/data/transforms.py:270-303): the MAE chain without the Gaussian smoothing."""

    def __init__(self, prob_flip: float = 0.1, offsets: float = 0.1, prob_shift: float = 0.5, seed: int | None = None) -> None:
        super().__init__(prob_flip, offsets, prob_shift, smooth=False, seed=seed)


class DataAugmentationDINO3D:
    """`DataAugmentationDINO3D` (src/data/transforms.py:39-105) for a BATCH of cached volumes in HBM: returns the list
    [global_1, global_2, local_1 .. local_n] of fp32 [B, C, *final_size] tensors the reference's collate would build.

    Per crop the reference runs ResizeWithPadOrCrop(224) -> (CenterSpatialCrop(192) for local crops) -> RandSpatialCrop
    (random size and corner) -> Resize(final_size, mode='area') -> [global crops: RandFlip(0.2) x 3,
    RandShiftIntensity(0.2, prob 0.5); global 1: RandGaussianSmooth(prob 0.2); global 2: RandAdjustContrast(gamma in
    [0.2, 1], prob 0.2)].  Here every crop of the batch is ONE gather launch (pad / crop / area resize / flips / shift),
    followed by the in-place smoothing / contrast launches on the samples that drew them.  Draws come from a numpy
    RandomState on the host, one independent set per sample and crop, as MONAI's do."""

    CANVAS, LOCAL_CANVAS = 224, 192

    def __init__(self, final_size, global_crops_size: int, local_crops_size: int, local_crops_number: int,
                 seed: int | None = None) -> None:
        import numpy as np
        self.final_size = tuple(int(v) for v in final_size)
        self.g, self.l, self.n_local = int(global_crops_size), int(local_crops_size), int(local_crops_number)
        self.R = np.random.RandomState(seed)

    def _canvas_origin(self, S: int) -> int:
        """Source coordinate of canvas voxel 0 along an axis of length S (symmetric pad, else centre crop)."""
        return -((self.CANVAS - S) // 2) if S < self.CANVAS else S // 2 - self.CANVAS // 2

    def randomize(self, batch: int, spatial):
        """Draws for `batch` samples of spatial shape `spatial`: dict of CPU tensors (boxes [ncrops*B, 7] ordered crop
        major, flips, offsets, sigmas for global 1, gammas for global 2)."""
        import numpy as np
        R = self.R
        ncrops = 2 + self.n_local
        boxes = np.zeros((ncrops, batch, 7), dtype=np.int32)
        flips = np.zeros((ncrops, batch), dtype=np.uint8)
        offs = np.zeros((ncrops, batch), dtype=np.float32)
        for cidx in range(ncrops):
            is_global = cidx < 2
            canvas = self.CANVAS if is_global else self.LOCAL_CANVAS
            lo = self.g if is_global else self.l
            hi = canvas if is_global else min(self.g, canvas)          # max_roi_size = global size for local crops
            size = R.randint(lo, hi + 1, size=(batch, 3))
            corner = (R.rand(batch, 3) * (canvas - size + 1)).astype(np.int64)
            extra = 0 if is_global else (self.CANVAS - self.LOCAL_CANVAS) // 2
            origin = np.array([self._canvas_origin(int(s)) for s in spatial])
            boxes[cidx, :, 0] = np.arange(batch)
            boxes[cidx, :, 1:4] = origin[None, :] + extra + corner
            boxes[cidx, :, 4:7] = size
            if is_global:
                for k in range(3):
                    flips[cidx] |= ((R.rand(batch) < 0.2).astype(np.uint8) << k)
                sh = R.rand(batch) < 0.5
                offs[cidx] = np.where(sh, R.uniform(-0.2, 0.2, batch), 0.0)
        sig = np.zeros((batch, 3), dtype=np.float32)
        on = R.rand(batch) < 0.2
        sig[on] = R.uniform(0.5, 1.0, (int(on.sum()), 3))
        gam = np.where(R.rand(batch) < 0.2, R.uniform(0.2, 1.0, batch), 0.0).astype(np.float32)
        return dict(boxes=torch.from_numpy(boxes.reshape(-1, 7)), flips=torch.from_numpy(flips.reshape(-1)),
                    offsets=torch.from_numpy(offs.reshape(-1)), sigma=torch.from_numpy(sig), gamma=torch.from_numpy(gam))

    def apply(self, vol: torch.Tensor, draws) -> List[torch.Tensor]:
        B = vol.shape[0]
        ncrops = 2 + self.n_local
        out = HF.crop_resize_area(vol, draws["boxes"], self.final_size, draws["flips"], draws["offsets"])
        crops = [out[i * B:(i + 1) * B] for i in range(ncrops)]            # views of one allocation, crop major
        sig = draws["sigma"]
        on = torch.nonzero(sig[:, 0] > 0).flatten()
        if on.numel() > 0:
            s_on = sig[on].double()
            tails = (s_on * 4.0 + 0.5).floor().clamp(min=1)
            radius = int(tails.max())
            x = torch.arange(-radius, radius + 1, dtype=torch.float64)
            taps = []
            for k in range(3):
                t = (0.70710678 / s_on[:, k]).unsqueeze(1)
                wk = (0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())).clamp(min=0)
                taps.append((wk * (x.abs().unsqueeze(0) <= tails[:, k].unsqueeze(1))).float())
            HF.gaussian_smooth(crops[0], taps, on.to(torch.int32))
        if bool((draws["gamma"] > 0).any()):
            HF.adjust_contrast_(crops[1], draws["gamma"])
        return crops

    def __call__(self, image: torch.Tensor) -> List[torch.Tensor]:
        return self.apply(image, self.randomize(image.shape[0], image.shape[2:]))
