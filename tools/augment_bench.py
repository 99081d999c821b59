"""HBM roofline of the on-GPU train-time augmentation (8(f) rank 3) at B = 256 cached volumes: python tools/augment_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200 import functional as HF
from headct_foundation_b200.data.transforms import MAE3DTrainAugment
dev = torch.device("cuda"); HBM = 6455.3


def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


B = 256
vol = torch.rand(B, 3, 96, 96, 96, device=dev).half()
aug = MAE3DTrainAugment(seed=0)
flips, offs, sig = aug.randomize(B)
ms = timeit(lambda: HF.flip_shift(vol, flips.to(dev), offs.to(dev)))
gb = vol.numel() * (2 + 4) / 1e9
print(f"flip_shift   B={B}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s ({gb / ms * 1e3 / HBM:.2f} of HBM copy bandwidth); algorithmic {gb:.2f} GB")
n_on = int((sig[:, 0] > 0).sum())
ms2 = timeit(lambda: aug.apply(vol, flips, offs, sig), n=5)
gb2 = gb + n_on * 3 * 96 ** 3 * 4 * 2 * 3 / 1e9
print(f"full augment B={B} ({n_on} smoothed samples): {ms2:.3f} ms  {gb2 / ms2 * 1e3:.0f} GB/s ({gb2 / ms2 * 1e3 / HBM:.2f}); algorithmic {gb2:.2f} GB (host tap construction included)")
import time
torch.cuda.synchronize(); t0 = time.perf_counter(); aug.apply(vol, flips, offs, sig); torch.cuda.synchronize()
print(f"one apply() wall clock incl. host work: {(time.perf_counter() - t0) * 1e3:.2f} ms")
on = torch.nonzero(sig[:, 0] > 0).flatten().to(torch.int32)
taps = [torch.rand(n_on, 9).to(dev) for _ in range(3)]
x32 = vol.float()
ms3 = timeit(lambda: HF.gaussian_smooth(x32, taps, on.to(dev)))
gb3 = n_on * 3 * 96 ** 3 * 4 * 2 * 3 / 1e9
print(f"gaussian (3 axes, {n_on} samples, radius 4): {ms3:.3f} ms  {gb3 / ms3 * 1e3:.0f} GB/s ({gb3 / ms3 * 1e3 / HBM:.2f} of HBM copy bandwidth)")

# ---- DINO multi-crop: 64 cached volumes of 224^3 x 3 (fp16) -> 4 crops of 96^3 each
from headct_foundation_b200.data.transforms import DataAugmentationDINO3D
Bd = 64
src = torch.rand(Bd, 3, 224, 224, 224, device=dev).half()
dino = DataAugmentationDINO3D((96, 96, 96), 112, 64, 2, seed=0)
draws = dino.randomize(Bd, src.shape[2:])
bx = draws["boxes"]
read_gb = float((bx[:, 4] * bx[:, 5] * bx[:, 6]).double().sum()) * 3 * 2 / 1e9
write_gb = 4 * Bd * 3 * 96 ** 3 * 4 / 1e9
msc = timeit(lambda: HF.crop_resize_area(src, draws["boxes"], (96, 96, 96), draws["flips"], draws["offsets"]), n=5)
print(f"DINO crop+area-resize+flip+shift, {4 * Bd} crops: {msc:.3f} ms  {(read_gb + write_gb) / msc * 1e3:.0f} GB/s "
      f"({(read_gb + write_gb) / msc * 1e3 / HBM:.2f} of HBM copy bandwidth); algorithmic {read_gb:.2f} GB read + {write_gb:.2f} GB written")
from headct_foundation_b200._cabi import lib as _lib
_lib().hct_crop_resize_set_rows(0)
mso = timeit(lambda: HF.crop_resize_area(src, draws["boxes"], (96, 96, 96), draws["flips"], draws["offsets"]), n=5)
_lib().hct_crop_resize_set_rows(1)
print(f"  (per-voxel gather kernel, for comparison: {mso:.3f} ms)")
msd = timeit(lambda: dino.apply(src, draws), n=5)
print(f"DINO full multi-crop augment B={Bd}: {msd:.3f} ms per batch ({Bd / msd * 1e3:.0f} volumes/s), "
      f"{int((draws['sigma'][:, 0] > 0).sum())} smoothed, {int((draws['gamma'] > 0).sum())} contrast-adjusted")
