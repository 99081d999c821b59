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
