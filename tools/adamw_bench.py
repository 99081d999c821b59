"""Timing of the multi-tensor clip + AdamW launches on the MAE ViT-B parameter set: python tools/adamw_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import headct_foundation_b200 as H
from headct_foundation_b200.configs import MAE_HEADCT
from headct_foundation_b200.optim import FusedAdamW
dev = torch.device("cuda")
m = H.MaskedAutoencoderViT(**MAE_HEADCT).to(dev)
ps = [p for p in m.parameters() if p.requires_grad]
for p in ps:
    p.grad = torch.randn_like(p) * 1e-3
opt = FusedAdamW(ps, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, clip_grad=3.0)
for _ in range(3): opt.step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): opt.step()
b.record(); torch.cuda.synchronize()
n = sum(p.numel() for p in ps)
ms = a.elapsed_time(b) / 10
gb = n * (4 + 7 * 4) / 1e9
print(f"clip + AdamW over {len(ps)} tensors / {n / 1e6:.1f} M parameters: {ms:.3f} ms ({gb / ms * 1e3:.0f} GB/s, {gb / ms * 1e3 / 6455.3:.2f} of HBM)")
