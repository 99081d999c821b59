"""Host cost of issuing one MAE training step: at batch 2 the GPU work is tiny, so wall clock per step ~ the time the
Python / ctypes side needs to enqueue the step's ~490 launches (the floor for small per-GPU batches)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import headct_foundation_b200 as H
from headct_foundation_b200.configs import MAE_HEADCT
from headct_foundation_b200.optim import FusedAdamW

dev = torch.device("cuda")
m = H.MaskedAutoencoderViT(**MAE_HEADCT).to(dev).train()
opt = FusedAdamW([p for p in m.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, clip_grad=3.0)
for B in (2, 8, 32):
    x = torch.rand(B, 3, 96, 96, 96, device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        loss, _, _ = m(x)
        loss.backward()
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step()
    t_issue = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 10
    print(f"B={B}: host issue {t_issue * 1e3:.2f} ms/step, wall {t_all * 1e3:.2f} ms/step, {B / t_all:.0f} volumes/s")

print("-- the same step replayed from one CUDA graph (utils/graphs.py)")
for B in (2, 8, 32, 64, 256):
    x = torch.rand(B, 3, 96, 96, 96, device=dev)
    gstep = H.GraphedTrainStep(m, opt, x)
    for _ in range(3):
        gstep(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        gstep(x)
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / 10
    print(f"B={B}: graph replay wall {t_all * 1e3:.2f} ms/step, {B / t_all:.0f} volumes/s")
    del gstep
