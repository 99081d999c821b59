"""Timing of the masked-MSE loss kernels at the MAE step shape (B = 256, 75 % masked) against the HBM roofline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200 import functional as HF
from headct_foundation_b200._cabi import call, stream_ptr
dev = torch.device("cuda")
HBM = 6455.3
N, L, P = 256, 512, 5184
pred = torch.randn(N, L + 1, P, device=dev).bfloat16()
imgs = torch.rand(N, 3, 96, 96, 96, device=dev)
mask = (torch.rand(N, L, device=dev) < 0.75).float()
ws = torch.empty(4 + N * L, device=dev)
dpred = torch.empty_like(pred)
dloss = torch.ones(1, device=dev)
st = stream_ptr(dev)


def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


f = timeit(lambda: call("hct_mae_loss_fwd", pred.data_ptr(), 1, imgs.data_ptr(), mask.data_ptr(), ws.data_ptr(), N, 3, 96, 96, 96, 12, 0, st))
b = timeit(lambda: call("hct_mae_loss_bwd", pred.data_ptr(), 1, imgs.data_ptr(), mask.data_ptr(), dloss.data_ptr(), ws[1:2].data_ptr(),
                        dpred.data_ptr(), N, 3, 96, 96, 96, 12, 0, st))
nm = float(mask.sum())
gf = nm * P * (2 + 4) / 1e9                       # masked rows: bf16 prediction + fp32 target
gb = (nm * P * (2 + 4) + N * (L + 1) * P * 2) / 1e9   # + the full bf16 gradient written (zeros on the kept rows)
print(f"mae_loss fwd {f:.3f} ms ({gf / f * 1e3:.0f} GB/s, {gf / f * 1e3 / HBM:.2f} of HBM)   bwd {b:.3f} ms ({gb / b * 1e3:.0f} GB/s, {gb / b * 1e3 / HBM:.2f} of HBM)")
