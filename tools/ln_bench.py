"""Timing of the LayerNorm kernels at the MAE step shapes against the HBM roofline: python tools/ln_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200 import functional as HF
dev = torch.device("cuda")
HBM = 6455.3


def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for name, rows in (("dec", 256 * 513), ("enc", 256 * 129)):
    D = 768
    x = torch.randn(rows, D, device=dev); w = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
    dy = torch.randn(rows, D, device=dev).bfloat16(); dres = torch.randn(rows, D, device=dev)
    y, mean, rstd = HF.layernorm_fwd(x, w, b, 1e-5, True, True)
    f = timeit(lambda: HF.layernorm_fwd(x, w, b, 1e-5, True, True))
    g = timeit(lambda: HF.layernorm_bwd(dy, x, w, mean, rstd, dres, True, want_colsum=True))
    from headct_foundation_b200._cabi import lib
    lib().hct_layernorm_set_bulk(0)
    g0 = timeit(lambda: HF.layernorm_bwd(dy, x, w, mean, rstd, dres, True, want_colsum=True))
    lib().hct_layernorm_set_bulk(1)
    print(f"   (ln_bwd with per-lane cp.async staging: {g0:.3f} ms)")
    bf, bb = rows * D * 6 / 1e9, rows * D * 16 / 1e9      # algorithmic GB: fwd 4+2 B/elem, bwd 2+4+4 in, 4+2 out
    print(f"{name} rows={rows}: ln_fwd {f:.3f} ms ({bf / f * 1e3:.0f} GB/s, {bf / f * 1e3 / HBM:.2f} of HBM)   "
          f"ln_bwd {g:.3f} ms ({bb / g * 1e3:.0f} GB/s, {bb / g * 1e3 / HBM:.2f} of HBM)")
