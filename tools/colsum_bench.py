"""Timing of the bias-gradient column sums at the step's shapes: python tools/colsum_bench.py"""
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
for name, rows, cols in (("dec dqkv", 256 * 513, 2304), ("enc dqkv", 256 * 129, 2304), ("dec d3", 256 * 513, 768)):
    x = torch.randn(rows, cols, device=dev).bfloat16()
    out = torch.zeros(cols, device=dev)
    ms = timeit(lambda: HF.colsum(x, cols, out=out))
    gb = rows * cols * 2 / 1e9
    ref = x[:4096].float().sum(0)
    got = HF.colsum(x[:4096].contiguous(), cols)
    err = ((got - ref).norm() / ref.norm()).item()
    print(f"{name}: {ms:.3f} ms ({gb / ms * 1e3:.0f} GB/s, {gb / ms * 1e3 / HBM:.2f} of HBM)  rel err {err:.1e}")
