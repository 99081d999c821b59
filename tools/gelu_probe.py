import os, sys
sys.path.insert(0, os.getcwd())
import torch
from headct_foundation_b200 import functional as HF
dev = torch.device("cuda")
def timeit(fn, iters=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
M, N, K = 131328, 3072, 768
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); out2 = torch.empty_like(out)
fl = 2.0 * M * N * K
for name, kw in (("bf16+bias", dict(epi=HF.EPI_BF16)), ("gelu no preact", dict(epi=HF.EPI_GELU_BF16)), ("gelu + preact", dict(epi=HF.EPI_GELU_BF16, out2=out2, ldo2=N))):
    ms = timeit(lambda: HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, bias=bias, **kw))
    print(f"{name:16s} {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s")
