"""ncu target: the two activation-epilogue GEMMs of a decoder block (fc1 + GELU + GELU' forward, fc2 dgrad * GELU') at B = 64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200 import functional as HF
sys.argv = sys.argv[:1]
import importlib.util
spec = importlib.util.spec_from_file_location("gb", os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_bench.py"))
dev = torch.device("cuda")
M = 64 * 513


def fwd(M, N, K, epi):
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    return lambda: HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=epi, bias=bias, out2=out2, ldo2=N)


def dgrad(M, N, K, epi):
    dY = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(K, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    aux = torch.randn(M, N, device=dev).bfloat16()
    cs = torch.zeros(N, device=dev)
    return lambda: HF.gemm(dY, W, M=M, N=N, K=K, lda=K, ldb=N, b_mn=True, out=out, ldo=N, epi=epi, aux=aux, ldaux=N, colsum=cs)


f = fwd(M, 3072, 768, HF.EPI_GELU_DERIV_BF16)
d = dgrad(M, 3072, 768, HF.EPI_MUL_BF16)
for _ in range(3):
    f(); d()
torch.cuda.synchronize()
print("ok")
