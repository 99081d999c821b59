"""A few launches of ONE GEMM shape / epilogue of the MAE step (for ncu captures): python tools/gemm_one.py [gelu|mul|res|bf16] [dec|enc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200 import functional as HF
dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "gelu"
M = 33024 if (len(sys.argv) > 2 and sys.argv[2] == "enc") else 131328
N, K = (768, 768) if which == "res" else (3072, 768)
bias = torch.randn(N, device=dev)
if which == "mul":        # dX[M,N] = dY[M,K] @ W[K,N] * aux
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(K, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16); aux = torch.randn(M, N, device=dev).bfloat16(); cs = torch.zeros(N, device=dev)
    run = lambda: HF.gemm(A, W, M=M, N=N, K=K, lda=K, ldb=N, b_mn=True, out=out, ldo=N, epi=HF.EPI_MUL_BF16, aux=aux, ldaux=N, colsum=cs)
else:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    f32 = which == "res"
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if which == "gelu" else None
    res = torch.randn(M, N, device=dev) if f32 else None
    epi = {"gelu": HF.EPI_GELU_DERIV_BF16, "res": HF.EPI_RES_F32, "bf16": HF.EPI_BF16}[which]
    run = lambda: HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=epi, bias=bias, out2=out2, ldo2=N, res=res, ldres=N)
for _ in range(4):
    run()
torch.cuda.synchronize()
print("ok", which, M, N, K)
