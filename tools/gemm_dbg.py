"""clock64 timeline of CTA 0 of one GEMM launch (hct_gemm_trace): where does the tile period go?
    python tools/gemm_dbg.py [gelu|mul|bf16|res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200 import functional as HF
from headct_foundation_b200._cabi import lib
dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "gelu"
M, N, K = 131328, 3072, 768
if which == "res":
    N = 768
A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16(); bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.float32 if which == "res" else torch.bfloat16)
out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16); aux = torch.randn(M, N, device=dev).bfloat16()
res = torch.randn(M, N, device=dev) if which == "res" else None
cs = torch.zeros(N, device=dev)
def run():
    if which == "gelu":
        HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_GELU_DERIV_BF16, bias=bias, out2=out2, ldo2=N)
    elif which == "mul":
        HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_MUL_BF16, aux=aux, ldaux=N, colsum=cs)
    elif which == "res":
        HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_RES_F32, bias=bias, res=res, ldres=N)
    else:
        HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=HF.EPI_BF16, bias=bias)
for _ in range(3): run()
torch.cuda.synchronize()
tr = torch.zeros(1536, dtype=torch.int64, device=dev)
lib().hct_gemm_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
lib().hct_gemm_trace(None)
t = tr.cpu()
mma = t[:512].view(64, 8); epi = t[512:].view(64, 16)
t0 = int(mma[0, 0])
print(f"{which}: M={M} N={N} K={K}   (cycles since the MMA warp's first event)")
print("tile | MMA: wait tempty, got, all MMAs issued+commit | EPI warp 4: wait tfull, got, then per unit (ld done, staged, drained) x4")
for i in range(2, 12):
    m = [int(v) - t0 if v > 0 else -1 for v in mma[i, :3]]
    e = [int(v) - t0 if v > 0 else -1 for v in epi[i, :14]]
    print(f"{i:3d}  | {m[0]:7d} {m[1]:7d} {m[2]:7d} | {e[0]:7d} {e[1]:7d} | " + " | ".join(f"{e[2+3*c]:7d} {e[3+3*c]:7d} {e[4+3*c]:7d}" for c in range(4)))
