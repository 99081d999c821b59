"""Per-shape timing of the tcgen05 GEMM on the shapes of one MAE step (B=256): python tools/gemm_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200 import functional as HF

dev = torch.device("cuda")
PEAK = 1395.3


def timeit(fn, iters=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def fwd(M, N, K, epi):
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    f32 = epi in (HF.EPI_RES_F32, HF.EPI_F32)
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16) if epi in (HF.EPI_GELU_BF16, HF.EPI_GELU_DERIV_BF16) else None
    res = torch.randn(M, N, device=dev) if epi == HF.EPI_RES_F32 else None
    return lambda: HF.gemm(A, B, M=M, N=N, K=K, lda=K, ldb=K, out=out, ldo=N, epi=epi, bias=bias, out2=out2, ldo2=N,
                           res=res, ldres=N)


def dgrad(M, N, K, epi):   # dX[M,N] = dY[M,K] @ W[K,N]
    dY = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(K, N, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    aux = torch.randn(M, N, device=dev).bfloat16() if epi in (HF.EPI_DGELU_BF16, HF.EPI_MUL_BF16) else None
    cs = torch.zeros(N, device=dev) if epi in (HF.EPI_DGELU_BF16, HF.EPI_MUL_BF16) else None
    return lambda: HF.gemm(dY, W, M=M, N=N, K=K, lda=K, ldb=N, b_mn=True, out=out, ldo=N, epi=epi, aux=aux, ldaux=N, colsum=cs)


def wgrad(T, N, K):
    dY = torch.randn(T, N, device=dev).bfloat16(); X = torch.randn(T, K, device=dev).bfloat16()
    out = torch.zeros(N, K, device=dev)
    return lambda: HF.gemm(dY, X, M=N, N=K, K=T, lda=N, ldb=K, a_mn=True, b_mn=True, out=out, ldo=K, epi=HF.EPI_ATOMIC_F32)


cases = []
for tag, M in (("dec", 131328), ("enc", 33024)):
    cases += [(f"{tag} fwd qkv", fwd(M, 2304, 768, HF.EPI_BF16), M, 2304, 768),
              (f"{tag} fwd proj+res", fwd(M, 768, 768, HF.EPI_RES_F32), M, 768, 768),
              (f"{tag} fwd fc1+gelu+gelu'", fwd(M, 3072, 768, HF.EPI_GELU_DERIV_BF16), M, 3072, 768),
              (f"{tag} fwd fc2+res", fwd(M, 768, 3072, HF.EPI_RES_F32), M, 768, 3072),
              (f"{tag} dgrad da(*gelu')", dgrad(M, 3072, 768, HF.EPI_MUL_BF16), M, 3072, 768),
              (f"{tag} dgrad dh2", dgrad(M, 768, 3072, HF.EPI_BF16), M, 768, 3072),
              (f"{tag} dgrad datt", dgrad(M, 768, 768, HF.EPI_BF16), M, 768, 768),
              (f"{tag} dgrad dh1", dgrad(M, 768, 2304, HF.EPI_BF16), M, 768, 2304),
              (f"{tag} wgrad fc2", wgrad(M, 768, 3072), 768, 3072, M),
              (f"{tag} wgrad fc1", wgrad(M, 3072, 768), 3072, 768, M),
              (f"{tag} wgrad proj", wgrad(M, 768, 768), 768, 768, M),
              (f"{tag} wgrad qkv", wgrad(M, 2304, 768), 2304, 768, M)]
cases += [("pred fwd", fwd(131328, 5184, 768, HF.EPI_BF16), 131328, 5184, 768),
          ("pred dgrad", dgrad(131328, 768, 5184, HF.EPI_BF16), 131328, 768, 5184),
          ("pred wgrad", wgrad(131328, 5184, 768), 5184, 768, 131328),
          ("embed fwd", fwd(32768, 768, 5184, HF.EPI_F32), 32768, 768, 5184),
          ("embed wgrad", wgrad(32768, 768, 5184), 768, 5184, 32768)]
_a = torch.randn(8192, 8192, device=dev).bfloat16(); _b = torch.randn(8192, 8192, device=dev).bfloat16()
_ms = timeit(lambda: torch.matmul(_a, _b), iters=10)
print(f"box reference: cuBLAS bf16 8192^3 {_ms:.3f} ms = {2 * 8192 ** 3 / _ms / 1e9:.0f} TFLOP/s")
del _a, _b
tot_ms = tot_fl = 0
mult = {"dec": 8, "enc": 12}
print(f"{'case':28s} {'ms':>8s} {'TFLOP/s':>9s} {'of peak':>8s}")
for name, fn, M, N, K in cases:
    ms = timeit(fn)
    fl = 2.0 * M * N * K
    print(f"{name:28s} {ms:8.3f} {fl / ms / 1e9:9.1f} {fl / ms / 1e9 / PEAK:8.2f}")
    k = mult.get(name.split()[0], 1)
    tot_ms += ms * k; tot_fl += fl * k
print(f"weighted step total: {tot_ms:.1f} ms, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s")
