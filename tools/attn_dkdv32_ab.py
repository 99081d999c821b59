import os, sys
sys.path.insert(0, "/root/repo")
import torch
from headct_foundation_b200._cabi import call, stream_ptr, lib
dev = torch.device("cuda")
def timeit(fn, iters=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for name, (B, S, H, hd) in {"dec": (256, 513, 16, 48), "enc": (256, 129, 12, 64), "vit": (64, 517, 12, 64)}.items():
    D = H * hd
    qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16(); out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
    do = torch.randn(B, S, D, device=dev).bfloat16(); lse = torch.empty(B, H, S, device=dev)
    dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, S, device=dev); st = stream_ptr(dev)
    call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st)
    r = []
    for mode in (0, 1, 0, 1):
        lib().hct_attention_set_dkdv32(mode)
        r.append(timeit(lambda: call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st)))
    lib().hct_attention_set_dkdv32(0)
    print(f"{name}: bwd 64-wide {r[0]:.3f} / {r[2]:.3f} ms   pipelined 32-wide {r[1]:.3f} / {r[3]:.3f} ms")
