import os, sys
sys.path.insert(0, "/root/repo")
import torch
from headct_foundation_b200._cabi import call, stream_ptr, lib
dev = torch.device("cuda")
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
for name, (B, S, H, hd) in {"dec512": (256, 512, 16, 48), "dec513": (256, 513, 16, 48), "enc128": (256, 128, 12, 64), "enc129": (256, 129, 12, 64)}.items():
    D = H * hd
    qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16(); out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
    do = torch.randn(B, S, D, device=dev).bfloat16(); lse = torch.empty(B, H, S, device=dev)
    dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, S, device=dev); st = stream_ptr(dev)
    f = timeit(lambda: call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st))
    b = timeit(lambda: call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st))
    print(f"{name}: fwd {f:.3f} ms  bwd {b:.3f} ms")
