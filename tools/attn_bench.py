"""Timing of the attention kernels at the MAE step shapes: python tools/attn_bench.py [modes, e.g. 2,1,0]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200._cabi import call, stream_ptr, lib

dev = torch.device("cuda")
MODES = [int(m) for m in (sys.argv[1] if len(sys.argv) > 1 else "2,3").split(",")]
NAMES = {2: "tcgen05 + bwd row-kernel tail", 3: "tcgen05 every tile", 1: "tcgen05+mma.sync fwd tail", 0: "mma.sync"}


def timeit(fn, iters=5):
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


for name, (B, S, H, hd) in {"dec": (256, 513, 16, 48), "enc": (256, 129, 12, 64), "vit": (64, 517, 12, 64)}.items():
    D = H * hd
    qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16()
    out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
    do = torch.randn(B, S, D, device=dev).bfloat16()
    lse = torch.empty(B, H, S, device=dev)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, S, device=dev)
    st = stream_ptr(dev)
    fl_f = 4.0 * B * H * S * S * hd
    for mode in MODES:
        lib().hct_attention_set_tcgen05(mode)
        f = timeit(lambda: call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st))
        b = timeit(lambda: call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st))
        print(f"{name} B={B} S={S} H={H} hd={hd} {NAMES[mode]}: fwd {f:.3f} ms ({fl_f / f / 1e9:.0f} TFLOP/s)  "
              f"bwd {b:.3f} ms ({2.5 * fl_f / b / 1e9:.0f} TFLOP/s algorithmic)")
lib().hct_attention_set_tcgen05(2)
print("-- backward with the tail block as its own chain step (hct_attention_set_merge_tail(0))")
lib().hct_attention_set_merge_tail(0)
for name, (B, S, H, hd) in {"dec": (256, 513, 16, 48), "enc": (256, 129, 12, 64), "vit": (64, 517, 12, 64)}.items():
    D = H * hd
    qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16(); out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
    do = torch.randn(B, S, D, device=dev).bfloat16(); lse = torch.empty(B, H, S, device=dev)
    dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, S, device=dev); st = stream_ptr(dev)
    call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st)
    b = timeit(lambda: call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(),
                            dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st))
    print(f"{name}: bwd {b:.3f} ms")
lib().hct_attention_set_merge_tail(1)

print("-- backward on the pipelined persistent kernels (hct_attention_set_bwd3(1))")
lib().hct_attention_set_bwd3(1)
for name, (B, S, H, hd) in {"dec": (256, 513, 16, 48), "enc": (256, 129, 12, 64), "vit": (64, 517, 12, 64)}.items():
    D = H * hd
    qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16(); out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
    do = torch.randn(B, S, D, device=dev).bfloat16(); lse = torch.empty(B, H, S, device=dev)
    dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, S, device=dev); st = stream_ptr(dev)
    call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st)
    b = timeit(lambda: call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(),
                            dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st))
    print(f"{name}: bwd3 {b:.3f} ms ({2 * 4.0 * B * H * S * S * hd / b / 1e9:.0f} TFLOP/s on 2 x forward flops)")
lib().hct_attention_set_bwd3(0)
