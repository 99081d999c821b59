"""A/B of the attention softmax arithmetic: exponential pairs on the FMA pipe (hct_attention_set_poly) 0 / 2 / 3 / 4 of 8,
forward and pipelined backward, at the MAE / DINO shapes; accuracy against an fp32 torch attention on a small batch.
    python tools/attn_poly_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200._cabi import call, stream_ptr, lib

dev = torch.device("cuda")


def timeit(fn, iters=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def ref(qkv, do, H, hd):
    B, S, _ = qkv.shape
    q, k, v = [t.reshape(B, S, H, hd).transpose(1, 2).float().requires_grad_() for t in qkv.chunk(3, dim=-1)]
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    o.backward(do.reshape(B, S, H, hd).transpose(1, 2).float())
    dq = torch.cat([t.grad.transpose(1, 2).reshape(B, S, H * hd) for t in (q, k, v)], dim=-1)
    return o.transpose(1, 2).reshape(B, S, H * hd), dq


shapes = {"dec": (256, 513, 16, 48), "enc": (256, 129, 12, 64), "vit": (64, 517, 12, 64)}
for name, (B, S, H, hd) in shapes.items():
    D = H * hd
    torch.manual_seed(0)
    qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16()
    out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
    do = torch.randn(B, S, D, device=dev).bfloat16()
    lse = torch.empty(B, H, S, device=dev)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, S, device=dev)
    st = stream_ptr(dev)
    with torch.backends.cuda.sdp_kernel(enable_flash=False, enable_mem_efficient=False, enable_math=True):
        o_ref, d_ref = ref(qkv[:4], do[:4], H, hd)
    lib().hct_attention_set_bwd3(0)      # the untouched two-kernel backward: the box-speed reference of this run
    b_old = timeit(lambda: call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st))
    lib().hct_attention_set_bwd3(1)
    print(f"{name}: two-kernel backward {b_old:.3f} ms (reference)")
    for poly, tma in ((-1, 1), (0, 1), (3, 1), (-1, 1), (0, 1)):
        lib().hct_attention_set_poly(poly, min(poly, 0))
        lib().hct_attention_set_bwd3_drain(tma)
        f = timeit(lambda: call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st))
        b = timeit(lambda: call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(),
                                dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st))
        torch.cuda.synchronize()
        eo = ((out[:4].float() - o_ref).norm() / o_ref.norm()).item()
        ed = ((dqkv[:4].float() - d_ref).norm() / d_ref.norm()).item()
        print(f"{name} B={B} S={S} H={H} hd={hd} fwd poly {poly}/8, bwd {min(poly, 0)}: fwd {f:.3f} ms  bwd {b:.3f} ms ({b / b_old:.3f} of the reference)   rel L2 err out {eo:.2e} dqkv {ed:.2e}", flush=True)
lib().hct_attention_set_poly(-1, 0); lib().hct_attention_set_bwd3_drain(1)
