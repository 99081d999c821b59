"""Forward attention: four-CTAs-per-SM kernel vs the pipelined persistent kernel (hct_attention_set_fwd2), one process.
    python tools/attn_fwd2_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200._cabi import call, stream_ptr, lib

dev = torch.device("cuda")


def timeit(fn, iters=10):
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


for name, (B, S, H, hd) in {"dec": (256, 513, 16, 48), "enc": (256, 129, 12, 64), "vit": (64, 517, 12, 64), "s512": (256, 512, 16, 48)}.items():
    D = H * hd
    torch.manual_seed(0)
    qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16()
    outs = []
    for fwd2, fold in ((0, 0), (0, 1), (0, 3), (0, 0), (0, 3)):
        lib().hct_attention_set_fwd2(fwd2)
        lib().hct_attention_set_tail_key(fold)
        out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, H, S, device=dev)
        st = stream_ptr(dev)
        f = timeit(lambda: call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st))
        torch.cuda.synchronize()
        outs.append((out, lse))
        print(f"{name} B={B} S={S} H={H} hd={hd} fwd2={fwd2} fold_tail_key={fold}: {f:.3f} ms ({4.0 * B * H * S * S * hd / f / 1e9:.0f} TFLOP/s)", flush=True)
    d = (outs[0][0].float() - outs[2][0].float()).norm() / outs[0][0].float().norm()
    dl = (outs[0][1] - outs[2][1]).abs().max()
    dt = (outs[0][0][:, -1].float() - outs[2][0][:, -1].float()).norm() / outs[0][0][:, -1].float().norm()
    print(f'   last row rel diff {dt.item():.2e}')
    print(f"   out rel diff {d.item():.2e}, lse max abs diff {dl.item():.2e}")
lib().hct_attention_set_fwd2(0); lib().hct_attention_set_tail_key(3)
