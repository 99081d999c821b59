"""One backward launch per mode at the decoder / encoder / DINO shapes (run under ncu --metrics gpu__time_duration.sum)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200._cabi import call, stream_ptr, lib
dev = torch.device("cuda")
for name, (B, S, H, hd) in {"dec": (256, 513, 16, 48), "enc": (256, 129, 12, 64), "vit": (64, 517, 12, 64)}.items():
    D = H * hd
    qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16(); out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
    do = torch.randn(B, S, D, device=dev).bfloat16(); lse = torch.empty(B, H, S, device=dev)
    dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, S, device=dev); st = stream_ptr(dev)
    call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st)
    for mode in (2, 3):
        lib().hct_attention_set_tcgen05(mode)
        for _ in range(2):
            call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st)
    torch.cuda.synchronize()
