"""One forward + backward of the tcgen05 attention at the decoder shape (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200._cabi import call, stream_ptr
dev = torch.device("cuda")
B, S, H, hd = (int(os.environ.get("PB", 64)), 513, 16, 48)
D = H * hd
qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16(); out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
do = torch.randn(B, S, D, device=dev).bfloat16(); lse = torch.empty(B, H, S, device=dev)
dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, S, device=dev); st = stream_ptr(dev)
for _ in range(3):
    call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st)
    call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st)
torch.cuda.synchronize()
print("ok")
