"""Event timeline of one CTA of the dK/dV attention-backward kernel (clock64 stamps written through hct_attention_trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200._cabi import call, stream_ptr, lib
dev = torch.device("cuda")
B, S, H, hd = 256, 513, 16, 48
D = H * hd
qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16(); out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
do = torch.randn(B, S, D, device=dev).bfloat16(); lse = torch.empty(B, H, S, device=dev)
dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, S, device=dev); st = stream_ptr(dev)
call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st)
def bwd():
    call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st)
for flags in [0]:
    for _ in range(2): bwd()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): bwd()
    b.record(); torch.cuda.synchronize()
    print(f"dbg={flags:2d}  bwd (delta + dkdv + dq) {a.elapsed_time(b) / 5:.3f} ms")
    tr = torch.zeros(768, dtype=torch.int64, device=dev)
    lib().hct_attention_trace(tr.data_ptr())
    bwd(); torch.cuda.synchronize()
    lib().hct_attention_trace(None)
    t = tr.cpu().view(3, 32, 8)
    t0 = int(t[t > 0].min())
    names = {0: ["wait qdo_empty", "got", "tma issued"],
             1: ["sdp: wait qdo_full", "got", "S/dP issued+commit", "wait p_full", "got", "dV/dK issued+commit"],
             2: ["wait s_full", "got", "computed", "stored+arrived p_full"]}
    for role, rn in enumerate(["producer", "mma", "softmax w2"]):
        print(f"-- {rn}: " + " | ".join(names[role]))
        for i in range(9):
            print(f"   blk {i}: " + " ".join(f"{int(v) - t0:7d}" if v > 0 else "      -" for v in t[role, i, :len(names[role])]))
