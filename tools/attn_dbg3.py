"""Event timeline of CTA 0 of the pipelined persistent attention-backward kernel (dK/dV pass), clock64 stamps written
through hct_attention_trace3.  python tools/attn_dbg3.py [dec|enc]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from headct_foundation_b200._cabi import call, stream_ptr, lib
dev = torch.device("cuda")
B, S, H, hd = (256, 129, 12, 64) if (len(sys.argv) > 1 and sys.argv[1] == "enc") else (256, 513, 16, 48)
D = H * hd
qkv = torch.randn(B, S, 3 * D, device=dev).bfloat16(); out = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
do = torch.randn(B, S, D, device=dev).bfloat16(); lse = torch.empty(B, H, S, device=dev)
dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, S, device=dev); st = stream_ptr(dev)
call("hct_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, H, hd, st)
lib().hct_attention_set_bwd3(1)
def bwd():
    call("hct_attention_bwd", qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), B, S, H, hd, st)
for _ in range(2): bwd()
torch.cuda.synchronize()
tr = torch.zeros(4 * 64 * 16, dtype=torch.int64, device=dev)
lib().hct_attention_trace3(tr.data_ptr())
bwd(); torch.cuda.synchronize()
lib().hct_attention_trace3(None)
t = tr.cpu().view(4, 64, 16)
t0 = int(t[t > 0].min())
names = {0: ["issue: start", "tiles+buffer ready", "S/dP issued"], 1: ["acc: wait p_full", "got", "acc_empty ok", "acc issued"],
         2: ["top", "stats stored", "pre-wait", "s_full", "computed", "arrived", "drained", "nxt pos", "drain: top", "done ok", "acc_empty arrived"], 3: ["top", "stats stored", "pre-wait", "s_full", "computed", "arrived", "drained", "nxt pos", "drain: top", "done ok", "acc_empty arrived"]}
for role, rn in enumerate(["mma S/dP issue", "mma accumulate", "softmax group 0 (warp 0)", "softmax group 1 (warp 8)"]):
    print(f"-- {rn}: " + " | ".join(names[role]))
    for i in range(30):
        if (t[role, i] > 0).any():
            print(f"   blk {i:2d}: " + " ".join(f"{int(v) - t0:7d}" if v > 0 else "      -" for v in t[role, i, :len(names[role])]))

