"""MAE step at batch 256 on one GPU: eager issue vs one CUDA-graph replay per step (GraphedTrainStep), same process.
    python tools/graph_vs_eager.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import headct_foundation_b200 as H
from headct_foundation_b200.configs import MAE_HEADCT
from headct_foundation_b200.optim import FusedAdamW

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda")
torch.manual_seed(0)
x = torch.rand(B, 3, 96, 96, 96, device=dev)


def timed(fn, warm=3, steps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def make():
    m = H.MaskedAutoencoderViT(**MAE_HEADCT).to(dev).train()
    o = FusedAdamW([p for p in m.parameters() if p.requires_grad], lr=1.5e-4, betas=(0.9, 0.95), weight_decay=0.05, clip_grad=3.0)
    return m, o


m, o = make()
def eager():
    o.zero_grad(set_to_none=True)
    m(x)[0].backward()
    o.step()
for rep in range(2):
    print(f"eager  B={B}: {timed(eager):.3f} ms/step", flush=True)
del m, o
torch.cuda.empty_cache()
m, o = make()
step = H.GraphedTrainStep(m, o, x)
for rep in range(2):
    print(f"graph  B={B}: {timed(lambda: step(x)):.3f} ms/step", flush=True)
