"""Experiment: how much of the MAE step is launch / inter-kernel overhead?  Times forward+backward eagerly and as a
replayed CUDA graph (torch.cuda.graph) at B = 256.  Not part of the product path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import headct_foundation_b200 as H
from headct_foundation_b200.configs import MAE_HEADCT

dev = torch.device("cuda")
B = int(os.environ.get("PB", 256))
torch.manual_seed(0)
model = H.MaskedAutoencoderViT(**MAE_HEADCT).to(dev).train()
x = torch.rand(B, 3, 96, 96, 96, device=dev)


def fb():
    for p in model.parameters():
        p.grad = None
    loss, _, _ = model(x)
    loss.backward()
    return loss


def timeit(fn, n=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


print(f"eager fwd+bwd: {timeit(fb):.2f} ms")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(2):
        fb()
torch.cuda.current_stream().wait_stream(s)
try:
    with torch.cuda.graph(g):
        loss = fb()
    print(f"graph fwd+bwd: {timeit(g.replay):.2f} ms   (loss {loss.item():.4f})")
except Exception as e:
    print("graph capture failed:", repr(e)[:300])
