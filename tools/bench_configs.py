"""Secondary configurations of BASELINE.json (configs[2..4]) on one B200: DINO step, fine-tune step, extraction sweep.

    python tools/bench_configs.py [dino|finetune|extract|all]
Prints one JSON line per measurement (not the driver's bench contract -- supporting numbers for profiles/)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import headct_foundation_b200 as H
from headct_foundation_b200 import configs as C
from headct_foundation_b200.optim import FusedAdamW

dev = torch.device("cuda")
PEAK = 1395.3


def timed(fn, warmup=3, steps=8):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def extract():
    torch.manual_seed(0)
    m = H.ViT(**C.VIT_EXTRACT).to(dev).eval()
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
        x = torch.rand(B, 3, 96, 96, 96, device=dev)
        with torch.no_grad():
            ms = timed(lambda: m(x), warmup=2, steps=5)
        tf = 100.921 * B / ms
        print(json.dumps({"config": "extract_feature (ViT-B 3D, S=513, 12 hidden states materialised)", "batch": B,
                          "ms": ms, "volumes_per_s": B / ms * 1e3, "algorithmic_tflops": tf, "frac_of_peak": tf / PEAK}), flush=True)
        if B <= 64:      # launch-bound range: the same forward replayed from a CUDA graph (utils/graphs.py)
            gf = H.GraphedForward(m, x)
            msg = timed(lambda: gf(x), warmup=2, steps=20)
            print(json.dumps({"config": "extract_feature, CUDA-graph replay", "batch": B, "ms": msg, "volumes_per_s": B / msg * 1e3,
                              "frac_of_peak": 100.921 * B / msg / PEAK, "speedup_vs_eager": ms / msg}), flush=True)
            del gf
        del x


def finetune(B=64):
    torch.manual_seed(0)
    m = H.ViT(**C.VIT_DOWNSTREAM).to(dev).train()
    clf = H.LinearClassifier(768, 2).to(dev).train()
    opt = FusedAdamW(list(m.parameters()), lr=1e-4, betas=(0.9, 0.999), weight_decay=0.05)
    opt2 = torch.optim.AdamW(clf.parameters(), lr=1e-2)
    x = torch.rand(B, 3, 96, 96, 96, device=dev)
    y = torch.randint(0, 2, (B,), device=dev)
    ce = torch.nn.CrossEntropyLoss()

    def step():
        opt.zero_grad(set_to_none=True); opt2.zero_grad(set_to_none=True)
        out, _ = m(x)
        loss = ce(clf(out[:, :1, :].squeeze(1)), y)
        loss.backward()
        opt.step(); opt2.step()
    ms = timed(step)
    tf = 298.687 * B / ms
    print(json.dumps({"config": "fine-tune step ViT-B + LinearClassifier + CE (vit_HeadCT_cq500 shape)", "batch": B, "ms": ms,
                      "volumes_per_s": B / ms * 1e3, "algorithmic_tflops": tf, "frac_of_peak": tf / PEAK}), flush=True)


def dino(B=64):
    torch.manual_seed(0)
    student = H.MultiCropWrapper(H.ViT(**C.VIT_DINO), H.DINOHead(**C.DINO_HEAD)).to(dev).train()
    teacher = H.MultiCropWrapper(H.ViT(**C.VIT_DINO), H.DINOHead(**C.DINO_HEAD)).to(dev).train()
    teacher.load_state_dict(student.state_dict())
    for p in teacher.parameters():
        p.requires_grad = False
    crit = H.DINOLoss(**C.DINO_LOSS).to(dev)
    opt = FusedAdamW([p for p in student.parameters() if p.requires_grad], lr=5e-4 * B / 256, betas=(0.9, 0.999),
                     weight_decay=0.04, clip_grad=3.0)
    crops = [torch.rand(B, 3, 96, 96, 96, device=dev) for _ in range(4)]
    losses = []

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.no_grad():
            t = teacher(crops[:2])["dino_output"]
        s = student(crops)["dino_output"]
        loss = crit(s, t, 0)
        loss.backward()
        opt.step()
        H.update_momentum_encoder(student, teacher, 0.999)
        losses.append(loss.detach())
    ms = timed(step, warmup=3, steps=5)
    tf = 1408.9 * B / ms
    print(json.dumps({"config": "DINO step: 4 student crops f+b, 2 teacher crops fwd, 65536-way head, loss, center, EMA teacher, "
                                "per-param clip + AdamW", "batch": B, "ms": ms, "volumes_per_s": B / ms * 1e3,
                      "algorithmic_tflops": tf, "frac_of_peak": tf / PEAK, "first_loss": float(losses[0]),
                      "last_loss": float(losses[-1])}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("extract", "all"):
        extract()
    if what in ("finetune", "all"):
        finetune()
    if what in ("dino", "all"):
        dino()
