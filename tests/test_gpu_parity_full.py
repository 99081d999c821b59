"""Parity at the benchmarked shape: the full mae_HeadCT.yaml model (12 + 8 blocks, 96^3 x 3 volumes) at batch 16
against the oracle executed in fp32 ON THE GPU (plain torch ops, TF32 off) -- loss, mask indices and per-parameter
gradients -- plus the 2-rank check that DDP-reduced gradients through our autograd nodes equal the mean of the two
ranks' stand-alone backward passes (main_pretrain_mae.py:139).  Round 1 compared the B = 256 path only with itself
(tests/test_gpu_fullsize.py) and with the oracle at B = 2."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def _oracle_mae_on_cuda(sd, x, noise, cfg):
    from oracle import headct_oracle as O
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        out = O.mae_forward(sdg, x, noise, patch=(cfg["patch_size"],) * 3, mask_ratio=cfg["mask_ratio"],
                            enc_heads=cfg["encoder_num_heads"], dec_heads=cfg["decoder_num_heads"],
                            norm_pix=cfg["norm_pix_loss"])
        out["loss"].backward()
        grads = {k: v.grad for k, v in sdg.items() if v.grad is not None}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    return out, grads


@pytest.mark.parametrize("batch", [16])
def test_mae_full_config_batch16_against_oracle_on_cuda(cuda, batch):
    import headct_foundation_b200 as H
    from oracle import synth
    cfg = synth.MAE_FULL
    sd = synth.to_device(synth.mae_state_dict(cfg, seed=21), cuda)
    model = H.MaskedAutoencoderViT(**cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda).train()
    x = synth.volume(batch, cfg["in_chans"], cfg["input_size"], 22).to(cuda)
    noise = synth.noise(batch, (cfg["input_size"] // cfg["patch_size"]) ** 3, seed=23).to(cuda)
    model.noise_override = noise
    loss, _, _ = model(x)
    loss.backward()
    latent, mask, ids_restore = model.forward_encoder(x)

    ref, rgrads = _oracle_mae_on_cuda(sd, x, noise, cfg)
    assert torch.equal(ids_restore, ref["ids_restore"])                       # bit exact
    assert torch.equal(mask, ref["mask"])
    rel = abs(loss.item() - ref["loss"].item()) / ref["loss"].item()
    assert rel < 1e-2, (loss.item(), ref["loss"].item())
    assert _cos(latent, ref["latent"]) > 0.9995
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    frozen = {k for k, p in model.named_parameters() if not p.requires_grad}          # decoder_pos_embed (mae.py:92)
    assert frozen == {"decoder_pos_embed"}
    assert set(grads) == set(rgrads) - frozen, set(grads) ^ set(rgrads)
    bad = {}
    for k, g in grads.items():
        r = rgrads[k]
        gn, rn = g.norm().item(), r.norm().item()
        c = _cos(g, r)
        # gradient norms within 5 % and directions within cos 0.99 (bf16 operands, fp32 accumulation), as at B = 2
        if abs(gn - rn) > 0.05 * rn + 1e-7 or c < 0.99:
            bad[k] = (gn, rn, c)
    assert not bad, bad


_DDP_WORKER = r'''
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["HCT_ROOT"])
import headct_foundation_b200 as H
from oracle import synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
cfg = synth.MAE_SMALL
sd = synth.mae_state_dict(cfg, seed=31)
L = (cfg["input_size"] // cfg["patch_size"]) ** 3
xs = [synth.volume(4, cfg["in_chans"], cfg["input_size"], 40 + r).to(dev) for r in range(world)]
ns = [synth.noise(4, L, seed=50 + r).to(dev) for r in range(world)]

def fresh():
    m = H.MaskedAutoencoderViT(**cfg); m.load_state_dict(sd, strict=True); return m.to(dev).train()

# stand-alone backward of every rank's batch, averaged on this rank (no communication)
want = None
for r in range(world):
    m = fresh(); m.noise_override = ns[r]
    m(xs[r])[0].backward()
    g = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    want = g if want is None else {k: want[k] + g[k] for k in g}
want = {k: v / world for k, v in want.items()}
# the same through torch DDP (bucketed NCCL all-reduce fired by our autograd nodes' gradient hooks)
m = fresh(); m.noise_override = ns[rank]
ddp = torch.nn.parallel.DistributedDataParallel(m, device_ids=[dev.index], bucket_cap_mb=1, gradient_as_bucket_view=True)
ddp(xs[rank])[0].backward()
torch.cuda.synchronize()
worst = 0.0
for k, p in m.named_parameters():
    if p.grad is None:
        continue
    d = (p.grad.double() - want[k].double()).norm() / want[k].double().norm().clamp_min(1e-30)
    worst = max(worst, d.item())
t = torch.tensor([worst], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"worst_rel": t.item(), "world": world, "params": len(want)}))
dist.destroy_process_group()
'''


def test_ddp_reduced_gradients_equal_mean_of_rank_gradients(cuda, tmp_path):
    """2 ranks, NCCL: gradients after DDP's all-reduce(mean) == mean over ranks of the stand-alone gradients.
    Both sides run the same kernels; the only differences are fp32 summation order (split-K wgrad atomics, NCCL ring)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    worker = tmp_path / "ddp_worker.py"
    worker.write_text(_DDP_WORKER)
    env = dict(os.environ, HCT_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(worker)], capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["params"] > 50
    assert out["worst_rel"] < 1e-4, out


_GRAPH_DDP_WORKER = r'''
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["HCT_ROOT"])
import headct_foundation_b200 as H
from headct_foundation_b200 import parallel
from headct_foundation_b200.optim import FusedAdamW
from oracle import synth
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
cfg = synth.MAE_SMALL
sd = synth.mae_state_dict(cfg, seed=31)
L = (cfg["input_size"] // cfg["patch_size"]) ** 3
x = synth.volume(4, cfg["in_chans"], cfg["input_size"], 40 + rank).to(dev)
noise = synth.noise(4, L, seed=50 + rank).to(dev)

def fresh(seed_shift):
    m = H.MaskedAutoencoderViT(**cfg)
    sdr = {k: (v + 0.01 * seed_shift if v.is_floating_point() else v) for k, v in sd.items()}   # ranks start apart ...
    m.load_state_dict(sdr, strict=True); m.noise_override = noise
    return m.to(dev).train()

def opt_of(m):
    return FusedAdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05, clip_grad=3.0)

# eager: backward, all-reduce(mean) of the gradients, update -- three steps
me = fresh(rank); parallel.broadcast_params_(me); oe = opt_of(me)          # ... and are put on rank 0's parameters
start = {k: p.detach().clone() for k, p in me.named_parameters()}
for _ in range(3):
    oe.zero_grad(set_to_none=True)
    me(x)[0].backward()
    parallel.allreduce_mean_grads_(list(me.parameters()))
    oe.step()
# the same three steps replayed from one CUDA graph per rank with the all-reduce captured inside
mg = fresh(rank); og = opt_of(mg)
step = H.GraphedTrainStep(mg, og, x)                                        # broadcasts from rank 0 itself
losses = [float(step(x).item()) for _ in range(3)]
torch.cuda.synchronize()
num = den_e = den_g = 0.0
spread = 0.0
for (k, pe), (_, pg) in zip(me.named_parameters(), mg.named_parameters()):
    ue, ug = (pe.detach() - start[k]).double().flatten(), (pg.detach() - start[k]).double().flatten()
    num += float(ue @ ug); den_e += float(ue @ ue); den_g += float(ug @ ug)
    lo, hi = pg.detach().clone(), pg.detach().clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    spread = max(spread, float((hi - lo).abs().max()))
if rank == 0:
    print(json.dumps({"cos_update": num / (den_e * den_g) ** 0.5, "rank_spread": spread, "losses": losses, "world": world}), flush=True)
del step                       # the graph holds NCCL kernels: it has to go before the communicator does
torch.cuda.synchronize()
dist.destroy_process_group()
'''


def test_graphed_step_with_captured_allreduce(cuda, tmp_path):
    """2 ranks, NCCL: GraphedTrainStep with the gradient all-reduce(mean) captured inside the graph
    (parallel.allreduce_mean_grads_) follows the eager backward -> all-reduce -> update trajectory, and the ranks'
    parameters stay bit-identical (same averaged gradients, deterministic update)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    worker = tmp_path / "graph_ddp_worker.py"
    worker.write_text(_GRAPH_DDP_WORKER)
    env = dict(os.environ, HCT_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29534", str(worker)], capture_output=True, text=True, env=env,
                       timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["world"] == 2
    assert out["rank_spread"] == 0.0, out                   # every rank holds the same parameters after three steps
    assert out["cos_update"] > 0.95, out                    # wgrad split-K atomics: runs differ in the last bits
    assert all(v == v and abs(v) < 1e3 for v in out["losses"]), out


def test_block_output_with_a_second_consumer(cuda):
    """A block's fp32 output gradient has a bf16 twin that rides to the next backward node in a side table
    (functional.put_bf16_shadow).  When the block output ALSO feeds a second loss term (ViT returns every hidden state,
    vit.py:161-166), autograd sums two incoming gradients: the twin is then stale and must not be served.  Checked
    against the oracle's blocks in fp32 for both orders of graph construction, and with an in-place gradient hook."""
    import headct_foundation_b200 as H
    from headct_foundation_b200 import functional as HF
    from oracle import headct_oracle as O, synth
    cfg = synth.VIT_SMALL
    sd = synth.vit_state_dict(cfg, seed=9)
    dim, heads = cfg["hidden_size"], cfg["num_heads"]
    blocks = []
    for i in range(2):
        b = H.AttentionBlock(dim, cfg["mlp_dim"], heads, qkv_bias=True)
        b.load_state_dict({k[len(f"blocks.{i}."):]: v for k, v in sd.items() if k.startswith(f"blocks.{i}.")}, strict=True)
        blocks.append(b.to(cuda).train())
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 37, dim, generator=g)
    w_mid = torch.randn(2, 37, dim, generator=g)
    w_out = torch.randn(2, 37, dim, generator=g)

    # oracle
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("blocks.")}
    xr = x.clone().requires_grad_(True)
    y1 = O.attention_block(xr, sdr, "blocks.0.", heads)
    y2 = O.attention_block(y1, sdr, "blocks.1.", heads)
    ((y2 * w_out).sum() + (y1 * w_mid).sum()).backward()

    def run(second_first: bool, hook: bool):
        for b in blocks:
            b.zero_grad()
        xc = x.to(cuda).requires_grad_(True)
        y1c, _ = blocks[0](xc)
        if hook:
            y1c.register_hook(lambda gr: gr.mul_(1.0))         # in-place edit of the gradient: bumps its version
        if second_first:
            side = (y1c * w_mid.to(cuda)).sum()
            y2c, _ = blocks[1](y1c)
        else:
            y2c, _ = blocks[1](y1c)
            side = (y1c * w_mid.to(cuda)).sum()
        ((y2c * w_out.to(cuda)).sum() + side).backward()
        assert _cos(xc.grad.cpu(), xr.grad) > 0.999, (second_first, hook)
        for name in ("mlp.linear2.weight", "attn.qkv.weight", "att_norm.weight"):
            got = dict(blocks[0].named_parameters())[name].grad.cpu()
            assert _cos(got, sdr["blocks.0." + name].grad) > 0.998, (name, second_first, hook)

    for second_first in (False, True):
        for hook in (False, True):
            run(second_first, hook)
    # and the plain chain still takes the fast path
    HF.shadow_stats(reset=True)
    xc = x.to(cuda).requires_grad_(True)
    y, _ = blocks[1](blocks[0](xc)[0])
    (y * w_out.to(cuda)).sum().backward()
    st = HF.shadow_stats()
    assert st["hit"] >= 1, st
