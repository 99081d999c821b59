"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic and of bench.py's multi-rank contract."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from headct_foundation_b200 import parallel as P
    res = {}
    assert P.world() == (rank, world)
    res["seed"] = P.rank_seed(42, rank)
    res["max"] = P.max_over_ranks([10.0 + rank, 5.0 - rank])
    # DINO center: sum of per-rank column sums, divided by (rows * world)  (losses.py:96-98)
    bc = torch.full((1, 8), float(rank + 1))
    P.allreduce_sum_(bc)
    res["center"] = (bc / (4 * world)).tolist()
    # gradient mean over ranks through torch DDP on gloo (what main_pretrain_mae.py:139 / bench.py do on NCCL)
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    ddp = torch.nn.parallel.DistributedDataParallel(lin, bucket_cap_mb=1)
    x = torch.full((2, 5), float(rank + 1))
    ddp(x).sum().backward()
    res["wgrad"] = float(lin.weight.grad[0, 0])                  # d/dw sum = sum of inputs = 2 (rank + 1); mean over ranks = 3
    res["lr"] = P.scaled_lr(1.5e-4, 256, world)
    # the exchange step in its capturable form: every .grad becomes the mean over ranks; parameters start from rank 0's
    torch.manual_seed(rank)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    P.broadcast_params_(net)
    res["w0"] = float(net[0].weight.sum())
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
    net[1].bias.grad = None                                       # a frozen tensor takes no part
    n_groups = P.allreduce_mean_grads_(list(net.parameters()))
    res["mean_grads"] = [None if p.grad is None else float(p.grad.flatten()[0]) for p in net.parameters()]
    res["groups"] = n_groups
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_helpers():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29611, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert (r0["seed"], r1["seed"]) == (42, 43)
    assert r0["max"] == [11.0, 5.0] == r1["max"]
    assert r0["center"] == r1["center"] == [[3.0 / 8] * 8]
    assert r0["wgrad"] == r1["wgrad"] == 3.0                      # mean over ranks of 2 (rank + 1)
    assert r0["lr"] == r1["lr"] == pytest.approx(3e-4)
    assert r0["w0"] == r1["w0"]                                   # broadcast_params_: both ranks hold rank 0's weights
    assert r0["mean_grads"] == r1["mean_grads"] == [1.5, 3.0, 4.5, None] and r0["groups"] == 3


def test_scaled_lr_rule():
    from headct_foundation_b200 import parallel as P
    assert P.scaled_lr(1.5e-4, 256, 8) == pytest.approx(1.5e-4 * 8)
    assert P.world() == (0, 1) and P.max_over_ranks([1.0, 2.0]) == [1.0, 2.0]      # single process: identity


@pytest.mark.timeout(600)
def test_bench_reference_arm_under_torchrun_two_ranks():
    """`bench.py --impl reference` under torchrun: rank 0 alone prints ONE JSON line, rank 1 exits 0 without work."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "3", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "volumes/s" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
