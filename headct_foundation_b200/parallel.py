"""Host-side helpers for the data-parallel path (one process per GPU, torch.distributed).

The hot path shards by batch only: the single exchange per step is the gradient all-reduce that torch
DDP performs on the buckets our autograd nodes fill, plus DINO's 256 KiB center all-reduce.  These helpers hold the
small amount of rank arithmetic around that: `functional.center_update` (DINO center) and `bench.py` (seeds, learning-rate
rule, max-over-ranks timing) call them, and the world_size-2 gloo tests exercise them on CPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized()


def world() -> Tuple[int, int]:
    return (dist.get_rank(), dist.get_world_size()) if is_dist() else (0, 1)


def rank_seed(base_seed: int, rank: int) -> int:
    """seed = SEED + rank (main_pretrain_mae.py:213) so every rank draws different mask noise."""
    return int(base_seed) + int(rank)


def scaled_lr(base_lr: float, batch_per_gpu: int, world_size: int) -> float:
    """lr = BASE_LR * (B * world) / 256 (main_pretrain_mae.py:149-152)."""
    return base_lr * batch_per_gpu * world_size / 256.0


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_dist() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t


def max_over_ranks(values: Sequence[float], device=None) -> List[float]:
    """Timing rule: a multi-GPU number is the max over ranks."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if is_dist() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def broadcast_params_(module: torch.nn.Module, src: int = 0) -> None:
    """Every rank starts from rank `src`'s parameters and buffers (what DistributedDataParallel does at construction);
    for the data-parallel paths that do not go through DDP (the CUDA-graph step)."""
    if not (is_dist() and dist.get_world_size() > 1):
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src)


def allreduce_mean_grads_(params: Sequence[torch.nn.Parameter], bucket_bytes: int = 64 << 20) -> int:
    """In place: every `.grad` becomes its mean over the ranks -- the one exchange step of the data-parallel path
    (main_pretrain_mae.py:139's DDP does it from autograd hooks; this is the form that can sit INSIDE a CUDA-graph capture,
    between backward and the optimizer update).  On NCCL the gradients go out in coalesced groups of ~`bucket_bytes` (one
    launch per group, reduced where they lie -- no flattening copy) with the averaging done by the collective; other
    backends (gloo in the CPU tests) sum per tensor and divide.  Returns the number of collective groups issued."""
    if not (is_dist() and dist.get_world_size() > 1):
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    world_size = dist.get_world_size()
    if dist.get_backend() != "nccl":
        for g in grads:
            dist.all_reduce(g)
            g.div_(world_size)
        return len(grads)
    groups: List[List[torch.Tensor]] = [[]]
    size = 0
    for g in grads:
        nbytes = g.numel() * g.element_size()
        if groups[-1] and size + nbytes > bucket_bytes:
            groups.append([])
            size = 0
        groups[-1].append(g)
        size += nbytes
    from torch.distributed.distributed_c10d import _coalescing_manager
    for grp in groups:
        with _coalescing_manager(device=grp[0].device, async_ops=False):
            for g in grp:
                dist.all_reduce(g, op=dist.ReduceOp.AVG)
    return len(groups)
