"""Host-side helpers for the data-parallel path (one process per GPU, torch.distributed).

The hot path shards by batch only: the single exchange per step is the gradient all-reduce that torch
DDP performs on the buckets our autograd nodes fill, plus DINO's 256 KiB center all-reduce.  These helpers hold the
small amount of rank arithmetic around that, and are what the world_size-2 gloo tests exercise on CPU.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

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


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced slice of n_items for this rank (first n % world ranks get one extra)."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


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


def mean_gradients_(params: Iterable[torch.Tensor], bucket_bytes: int = 64 << 20) -> int:
    """Bucketed all-reduce(mean) of .grad over ranks -- the semantic of DDP's reducer (main_pretrain_mae.py:139),
    for callers that drive backward without the DDP wrapper.  Returns the number of buckets reduced."""
    if not is_dist() or dist.get_world_size() == 1:
        return 0
    ws = dist.get_world_size()
    bucket: List[torch.Tensor] = []
    size = n_buckets = 0

    def flush():
        nonlocal bucket, size, n_buckets
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat)
        flat.div_(ws)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0
        n_buckets += 1

    for p in params:
        if p.grad is None:
            continue
        bucket.append(p.grad)
        size += p.grad.numel() * p.grad.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
    return n_buckets
