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
