"""Per-image sharding across GPUs: every stage of the path is independent per image, so ranks
own contiguous batch slices and no collective sits on the inference data path.  The only communication is
the max-over-ranks of the timed region (and, for training, the weight-gradient all-reduce that lives outside
this path)."""
from __future__ import annotations

import os


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_slice(global_batch: int, rank: int, world: int) -> slice:
    """Contiguous slice of images owned by ``rank``: image b -> rank floor(b*world/global_batch)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    lo = (global_batch * rank) // world
    hi = (global_batch * (rank + 1)) // world
    return slice(lo, hi)


def max_over_ranks(value: float, device=None) -> float:
    """All-reduce(MAX) of a scalar over the default process group (no-op without one)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
