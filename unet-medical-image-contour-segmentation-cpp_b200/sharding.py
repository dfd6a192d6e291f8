"""Slice sharding across ranks: slices are independent, so each rank takes a
contiguous block and no data-path collective exists.  torch.distributed only carries the barrier and the
max-over-ranks of the timing; both work with nccl (GPU) and gloo (CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_slices: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: block sizes differ by at most one, earlier ranks get the
    extra slice (cfg3: 256 slices over 2/4/8 GPUs -> 128/64/32 each)."""
    if not (0 <= rank < world) or n_slices < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_slices, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a scalar over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
