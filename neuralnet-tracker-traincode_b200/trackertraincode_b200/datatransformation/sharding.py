"""Per-sample sharding of the augmentation across ranks: one process per GPU, each augments its own slice
of the global batch, no collective on the data path.  The only cross-rank facts are *which* samples a rank owns and the
global id of each sample in the Philox noise stream (so a sample's noise does not depend on the world size)."""
from __future__ import annotations

import os
from typing import Tuple


def rank_world() -> Tuple[int, int]:
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun environment, else (0, 1)."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of n samples owned by `rank`; sizes differ by at most one, every sample has one owner."""
    assert 0 <= rank < world
    return (rank * n) // world, ((rank + 1) * n) // world


def global_sample_offset(step: int, rank: int, world: int, local_batch: int) -> int:
    """Id of this rank's first sample at `step` in the global sample stream: step * global_batch + rank * local_batch."""
    return (step * world + rank) * local_batch


def max_over_ranks(seconds: float, device=None) -> float:
    """The job's time for a step is the slowest rank's (all_reduce MAX); identity when not distributed."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
