"""Multi-GPU plumbing of the spatial/mixer path: one process per GPU, sources sharded over the ranks,
one sum all-reduce of the small output tile per callback. The reference has no
counterpart (it is single-process); the collective goes through torch.distributed (NCCL on GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_sources(n_sources: int, rank: int, world: int) -> np.ndarray:
    """Round-robin: source s belongs to rank s mod world (keeps the ranks balanced as sources finish)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return np.arange(rank, n_sources, world, dtype=np.int64)


def allreduce_tile(tile, group=None):
    """Sums the per-rank partial tiles in place; `tile` is a torch tensor on the rank's device."""
    import torch.distributed as dist

    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tile, op=dist.ReduceOp.SUM, group=group)
    return tile
