"""Baseline sharding across the GPUs of one box.

Every statistic of the path is per P x P tile and tiles never straddle waterfalls, so the
baseline axis is split contiguously over ranks with no data-path exchange; each rank is
semantically one `Preprocessor` over its slice (the reference's own unit of parallelism,
synthetic_generator.py:55-107).  The only collective is a sum of the metric counts.
"""
from __future__ import annotations

import torch


def baseline_shard(n_baselines: int, world: int, rank: int) -> slice:
    """Contiguous, balanced partition: the first `n % world` ranks get one extra baseline
    (351 over 8 -> 44 x 7 + 43)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_baselines, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def allreduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM of an int64 count tensor ({TP, FP, FN}, moments...) over the process group
    (NCCL for CUDA tensors, gloo for CPU tensors in the tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
