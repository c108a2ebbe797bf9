"""Multi-GPU plumbing for the paths that shard without a data-path collective: the sampler and
the mel-STFT / FGLA codec are independent per stereo item, so N GPUs run N replicas over a partition of the items
(seeds / batch rows).  torch.distributed is only used to agree on the partition and to time the job as the maximum
over ranks."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of range(n_items): the first n_items % world ranks get one extra item."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_seeds(base_seed: int, n_items: int, rank: int, world: int) -> List[int]:
    """Per-item sampler seeds owned by this rank (item i of the global batch always gets base_seed + i,
    independent of the number of ranks, so results do not depend on the sharding)."""
    lo, hi = shard_range(n_items, rank, world)
    return [base_seed + i for i in range(lo, hi)]


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Job time = slowest rank (all-reduce MAX; identity without a process group)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count: int, dist=None, device=None) -> Sequence[int]:
    """Items processed per rank (all-gather), for the whole-job throughput numerator."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [int(count)]
    t = torch.tensor([int(count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]
