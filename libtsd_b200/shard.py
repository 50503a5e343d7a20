"""Multi-GPU plumbing: channels are independent units, so a batch is split into contiguous
channel ranges, one per rank (one process per GPU), with NO collective on the data path.  The only
collectives are the optional final gather of the output shards (to the consumer) and the reduction of
timings.  Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def channel_shard(nchan: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [start, start+count) of channels owned by `rank`; remainders go to the first ranks."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(nchan, world)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def gather_channels(y_local, nchan: int, group=None):
    """All-gather the per-rank output shards [count_r, n] into [nchan, n] (same n on every rank).
    Uneven shards are padded to the largest one for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = y_local.shape[1]
    counts = [channel_shard(nchan, r, world)[1] for r in range(world)]
    cmax = max(counts)
    pad = torch.zeros((cmax, n), dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def max_over_ranks(value: float, device="cpu", group=None) -> float:
    """Timing reduction used by bench.py: the slowest rank defines the step time."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
