"""Multi-GPU plumbing: replicas only.

Utterances are independent, the graph is read-only: every rank holds a graph
replica and decodes its own shard of the utterances; there is no collective on
the search path.  torch.distributed is used for three things only: the barrier
around the timed region, the MAX-reduction of the per-rank time, and gathering
the (small) results.  The same functions run on gloo (CPU tests) and nccl.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [begin, end) of `n_items` for `rank`; sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_by_length(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy length-balanced assignment (longest first) of utterances to ranks."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(lengths[i])
    for r in range(world_size):
        out[r].sort()
    return out


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (the per-rank elapsed time)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_objects(obj, dst: int = 0):
    """Gathers one picklable object per rank on `dst` (results are a few KB per utterance)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out
