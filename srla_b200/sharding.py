"""Multi-GPU sharding of the encode path.

Blocks and files are independent, so N GPUs split a batch by contiguous ranges of whole streams:
rank r encodes streams [lo, hi) and writes its own outputs.  No data-path collective exists; only
sizes / timings are ever exchanged (bench.py uses torch.distributed for the barrier and the max).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one) range of `num_items` for `rank`."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(num_items, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi
