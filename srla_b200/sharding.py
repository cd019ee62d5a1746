"""Multi-GPU sharding of the encode path and the host-side placement that goes with it.

Blocks and files are independent, so N GPUs split a batch by contiguous ranges of whole streams: rank r encodes
streams [lo, hi) and writes its own outputs.  No data-path collective exists; only sizes / timings are ever exchanged
(bench.py uses torch.distributed for the barrier and the max).  What limits the end-to-end rate of N ranks on one host is
the host side -- PCIe ingest and the host threads that feed it -- so every rank keeps its page-locked staging memory and
its feeder threads on the CPUs and the memory node next to its GPU (`pin_rank_to_gpu_cpus`), and the ranks that share a
node split its CPUs instead of piling onto the same cores.
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple


def shard_range(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most one) range of `num_items` for `rank`."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(num_items, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11] (the format of sysfs cpulist files)"""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_locality(pci_bus_id: str, sysfs_root: str = "/sys/bus/pci/devices") -> Tuple[Optional[int], List[int]]:
    """(NUMA node, CPUs next to it) of the PCI device `dddd:bb:dd.f`, read from sysfs; (None, []) when sysfs does not say."""
    base = os.path.join(sysfs_root, pci_bus_id.lower())
    node: Optional[int] = None
    cpus: List[int] = []
    try:
        with open(os.path.join(base, "numa_node")) as f:
            v = int(f.read().strip())
            node = v if v >= 0 else None
    except (OSError, ValueError):
        pass
    try:
        with open(os.path.join(base, "local_cpulist")) as f:
            cpus = parse_cpulist(f.read())
    except (OSError, ValueError):
        pass
    return node, cpus


def split_cpus(cpus: List[int], sharers: int, index: int) -> List[int]:
    """the `index`-th of `sharers` contiguous, balanced parts of `cpus` (every part non-empty while cpus last)"""
    if sharers <= 1 or not cpus:
        return list(cpus)
    lo, hi = shard_range(len(cpus), index, sharers)
    if hi <= lo:                                   # more sharers than CPUs: wrap around
        return [cpus[index % len(cpus)]]
    return cpus[lo:hi]


def pin_rank_to_gpu_cpus(pci_bus_ids: List[str], local_rank: int, local_world: int) -> dict:
    """Restrict the calling process (and every thread it starts later: the library's feeder pool, the CUDA driver threads)
    to its share of the CPUs next to GPU `local_rank`, so that page-locked staging memory is first touched on that node
    and the feeder threads of different ranks do not compete for cores.  `pci_bus_ids[i]` = PCI address of local GPU i.
    Returns what was done (for the bench record)."""
    allowed = sorted(os.sched_getaffinity(0))
    node, cpus = gpu_locality(pci_bus_ids[local_rank])
    cpus = [c for c in cpus if c in allowed] or allowed
    # ranks whose GPU reports the same CPU set share it
    same = [r for r in range(local_world) if (gpu_locality(pci_bus_ids[r])[1] or allowed) == (gpu_locality(pci_bus_ids[local_rank])[1] or allowed)]
    mine = split_cpus(cpus, len(same), same.index(local_rank)) if local_rank in same else cpus
    info = {"numa_node": node, "gpu_local_cpus": len(cpus), "ranks_sharing_them": len(same), "cpus_of_this_rank": len(mine)}
    try:
        os.sched_setaffinity(0, set(mine))
        info["pinned"] = True
    except OSError:
        info["pinned"] = False
    return info
