"""Multi-GPU plumbing for the hook-chain path: payloads are independent units, so a batch is
partitioned across ranks (size-balanced) and every rank scans its own shard; the ONLY exchange is
one all-gather of the per-unit verdict bitmaps (24-ish bytes per payload; payload bytes never cross
NVLink).  torch.distributed is plumbing here: NCCL on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def partition_units(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Greedy size-balanced assignment (longest first) of unit indices to ranks; each rank's list is
    returned in ascending unit order so verdicts can be scattered back deterministically."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    loads = [0] * world
    parts: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += sizes[i] + 1
    for p in parts:
        p.sort()
    return parts


def gather_verdicts(local: "np.ndarray | object", counts: Sequence[int], words: int, group=None):
    """All-gather variable-length verdict arrays (AllGatherv emulated by padding to the largest shard,
    one collective).  `local` is a torch tensor (int64, n_local * words) on the rank's device.
    Returns the list of per-rank tensors trimmed to their true lengths."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    mx = max(counts) * words
    buf = torch.zeros(mx, dtype=torch.int64, device=local.device)
    buf[: local.numel()] = local
    out = torch.empty(world * mx, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return [out[r * mx: r * mx + counts[r] * words] for r in range(world)]


def scatter_back(parts: List[List[int]], gathered, words: int, n_units: int) -> np.ndarray:
    """Global verdict array (n_units * words, uint64) from the per-rank gathered tensors."""
    full = np.zeros(n_units * words, dtype=np.uint64)
    for idxs, t in zip(parts, gathered):
        a = t.cpu().numpy().view(np.uint64).reshape(-1, words)
        for j, u in enumerate(idxs):
            full[u * words:(u + 1) * words] = a[j]
    return full
