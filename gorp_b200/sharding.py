"""Multi-GPU sharding of the extraction path: lines are independent, so a batch shards as contiguous,
line-aligned ranges of the UTF-16 text, one per rank (one process per GPU), tables replicated, and NO text or span
data ever crosses devices. The only collective is the optional all-reduce of the per-extraction histogram
(E + 2 int64 values: NCCL over NVLink on GPUs, gloo in the CPU tests); global row numbers come from an all-gather of
the per-rank line counts (one int64 per rank)."""
from __future__ import annotations

import numpy as np


def line_aligned_range(text: np.ndarray, rank: int, world: int):
    """Units [u0, u1) of `text` owned by `rank`: nominal cuts k*N/world moved FORWARD to just after the next '\\n'
    (a rank owns the lines that start in its range; the last rank ends at N)."""
    n = int(text.size)

    def cut(k):
        if k <= 0:
            return 0
        if k >= world:
            return n
        p = n * k // world
        while p < n and (p == 0 or text[p - 1] != 0x0A):
            p += 1
        return p
    return cut(rank), cut(rank + 1)


def line_range(n_lines: int, rank: int, world: int):
    """Lines [l0, l1) of a List<String> batch owned by `rank`."""
    return n_lines * rank // world, n_lines * (rank + 1) // world


def allreduce_histogram(hist):
    """In-place SUM of a per-extraction histogram tensor over all ranks (no-op for a single process)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def global_row_base(n_local_lines: int, device=None):
    """(first global row of this rank, total rows) from an all-gather of the per-rank line counts."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0, int(n_local_lines)
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([int(n_local_lines)], dtype=torch.int64, device=device)
    allc = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    counts = [int(t.item()) for t in allc]
    return sum(counts[:rank]), sum(counts)


def rebase_line_offsets(line_off_local: np.ndarray, u0: int) -> np.ndarray:
    """Shard-relative line offsets -> offsets in the whole text."""
    return line_off_local.astype(np.int64) + int(u0)
