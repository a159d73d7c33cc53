"""Row-range sharding of columns across the GPUs of one box (BASELINE.json north_star (4)).

One process per GPU (torchrun); `torch.distributed` is plumbing only.  Element-wise ops, compare,
merge and shard-local take need NO communication: each rank owns a contiguous row range of every
column (values + the matching slice of every bitmap) and runs the same single-GPU kernels on it.
The only exchange on the path is for compaction outputs: ranks all-gather their selected-row
counts (8 bytes each, NCCL over NVLink on GPUs, gloo on CPU tests) and prefix-sum them into global
output offsets.  The reference has no multi-device support at all.
"""
from __future__ import annotations

from typing import List, Tuple

# shard boundaries are multiples of 1024 rows: every validity/boolean bitmap then splits on a
# 128-byte line (1024 bits) and every value column on a >= 1 KiB boundary, so each shard keeps the
# 16-byte alignment the vector kernels want
SHARD_ALIGN = 1024


def row_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the rows rank `rank` owns: near-equal contiguous ranges, aligned."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    blocks = (n_rows + SHARD_ALIGN - 1) // SHARD_ALIGN
    base, extra = divmod(blocks, world)
    b = rank * base + min(rank, extra)
    e = b + base + (1 if rank < extra else 0)
    return min(b * SHARD_ALIGN, n_rows), min(e * SHARD_ALIGN, n_rows)


def exclusive_offsets(counts: List[int]) -> Tuple[List[int], int]:
    offs, acc = [], 0
    for c in counts:
        offs.append(acc)
        acc += int(c)
    return offs, acc


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def exchange_counts(local_count: int, device=None) -> Tuple[List[int], int]:
    """All-gather one u64 count per shard and scan: returns (per-rank global offsets, total).
    With no process group this is the single-shard identity."""
    dist = _dist()
    if dist is None:
        return [0], int(local_count)
    import torch
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    mine = torch.tensor([int(local_count)], dtype=torch.int64, device=dev)
    gathered = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, mine)
    return exclusive_offsets([int(t.item()) for t in gathered])


def max_over_ranks(value: float, device=None) -> float:
    """device-time of a multi-GPU step = max over ranks"""
    dist = _dist()
    if dist is None:
        return float(value)
    import torch
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    dist = _dist()
    if dist is None:
        return float(value)
    import torch
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier() -> None:
    dist = _dist()
    if dist is not None:
        dist.barrier()


def sharded_filter(array, mask):
    """Filter this rank's shard and report where its output sits in the global result:
    returns (local filtered array, global offset of its first row, global row count)."""
    out = array.filter(mask)
    offsets, total = exchange_counts(out.len)
    dist = _dist()
    rank = dist.get_rank() if dist is not None else 0
    return out, offsets[rank], total
