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
    """Filter this rank's shard and learn where its output sits in the global result:
    returns (local filtered array, global offset of its first row, global row count).

    On GPUs (NCCL) the per-shard counts never leave the device before the exchange: the count
    kernel writes this rank's total into a device word, `all_gather_into_tensor` runs on the SAME
    stream (torch sees the library's stream as an ExternalStream), and one 8*world-byte readback
    gives every rank all counts — a single host synchronisation per filter instead of one for the
    local count plus one per gathered value."""
    dist = _dist()
    if dist is None or dist.get_backend() != "nccl":
        out = array.filter(mask)
        offsets, total = exchange_counts(out.len)
        rank = dist.get_rank() if dist is not None else 0
        return out, offsets[rank], total
    import torch
    from .array import ArrowComputePipeline
    dev = array.gpu_device
    tdev = torch.device("cuda", dev.ordinal)
    rank, world = dist.get_rank(), dist.get_world_size()
    with torch.cuda.stream(torch.cuda.ExternalStream(dev.stream_ptr, device=tdev)):
        mine = torch.zeros(1, dtype=torch.int64, device=tdev)
        pipeline = ArrowComputePipeline(dev, "sharded_filter")
        plan = array.filter_count_op(mask, pipeline, total_ptr=mine.data_ptr())
        gathered = torch.empty(world, dtype=torch.int64, device=tdev)
        dist.all_gather_into_tensor(gathered, mine)
        counts = gathered.cpu().tolist()          # the one synchronisation
    offsets, total = exclusive_offsets(counts)
    out = array.filter_scatter_op(plan, int(counts[rank]), pipeline)
    pipeline.finish()
    return out, offsets[rank], total
