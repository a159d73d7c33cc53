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


_EXCHANGE_CTX: dict = {}


class CountExchange:
    """Device-side exchange of one u64 per rank between the GPUs of the box (agpu_exchange_post /
    agpu_exchange_wait): every rank owns a slot area in IPC-exportable memory that all peers have
    mapped; a post is one 8-byte store per peer over NVLink, a wait polls local memory.  No host
    synchronisation and no collective-library call on the data path; torch.distributed is used
    once, here, to hand the IPC handles around."""

    def __init__(self, device):
        import ctypes as C
        from ._ffi import check, lib
        from .array import ArrowGpuBuffer
        dist = _dist()
        self.device = device
        self.rank = dist.get_rank() if dist else 0
        self.world = dist.get_world_size() if dist else 1
        nbytes = lib().agpu_exchange_bytes(self.world)
        p = C.c_void_p()
        check(lib().agpu_ipc_alloc(device.handle, nbytes, C.byref(p)), "agpu_ipc_alloc")
        self.slots = ArrowGpuBuffer(device, p.value, nbytes, kind="ipc")
        check(lib().agpu_memset(device.handle, self.slots.ptr, 0, nbytes), "agpu_memset")
        device.sync()                       # zeroed before anybody can learn the handle
        mine = device.ipc_export(self.slots)
        handles = [mine]
        if dist:
            handles = [None] * self.world
            dist.all_gather_object(handles, mine)
        self.peers = [self.slots if r == self.rank else device.ipc_open(h, nbytes) for r, h in enumerate(handles)]
        self.ptrs = (C.c_void_p * self.world)(*[b.ptr for b in self.peers])
        self.seq = 0
        barrier()

    def post(self, value_ptr: int) -> None:
        """enqueue: store the u64 at device address `value_ptr` into every rank's slot area"""
        from ._ffi import check, lib
        check(lib().agpu_exchange_post(self.device.handle, value_ptr, self.ptrs, self.rank, self.world, self.seq),
              "agpu_exchange_post")

    def wait(self, out_ptr: int, timeout_ms: int = 10000) -> None:
        """enqueue: wait on the GPU for all ranks' values of the current exchange and write
        offsets / total / status / counts (2*world+2 u64) at device address `out_ptr`"""
        from ._ffi import check, lib
        check(lib().agpu_exchange_wait(self.device.handle, self.slots.ptr, self.world, self.seq, out_ptr, timeout_ms),
              "agpu_exchange_wait")
        self.seq += 1

    def close(self) -> None:
        barrier()
        self.peers = []


def count_exchange(device) -> CountExchange:
    """the CountExchange of a device handle (created on first use; collective: every rank must call)"""
    key = (id(device), "peer-exchange")
    ctx = _EXCHANGE_CTX.get(key)
    if ctx is None:
        ctx = _EXCHANGE_CTX[key] = CountExchange(device)
    return ctx


def parse_exchange_result(words, rank: int, world: int):
    """(offsets[world], total, counts[world]) from the 2*world+2 u64 agpu_exchange_wait writes"""
    words = [int(w) for w in words]
    if words[world + 1] != 0:
        from ._ffi import AgpuError
        raise AgpuError(-5, "agpu_exchange_wait")
    return words[:world], words[world], words[world + 2: 2 * world + 2]


def _read_small(dev, buffer, nbytes):
    """the one host synchronisation of a sharded filter: a few dozen bytes into a PINNED staging
    array (a pageable destination makes the driver stage the copy: ~15 us more)"""
    key = (id(dev), "pinned-result")
    stage = _EXCHANGE_CTX.get(key)
    if stage is None:
        import numpy as np
        stage = _EXCHANGE_CTX[key] = dev.pinned_empty(1024, np.uint8)
    dev.read_into(buffer, stage[:nbytes], wait=True)
    return stage[:nbytes].copy()


class PendingShardedFilter:
    """A sharded filter whose kernels are all enqueued: the compacted rows sit in `capacity`-row
    buffers on the device, the counts of all shards are (or will be) on the device too.  Nothing
    has synchronised with the host yet; `result()` does, once, and returns what `sharded_filter`
    returns.  Dependent device work can be enqueued before calling it."""

    def __init__(self, array, plan, out, capacity, info, info_kind, rank, world, gathered=None):
        self.array, self.plan, self.out, self.capacity = array, plan, out, capacity
        self.info, self.info_kind, self.rank, self.world, self.gathered = info, info_kind, rank, world, gathered
        self._resolved = None

    def result(self):
        if self._resolved is not None:
            return self._resolved
        import numpy as np
        from .array import NullBitBufferGpu
        dev = self.array.gpu_device
        if self.info_kind == "peer":
            words = _read_small(dev, self.info, (2 * self.world + 2) * 8).view(np.uint64)
            offsets, total, counts = parse_exchange_result(words, self.rank, self.world)
        elif self.info_kind == "nccl":
            import torch
            with torch.cuda.stream(self.info):
                counts = self.gathered.cpu().tolist()      # one synchronisation, after everything was enqueued
            offsets, total = exclusive_offsets(counts)
        else:                                              # single shard: the count pass's total
            counts = [int(_read_small(dev, self.info, 8).view(np.uint64)[0])]
            offsets, total = [0], counts[0]
        count = int(counts[self.rank])
        out = self.out
        if count > self.capacity:      # the estimate was too small: redo the scatter with the exact size
            out = self.array.filter_scatter_op(self.plan, count, None)
        nb = None
        if out.null_buffer is not None:
            nb = NullBitBufferGpu(out.null_buffer.bit_buffer, count, dev)
        final = type(out)(out.data, dev, count, nb)
        self._resolved = (final, int(offsets[self.rank]), int(total))
        self.plan = self.out = None
        return self._resolved


def sharded_filter_async(array, mask, capacity=None, exchange="peer") -> PendingShardedFilter:
    """Enqueue a sharded filter without ANY host synchronisation:
        count kernel (its last CTA posts this shard's count to every peer) -> scatter kernel -> wait for the peers
    The output is sized for `capacity` rows (default: the shard's row count, always enough), so the
    scatter never waits for the count to reach the host.  exchange: "peer" = one 8-byte store per
    peer over NVLink + a polling kernel (CountExchange); "nccl" = all_gather_into_tensor on the same
    stream (kept for comparison)."""
    from .array import ArrowComputePipeline
    dist = _dist()
    dev = array.gpu_device
    rank = dist.get_rank() if dist else 0
    world = dist.get_world_size() if dist else 1
    cap = array.len if capacity is None else min(int(capacity), array.len)
    pipeline = ArrowComputePipeline(dev, "sharded_filter")
    if dist is None or dist.get_backend() != "nccl":
        plan = array.filter_count_op(mask, pipeline)
        out = array.filter_scatter_op(plan, cap, pipeline)
        pipeline.finish()
        if dist is None:
            return PendingShardedFilter(array, plan, out, cap, plan.total, "local", rank, world)
        # CPU/gloo process groups (host-logic tests): counts travel through the host
        import numpy as np
        count = int(dev.retrive_data(plan.total, 8).view(np.uint64)[0])
        offsets, total = exchange_counts(count)
        pend = PendingShardedFilter(array, plan, out, cap, None, "host", rank, world)
        pend._resolved = (type(out)(out.data, dev, count, out.null_buffer), offsets[rank], total)
        return pend
    if exchange == "peer":
        ex = count_exchange(dev)
        # the count kernel posts the shard's total to every peer itself (its last CTA): no post launch
        plan = array.filter_count_op(mask, pipeline, post=(ex.ptrs, ex.rank, ex.world, ex.seq))
        out = array.filter_scatter_op(plan, cap, pipeline)
        info = dev.create_empty_buffer((2 * world + 2) * 8)
        ex.wait(info.ptr)
        pipeline.finish()
        return PendingShardedFilter(array, plan, out, cap, info, "peer", rank, world)
    import torch
    ctx = _EXCHANGE_CTX.get(id(dev))
    if ctx is None:   # per device handle: its stream as a torch stream
        tdev = torch.device("cuda", dev.ordinal)
        ctx = torch.cuda.ExternalStream(dev.stream_ptr, device=tdev)
        torch.cuda.synchronize(tdev)
        _EXCHANGE_CTX[id(dev)] = ctx
    ext = ctx
    with torch.cuda.stream(ext):
        mine = torch.zeros(1, dtype=torch.int64, device=torch.device("cuda", dev.ordinal))
        gathered = torch.zeros(world, dtype=torch.int64, device=mine.device)
        plan = array.filter_count_op(mask, pipeline, total_ptr=mine.data_ptr())
        dist.all_gather_into_tensor(gathered, mine)
        out = array.filter_scatter_op(plan, cap, pipeline)
        pipeline.finish()
    pend = PendingShardedFilter(array, plan, out, cap, ext, "nccl", rank, world, gathered=gathered)
    pend._keep = mine
    return pend


def sharded_filter(array, mask, capacity=None, exchange="peer"):
    """Filter this rank's shard and learn where its output sits in the global result:
    returns (local filtered array, global offset of its first row, global row count).

    All kernels of the filter AND the count exchange are enqueued before the host looks at anything
    (sharded_filter_async); the single synchronisation is the final read of the 8*(2*world+2)-byte
    result block.  Round 1 synchronised between the count and the scatter pass (all_gather +
    readback + stream sync), which left the GPU idle for a host round trip per filter."""
    return sharded_filter_async(array, mask, capacity, exchange).result()


# ------------------------------------------------------------------------------------------
# reductions over all shards (local partial -> one tiny exchange)
# ------------------------------------------------------------------------------------------
def combine_partial_sums(partials, np_dtype):
    """The global sum is DEFINED as the left fold, in rank order, of the per-shard sums in the
    column's own type: f32 = sequential f32 additions (each shard's partial already follows the
    reference's 256-wide pairwise tree, aggregate.wgsl:28-37), i32/u32 = wrapping.  Every rank
    computes the same value from the same gathered partials — no dependence on the collective's
    reduction order."""
    import numpy as np
    dt = np.dtype(np_dtype)
    if dt.kind == "f":
        acc = dt.type(0)
        for v in partials:
            acc = dt.type(acc + dt.type(v))
        return acc
    bits = dt.itemsize * 8
    total = sum(int(v) for v in partials) & ((1 << bits) - 1)
    if dt.kind == "i" and total >= 1 << (bits - 1):
        total -= 1 << bits
    return dt.type(total)


def gather_words(local_word: int, device=None) -> List[int]:
    """all-gather one 32-bit pattern per rank (host value in, list of every rank's value out)"""
    dist = _dist()
    if dist is None:
        return [int(local_word)]
    import torch
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    mine = torch.tensor([int(local_word)], dtype=torch.int64, device=dev)
    gathered = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(gathered, mine)
    return [int(t.item()) for t in gathered]


def _device_word_gather(dev, fill):
    """NCCL path shared by the reductions: `fill(ptr)` enqueues a kernel that writes this rank's
    32-bit partial at device address `ptr`; the all-gather runs on the SAME stream and one
    4*world-byte readback returns every rank's partial (a single host synchronisation)."""
    import numpy as np
    import torch
    dist = _dist()
    world = dist.get_world_size()
    key = (id(dev), "words")
    ctx = _EXCHANGE_CTX.get(key)
    if ctx is None:
        tdev = torch.device("cuda", dev.ordinal)
        ctx = (torch.cuda.ExternalStream(dev.stream_ptr, device=tdev),
               torch.zeros(1, dtype=torch.int32, device=tdev), torch.zeros(world, dtype=torch.int32, device=tdev),
               torch.zeros(world, dtype=torch.int32).pin_memory())
        torch.cuda.synchronize(tdev)
        _EXCHANGE_CTX[key] = ctx
    ext, mine, gathered, host = ctx
    with torch.cuda.stream(ext):
        fill(mine.data_ptr())
        dist.all_gather_into_tensor(gathered, mine)
        host.copy_(gathered, non_blocking=True)
        ext.synchronize()
        return host.numpy().view(np.uint32).copy()


def sharded_sum(array):
    """Sum over ALL shards of a row-range-sharded f32/i32/u32 column (numpy scalar of the column's
    type, identical on every rank): agpu_sum on the shard, all-gather of the 4-byte partials,
    rank-ordered fold (combine_partial_sums)."""
    import numpy as np
    from . import _ffi
    dist = _dist()
    np_dtype = array.NP
    if dist is not None and dist.get_backend() == "nccl":
        dev = array.gpu_device
        words = _device_word_gather(dev, lambda ptr: _ffi.check(
            _ffi.lib().agpu_sum(dev.handle, array.DTYPE, array.data.ptr, array.len, ptr), "sum"))
        return combine_partial_sums(words.view(np_dtype), np_dtype)
    local = array.sum().raw_values()          # one-element array (Sum::sum, aggregate_kernels.rs:24-52)
    word = int(np.asarray(local, dtype=np_dtype).view(np.uint32)[0])
    words = np.array(gather_words(word), dtype=np.uint32)
    return combine_partial_sums(words.view(np_dtype), np_dtype)


def _sharded_flag(bits, fn_name, combine):
    import numpy as np
    from . import _ffi
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        dev = bits.gpu_device
        words = _device_word_gather(dev, lambda ptr: _ffi.check(
            getattr(_ffi.lib(), fn_name)(dev.handle, bits.data.ptr, bits.len, ptr), fn_name))
        return bool(combine(int(w) for w in words))
    local = int(getattr(bits, fn_name[len("agpu_"):])())
    return bool(combine(gather_words(local)))


def sharded_any(bits) -> bool:
    """LogicalContains::any over all shards of a sharded BooleanArrayGPU"""
    return _sharded_flag(bits, "agpu_any", max)


def sharded_all(bits) -> bool:
    """LogicalContains::all over all shards (an empty shard counts as all-true)"""
    return _sharded_flag(bits, "agpu_all", min)


class ShardedColumn:
    """One column split into contiguous row ranges, one shard per rank/GPU, with every shard
    mapped into every process (CUDA IPC) so kernels can read peer shards directly over NVLink.

    `take_global(indexes)` gathers rows by GLOBAL row number: one kernel, no request/response
    exchange (the reference has neither sharding nor a global take)."""

    def __init__(self, array_cls, values, valid, n_total: int, device):
        """values/valid: this rank's shard (numpy) of a column of n_total rows"""
        import numpy as np
        from .array import NullBitBufferGpu, pack_bits
        dist = _dist()
        self.rank = dist.get_rank() if dist else 0
        self.world = dist.get_world_size() if dist else 1
        self.cls, self.device, self.n_total = array_cls, device, n_total
        self.begin = [row_range(n_total, r, self.world)[0] for r in range(self.world)] + [n_total]
        assert len(values) == self.begin[self.rank + 1] - self.begin[self.rank]
        if isinstance(values, array_cls):      # an existing device array: copy it into exportable memory
            src, rows = values, values.len
            data = self._ipc_copy(device, src.data, rows * array_cls.NP.itemsize)
            nb = None
            if src.null_buffer is not None:
                nb = NullBitBufferGpu(self._ipc_copy(device, src.null_buffer.bit_buffer, (rows + 31) // 32 * 4), rows, device)
        else:
            rows = len(values)
            data = device.create_ipc_buffer_with_data(np.ascontiguousarray(values.astype(array_cls.NP, copy=False)))
            nb = None
            if valid is not None:
                nb = NullBitBufferGpu(device.create_ipc_buffer_with_data(pack_bits(valid)), rows, device)
        self.local = array_cls(data, device, rows, nb)
        # exchange (handle, has_validity) with every rank and map the peers' shards
        mine = (device.ipc_export(data), device.ipc_export(nb.bit_buffer) if nb is not None else None)
        everyone = [mine]
        if dist:
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine)
        self.values, self.validity = [], []
        for r, (vh, bh) in enumerate(everyone):
            rows = self.begin[r + 1] - self.begin[r]
            if r == self.rank:
                self.values.append(data)
                self.validity.append(nb.bit_buffer if nb is not None else None)
            else:
                self.values.append(device.ipc_open(vh, rows * array_cls.NP.itemsize) if rows else None)
                self.validity.append(device.ipc_open(bh, (rows + 31) // 32 * 4) if (bh is not None and rows) else None)
        self.has_validity = any(v is not None for v in self.validity)

    @staticmethod
    def _ipc_copy(device, buffer, nbytes):
        import ctypes as C
        from ._ffi import check, lib
        from .array import ArrowGpuBuffer
        p = C.c_void_p()
        check(lib().agpu_ipc_alloc(device.handle, max(nbytes, 16), C.byref(p)), "agpu_ipc_alloc")
        out = ArrowGpuBuffer(device, p.value, nbytes, kind="ipc")
        check(lib().agpu_d2d(device.handle, out.ptr, buffer.ptr, nbytes), "agpu_d2d")
        device.sync()
        return out

    def take_global(self, indexes):
        """out[j] = column[indexes[j]] for this rank's (local) UInt32 index array"""
        import ctypes as C
        from ._ffi import check, lib
        from .array import NullBitBufferGpu, bitmap_words
        dev, m = self.device, indexes.len
        vals = (C.c_void_p * self.world)(*[b.ptr if b is not None else None for b in self.values])
        bits = (C.c_void_p * self.world)(*[b.ptr if b is not None else None for b in self.validity])
        begin = (C.c_uint64 * (self.world + 1))(*self.begin)
        nb = None
        if self.has_validity:
            nb = NullBitBufferGpu(dev.create_empty_buffer(bitmap_words(m) * 4), m, dev)
        out = self.cls.empty(m, dev, nb)
        check(lib().agpu_take_sharded(dev.handle, self.cls.DTYPE, self.world, vals, bits if self.has_validity else None,
                                      begin, indexes.data.ptr, out.data.ptr, m, nb.bit_buffer.ptr if nb else None),
              "take_sharded")
        return out

    def close(self):
        """unmap the peers' shards (after a barrier: nobody may still be reading ours)"""
        barrier()
        self.values, self.validity = [], []
