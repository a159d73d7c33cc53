"""Run under torchrun (one rank per GPU, NCCL):  row-range shards of one global column are
processed independently; the filter's per-shard counts are exchanged on the device.  Every rank
checks its shard of the result against numpy applied to the GLOBAL column.

    python -m torch.distributed.run --nproc-per-node N tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import arrow_gpu_b200 as ag
    from arrow_gpu_b200 import sharded

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = ag.GpuDevice(local)
    n = 3_000_017                     # not a multiple of the shard alignment
    rng = np.random.default_rng(123)  # same global column on every rank
    vals = rng.integers(-2**31, 2**31, n).astype(np.int32)
    valid = rng.random(n) < 0.9
    keep = rng.random(n) < 0.37
    kvalid = rng.random(n) < 0.95
    other = rng.integers(-2**31, 2**31, n).astype(np.int32)

    b, e = sharded.row_range(n, rank, world)
    a = ag.Int32ArrayGPU.from_numpy(vals[b:e], valid[b:e], dev)
    o = ag.Int32ArrayGPU.from_numpy(other[b:e], None, dev)
    m = ag.BooleanArrayGPU.from_numpy(keep[b:e], kvalid[b:e], dev)

    # element-wise + compare + merge: no communication, shard == slice of the global result
    assert np.array_equal(a.add(o).raw_values(), (vals[b:e].astype(np.int64) + other[b:e]).astype(np.int32))
    assert np.array_equal(a.gt(o).raw_values(), vals[b:e] > other[b:e])
    assert np.array_equal(a.merge(o, m).raw_values(), np.where(keep[b:e], vals[b:e], other[b:e]))

    # filter: local compaction + count exchange -> global placement.  Both exchange paths: the
    # device-side peer-slot exchange (default) and NCCL all_gather on the same stream
    sel = keep & kvalid
    want = vals[sel]
    for how in ("peer", "nccl"):
        out, offset, total = sharded.sharded_filter(a, m, exchange=how)
        assert total == int(sel.sum()), (how, total, int(sel.sum()))
        assert offset == int(sel[:b].sum()), (how, rank, offset)
        got = out.raw_values()
        assert np.array_equal(got, want[offset:offset + len(got)]), how
        assert np.array_equal(out.null_buffer.flags(), valid[sel][offset:offset + len(got)]), how
    # many filters enqueued back to back before the host looks at any of them (exercises the slot
    # ring of the exchange: every one must see its own counts), then resolved in order
    masks, pend = [], []
    for k in range(9):
        kk = np.random.default_rng(500 + k).random(n) < (0.1 + 0.1 * k)
        masks.append(kk)
        pend.append(sharded.sharded_filter_async(a, ag.BooleanArrayGPU.from_numpy(kk[b:e], None, dev)))
    for kk, p in zip(masks, pend):
        o, off, tot = p.result()
        assert tot == int(kk.sum()) and off == int(kk[:b].sum()), (rank, tot, off)
        assert np.array_equal(o.raw_values(), vals[kk][off:off + o.len])
    # an output capacity that is too small is detected from the exchanged count and redone
    o, off, tot = sharded.sharded_filter(a, m, capacity=10)
    assert tot == int(sel.sum()) and np.array_equal(o.raw_values(), want[off:off + o.len])
    assert np.array_equal(o.null_buffer.flags(), valid[sel][off:off + o.len])

    # the global result reassembled from the shards (all_gather of the ragged pieces via padding)
    t = torch.zeros(n, dtype=torch.int32, device=f"cuda:{local}")
    t[offset:offset + len(got)] = torch.from_numpy(got).to(t.device)
    dist.all_reduce(t)
    assert np.array_equal(t[:total].cpu().numpy(), want)
    # reductions over all shards: per-shard partial on the device, one all-gather, rank-ordered fold
    partials = [vals[slice(*sharded.row_range(n, r, world))].astype(np.int64).sum() for r in range(world)]
    assert sharded.sharded_sum(a) == np.int32(np.int64(sum(partials)).astype(np.int32))
    fvals = rng.uniform(-100, 100, n).astype(np.float32)
    fa = ag.Float32ArrayGPU.from_numpy(fvals[b:e], None, dev)
    import oracle as O
    fparts = [O.sum(O.F32, fvals[slice(*sharded.row_range(n, r, world))]) for r in range(world)]
    got_f = sharded.sharded_sum(fa)
    assert got_f == sharded.combine_partial_sums(np.array(fparts, dtype=np.float32), np.float32), (got_f, fparts)
    flags = np.zeros(n, dtype=bool)
    flags[n - 5] = True                                  # a single set bit, in the last shard
    fl = ag.BooleanArrayGPU.from_numpy(flags[b:e], None, dev)
    assert sharded.sharded_any(fl) is True and sharded.sharded_all(fl) is False
    assert sharded.sharded_any(fl.bitwise_and(fl.bitwise_not())) is False
    assert sharded.sharded_all(fl.bitwise_or(fl.bitwise_not())) is True

    # global-index take over peer memory (CUDA IPC + NVLink loads inside the gather kernel)
    col = sharded.ShardedColumn(ag.Int32ArrayGPU, vals[b:e], valid[b:e], n, dev)
    gidx = np.random.default_rng(1000 + rank).integers(0, n, 200_003).astype(np.uint32)
    gidx[:3] = [0, n - 1, n + 5]                     # first row, last row, out of range (reads 0)
    taken = col.take_global(ag.UInt32ArrayGPU.from_numpy(gidx, None, dev))
    safe = np.minimum(gidx, n - 1)
    want_vals = np.where(gidx < n, vals[safe], 0)
    want_valid = np.where(gidx < n, valid[safe], False)
    assert np.array_equal(taken.raw_values(), want_vals)
    assert np.array_equal(taken.null_buffer.flags(), want_valid)
    col.close()
    del col
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU check ok: world={world}, {total} of {n} rows kept")
    del a, o, m, out, taken
    dev.sync()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
