"""torchrun probe: throughput of the global take kernel reading a PEER shard over NVLink, for
sequential and random global row numbers (all aimed at the other rank's shard)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    import arrow_gpu_b200 as ag
    from arrow_gpu_b200 import sharded
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = ag.GpuDevice(local)
    rows = 1 << 28                      # 1 GiB shard per rank
    n = rows * world
    shard = np.arange(rank * rows, (rank + 1) * rows, dtype=np.int64).astype(np.int32)
    col = sharded.ShardedColumn(ag.Int32ArrayGPU, shard, None, n, dev)
    peer = (rank + 1) % world
    m = 1 << 26
    rng = np.random.default_rng(rank)
    cases = {
        "local sequential": np.arange(m, dtype=np.uint32) + rank * rows,
        "local random": rng.integers(rank * rows, (rank + 1) * rows, m).astype(np.uint32),
        "peer sequential": np.arange(m, dtype=np.uint32) + peer * rows,
        "peer random": rng.integers(peer * rows, (peer + 1) * rows, m).astype(np.uint32),
        "peer random within 64 MiB": rng.integers(peer * rows, peer * rows + (1 << 24), m).astype(np.uint32),
        "all shards random": rng.integers(0, n, m).astype(np.uint32),
        "all shards sequential": (np.arange(m, dtype=np.uint64) * (n // m)).astype(np.uint32),
    }
    for name, idx in cases.items():
        gi = ag.UInt32ArrayGPU.from_numpy(idx, None, dev)
        out = col.take_global(gi)
        assert np.array_equal(out.raw_values()[:1000], idx[:1000].astype(np.int32))
        dist.barrier()
        ts = []
        for _ in range(3):
            e0 = dev.record_event()
            out = col.take_global(gi)
            e1 = dev.record_event()
            dev.sync()
            ts.append(e0.elapsed_ms(e1))
        if rank == 0:
            t = min(ts)
            print(f"{name:28s} {t:9.3f} ms  {m / t / 1e6:9.2f} G rows/s  {m * 4 / t / 1e6:8.1f} GB/s gathered payload", flush=True)
    col.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
