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
    rows = int(os.environ.get("PROBE_ROWS", 1 << 28))   # default: 1 GiB shard per rank
    n = rows * world
    shard = np.arange(rank * rows, (rank + 1) * rows, dtype=np.int64).astype(np.int32)
    col = sharded.ShardedColumn(ag.Int32ArrayGPU, shard, None, n, dev)
    peer = (rank + 1) % world
    m = int(os.environ.get("PROBE_M", 1 << 26))
    rng = np.random.default_rng(rank)
    cases = {
        "local sequential": np.arange(m, dtype=np.uint32) + rank * rows,
        "local random": rng.integers(rank * rows, (rank + 1) * rows, m).astype(np.uint32),
        "peer sequential": np.arange(m, dtype=np.uint32) + peer * rows,
        "peer random": rng.integers(peer * rows, (peer + 1) * rows, m).astype(np.uint32),
        "peer random within 64 MiB": rng.integers(peer * rows, peer * rows + (1 << 24), m).astype(np.uint32),
        "peer random within 1 GiB": rng.integers(peer * rows, peer * rows + min(rows, 1 << 28), m).astype(np.uint32),
        "peer random in [1 GiB, end)": rng.integers(peer * rows + min(rows - 1, 1 << 28), (peer + 1) * rows, m).astype(np.uint32),
        "peer random within 1.5 GiB": rng.integers(peer * rows, peer * rows + min(rows, 3 << 27), m).astype(np.uint32),
        "peer random within 1.75 GiB": rng.integers(peer * rows, peer * rows + min(rows, 7 << 26), m).astype(np.uint32),
        "all shards random": rng.integers(0, n, m).astype(np.uint32),
        "all shards sequential": (np.arange(m, dtype=np.uint64) * (n // m)).astype(np.uint32),
    }
    if os.environ.get("PROBE_FEW"):
        cases = {k: v for k, v in cases.items() if k.startswith("peer random") or k == "local random"}
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
            if os.environ.get("PROBE_VERBOSE"):
                print("   per call ms:", [round(x, 3) for x in ts], flush=True)
            print(f"{name:28s} {t:9.3f} ms  {m / t / 1e6:9.2f} G rows/s  {m * 4 / t / 1e6:8.1f} GB/s gathered payload", flush=True)
    col.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
