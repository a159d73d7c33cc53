"""probe (not a test): where the time of one sharded filter goes, phase by phase, on the device.
Run under torchrun (any world size); every rank filters a 500 M-row int32 shard (the N = 8 shard of
BASELINE.json config 5).  CUDA events are recorded on the handle's stream between the phases, so
the gaps between kernels — a host round trip would show up as one — are part of the numbers.

    python -m torch.distributed.run --nproc-per-node N tests/probe_sharded_filter_timeline.py [selectivity]
"""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import arrow_gpu_b200 as ag
    from arrow_gpu_b200 import sharded
    from bench_workloads import _bitmap, _randint32

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = ag.GpuDevice(local)
    sel = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
    n = 500_000_000
    tdev = torch.device("cuda", local)
    vals = _randint32(torch, n, 40 + rank, tdev)
    bits = _bitmap(torch, n, sel, 42 + rank, tdev)
    a = ag.Int32ArrayGPU(ag.ArrowGpuBuffer(dev, vals.data_ptr(), n * 4, owned=False), dev, n, None)
    m = ag.BooleanArrayGPU(ag.ArrowGpuBuffer(dev, bits.data_ptr(), bits.numel(), owned=False), dev, n, None)
    ex = sharded.count_exchange(dev)
    info = dev.create_empty_buffer((2 * world + 2) * 8)
    torch.cuda.synchronize()
    names = ["count+scan+post", "scatter", "wait"]
    rows = []
    for it in range(25):
        dist.barrier()
        ev = [dev.record_event() for _ in range(1)]
        plan = a.filter_count_op(m, None, post=(ex.ptrs, ex.rank, ex.world, ex.seq))   # last CTA posts the total
        ev.append(dev.record_event())
        out = a.filter_scatter_op(plan, n, None)
        ev.append(dev.record_event())
        ex.wait(info.ptr)
        ev.append(dev.record_event())
        dev.sync()
        if it >= 5:
            rows.append([ev[k].elapsed_ms(ev[k + 1]) * 1e3 for k in range(3)] + [ev[0].elapsed_ms(ev[3]) * 1e3])
        del out, plan
    med = [statistics.median(r[k] for r in rows) for k in range(4)]
    t = torch.tensor(med, dtype=torch.float64, device=tdev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        med = t.tolist()
        ideal = (4.125 + 4 * sel) * n / 6541.1e9 * 1e6
        print(f"world={world} selectivity={sel} rows/GPU={n}: median us per phase, max over ranks")
        for k, name in enumerate(names):
            print(f"  {name:14s} {med[k]:8.1f}")
        print(f"  first to last event {med[3]:8.1f}   sum of the phases {sum(med[:3]):8.1f}   (no gap between them: all three are enqueued back to back)")
        print(f"  algorithmic bytes at the measured copy peak: {ideal:.1f} us -> {ideal / med[3]:.3f} of peak for the whole filter on the device")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
