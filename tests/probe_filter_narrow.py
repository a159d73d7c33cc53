"""probe (not a test): filter of 1-/2-byte rows and of f32 rows with validity at 256 Mi rows;
run plain for timings, or under `ncu -k regex:filter_scatter --set full` for the profile"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

import arrow_gpu_b200 as ag

dev = ag.GpuDevice(0)
n = 1 << 28
g = torch.Generator(device="cuda").manual_seed(5)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5


def column(cls, tdt):
    t = torch.randint(0, 120, (n,), dtype=torch.int32, device="cuda", generator=g).to(tdt)
    return t, cls(ag.ArrowGpuBuffer(dev, t.data_ptr(), t.numel() * t.element_size(), owned=False), dev, n, None)


def bitmap():
    t = torch.randint(-2**31, 2**31 - 1, (n // 32,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    return t, ag.ArrowGpuBuffer(dev, t.data_ptr(), n // 8, owned=False)


keep = []
tm, mbuf = bitmap()
mask = ag.BooleanArrayGPU(mbuf, dev, n, None)
tv, vbuf = bitmap()
torch.cuda.synchronize()
for name, cls, tdt, with_v, bpr in (("i8", ag.Int8ArrayGPU, torch.int8, False, 1.625), ("u16", ag.UInt16ArrayGPU, torch.int16, False, 3.125),
                                    ("i32", ag.Int32ArrayGPU, torch.int32, False, 6.125), ("f32+validity", ag.Float32ArrayGPU, torch.float32, True, 6.3125)):
    t, col = column(cls, tdt)
    if with_v:
        col.null_buffer = ag.NullBitBufferGpu(vbuf, n, dev)
    col.filter(mask)
    dev.sync()
    ts = []
    for _ in range(reps):
        e0 = dev.record_event()
        out = col.filter(mask)
        e1 = dev.record_event()
        dev.sync()
        ts.append(e0.elapsed_ms(e1))
    if not ts:
        continue
    print(f"{name:14s} filter s=0.5: min {min(ts):.4f} ms  {bpr * n / min(ts) / 1e6:7.1f} GB/s  frac {bpr * n / min(ts) / 1e6 / 6541.1:.3f}  rows out {out.len}")
    del t, col, out
