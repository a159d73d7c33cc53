"""probe (not a test): fused integer chains vs the same ops one kernel each, 256 Mi rows"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

import arrow_gpu_b200 as ag
from arrow_gpu_b200 import kernels as K

dev = ag.GpuDevice(0)
n = 1 << 28
g = torch.Generator(device="cuda").manual_seed(5)
PEAK = 6541.1


def column(cls, tdt):
    t = torch.randint(-100, 100, (n,), dtype=torch.int32, device="cuda", generator=g).to(tdt)
    return t, cls(ag.ArrowGpuBuffer(dev, t.data_ptr(), t.numel() * t.element_size(), owned=False), dev, n, None)


def timeit(fn, reps=5):
    fn()
    dev.sync()
    ts = []
    for _ in range(reps):
        e0 = dev.record_event()
        out = fn()
        e1 = dev.record_event()
        dev.sync()
        ts.append(e0.elapsed_ms(e1))
        del out
    return min(ts)


for name, cls, tdt, es in (("i8", ag.Int8ArrayGPU, torch.int8, 1), ("u16", ag.UInt16ArrayGPU, torch.int16, 2),
                           ("i32", ag.Int32ArrayGPU, torch.int32, 4)):
    keep = [column(cls, tdt) for _ in range(3)]
    a, b, c = (k[1] for k in keep)
    s = cls.from_slice([3], dev)
    chains = {
        "add b, and c, mul s": ([("add", b), ("bitwise_and", c), ("mul", K.DeviceScalar(s))],
                                lambda: a.add(b).bitwise_and(c).mul_scalar(s), 4 * es),
        "not, add s, xor b": ([("bitwise_not",), ("add", K.DeviceScalar(s)), ("bitwise_xor", b)],
                              lambda: a.bitwise_not().add_scalar(s).bitwise_xor(b), 3 * es),
        "mul b, add c, gt b": ([("mul", b), ("add", c), ("gt", b)],
                               lambda: a.mul(b).add(c).gt(b), 3 * es + 0.125),
    }
    for label, (steps, unfused, bpr) in chains.items():
        tf = timeit(lambda: K.fused_chain_int(a, steps))
        tu = timeit(unfused)
        print(f"{name:4s} [{label:22s}] fused {tf:.4f} ms ({bpr * n / tf / 1e6:7.1f} GB/s, frac {bpr * n / tf / 1e6 / PEAK:.3f})"
              f"   unfused {tu:.4f} ms   speed-up {tu / tf:.2f}x")
    del keep, a, b, c

# f32 arithmetic-only chains (the light interpreter, specialised on the column count)
tf32 = [torch.empty(n, dtype=torch.float32, device="cuda").uniform_(-5, 5, generator=g) for _ in range(3)]
fa, fb, fc = (ag.Float32ArrayGPU(ag.ArrowGpuBuffer(dev, t.data_ptr(), n * 4, owned=False), dev, n, None) for t in tf32)
sf = ag.Float32ArrayGPU.from_slice([1.5], dev)
for label, steps, unfused, bpr in (
        ("mul s, add s, abs", [("mul", K.DeviceScalar(sf)), ("add", K.DeviceScalar(sf)), ("abs",)],
         lambda: fa.mul_scalar(sf).add_scalar(sf).abs(), 8),
        ("mul b, add s, sqrt", [("mul", fb), ("add", K.DeviceScalar(sf)), ("sqrt",)],
         lambda: fa.mul(fb).add_scalar(sf).sqrt(), 12),
        ("mul b, add c, max s", [("mul", fb), ("add", fc), ("max", 2.0)], None, 16),
        ("sin, mul b, add c (heavy)", [("sin",), ("mul", fb), ("add", fc)], lambda: fa.sin().mul(fb).add(fc), 16)):
    tf = timeit(lambda: K.fused_chain(fa, steps))
    tu = timeit(unfused) if unfused else float("nan")
    print(f"f32  [{label:26s}] fused {tf:.4f} ms ({bpr * n / tf / 1e6:7.1f} GB/s, frac {bpr * n / tf / 1e6 / PEAK:.3f})"
          f"   unfused {tu:.4f} ms   speed-up {tu / tf:.2f}x")

# fused cast + arithmetic: u8 -> f32, * scale, + offset (5 B/row) vs cast, mul_scalar, add_scalar
tu8 = torch.randint(0, 256, (n,), dtype=torch.int32, device="cuda", generator=g).to(torch.uint8)
u8 = ag.UInt8ArrayGPU(ag.ArrowGpuBuffer(dev, tu8.data_ptr(), n, owned=False), dev, n, None)
tf = timeit(lambda: K.fused_chain(u8, [("mul", K.DeviceScalar(sf)), ("add", K.DeviceScalar(sf))]))
tu = timeit(lambda: u8.cast(ag.Float32ArrayGPU).mul_scalar(sf).add_scalar(sf))
print(f"u8   [cast f32, mul s, add s      ] fused {tf:.4f} ms ({5 * n / tf / 1e6:7.1f} GB/s, frac {5 * n / tf / 1e6 / PEAK:.3f})"
      f"   unfused {tu:.4f} ms   speed-up {tu / tf:.2f}x")

# shift by a per-row u32 counts column inside a chain (i8: 1 + 4 + 1 + 1 = 7 B/row; u16: 2 + 4 + 2 + 2 = 10)
tc = torch.randint(0, 8, (n,), dtype=torch.int32, device="cuda", generator=g)
cnt = ag.UInt32ArrayGPU(ag.ArrowGpuBuffer(dev, tc.data_ptr(), n * 4, owned=False), dev, n, None)
for name, cls, tdt, es in (("i8", ag.Int8ArrayGPU, torch.int8, 1), ("u16", ag.UInt16ArrayGPU, torch.int16, 2)):
    keep = [column(cls, tdt) for _ in range(2)]
    a, b = (k[1] for k in keep)
    steps = [("bitwise_shl", cnt), ("add", b)]
    tf = timeit(lambda: K.fused_chain_int(a, steps))
    tu = timeit(lambda: a.bitwise_shl(cnt).add(b))
    bpr = 3 * es + 4
    print(f"{name:4s} [shl cnt, add b          ] fused {tf:.4f} ms ({bpr * n / tf / 1e6:7.1f} GB/s, frac {bpr * n / tf / 1e6 / PEAK:.3f})"
          f"   unfused {tu:.4f} ms   speed-up {tu / tf:.2f}x")
