import sys, os
sys.path.insert(0, os.getcwd())
import torch
import arrow_gpu_b200 as ag
dev = ag.GpuDevice(0)
n = 4_000_000_000
t = torch.empty(n, dtype=torch.float32, device="cuda").uniform_(-1, 1)
torch.cuda.synchronize()
a = ag.Float32ArrayGPU(ag.ArrowGpuBuffer(dev, t.data_ptr(), n * 4, owned=False), dev, n, None)
bits = ag.BooleanArrayGPU(ag.ArrowGpuBuffer(dev, t.data_ptr(), n * 4, owned=False), dev, n * 32 - 7, None)
def timeit(fn, reps=5):
    fn(); dev.sync()
    ts = []
    for _ in range(reps):
        e0 = dev.record_event(); fn(); e1 = dev.record_event(); dev.sync(); ts.append(e0.elapsed_ms(e1))
    return min(ts), sum(ts) / len(ts)
for name, fn, nbytes in (("f32.sum (read 16 GB)", lambda: a.sum(), n * 4), ("bitmap.all (read 16 GB)", lambda: bits.all(), n * 4),
                         ("f32.neg (copy 32 GB)", lambda: a.neg(), n * 8), ("f32.gt (read 32GB... same col)", lambda: a.gt(a), n * 4 + n / 8)):
    mn, av = timeit(fn)
    print(f"{name:34s} min {mn:8.3f} ms  avg {av:8.3f} ms  {nbytes / mn / 1e6:8.1f} GB/s")
