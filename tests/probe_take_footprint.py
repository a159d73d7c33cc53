"""probe (not a test): random 4-byte gathers as a function of the SOURCE footprint.
m = 256 Mi output rows each time; indices uniform in [0, footprint) of ONE 16 GiB source column.
Prints ns per row and the implied DRAM bytes per gather if the kernel were purely DRAM-bound at
6.5 TB/s (an upper bound on bytes moved).  Also: all indices in one window at a random offset."""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

import arrow_gpu_b200 as ag

dev = ag.GpuDevice(0)
N = 1 << 32          # 4 Gi rows = 16 GiB source
m = 1 << 28
src_t = torch.empty(N, dtype=torch.int32, device="cuda")
for s in range(0, N, 1 << 28):
    src_t[s:s + (1 << 28)] = torch.arange(s, s + (1 << 28), dtype=torch.int64, device="cuda").to(torch.int32)
src = ag.Int32ArrayGPU(ag.ArrowGpuBuffer(dev, src_t.data_ptr(), N * 4, owned=False), dev, N, None)
g = torch.Generator(device="cuda").manual_seed(7)
torch.cuda.synchronize()


def run(label, idx_t):
    idx = ag.UInt32ArrayGPU(ag.ArrowGpuBuffer(dev, idx_t.data_ptr(), m * 4, owned=False), dev, m, None)
    src.take(idx)
    dev.sync()
    ts = []
    for _ in range(3):
        e0 = dev.record_event()
        out = src.take(idx)
        e1 = dev.record_event()
        dev.sync()
        ts.append(e0.elapsed_ms(e1))
    t = min(ts)
    print(f"{label:44s} {t:8.3f} ms  {t * 1e6 / m:6.3f} ns/row  {m / t / 1e6:7.2f} G rows/s  <= {t * 1e-3 * 6.5e12 / m - 8:6.1f} B/gather at 6.5 TB/s")
    del out


for bits in (24, 26, 28, 29, 30, 31, 32):
    fp = 1 << bits
    idx_t = torch.randint(0, fp, (m,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    run(f"uniform in [0, 2^{bits}) rows = {fp * 4 >> 20} MiB", idx_t)
    del idx_t
# one 256 MiB / 1 GiB window somewhere in the middle of the 16 GiB column
for bits in (26, 28):
    base = (5 << 28) + 12345 * 1024
    idx_t = (torch.randint(0, 1 << bits, (m,), dtype=torch.int64, device="cuda", generator=g) + base).to(torch.int32)
    run(f"uniform in a 2^{bits}-row window at row {base}", idx_t)
    del idx_t
# indices grouped by 1 GiB window (16 consecutive runs), random inside each: what a windowed take would see
chunks = []
per = m // 16
for w in range(16):
    chunks.append((torch.randint(0, 1 << 28, (per,), dtype=torch.int64, device="cuda", generator=g) + (w << 28)).to(torch.int32))
run("grouped by 1 GiB window, random inside", torch.cat(chunks))
# same multiset of indices, fully shuffled
allidx = torch.cat(chunks)
perm = torch.randperm(m, device="cuda", generator=g)
run("the same indices shuffled", allidx[perm])
