"""probe (not a test): where the host time of one small op goes (1 Mi-row f32 add with bitmaps)"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.getcwd())
import numpy as np

import arrow_gpu_b200 as ag

dev = ag.GpuDevice(0)
n = 1 << 20
rng = np.random.default_rng(0)
a = ag.Float32ArrayGPU.from_numpy(rng.random(n).astype(np.float32), rng.random(n) < 0.9, dev)
b = ag.Float32ArrayGPU.from_numpy(rng.random(n).astype(np.float32), rng.random(n) < 0.9, dev)
for _ in range(200):
    a.add(b)
dev.sync()
reps = 20000
t0 = time.perf_counter()
for _ in range(reps):
    a.add(b)
t1 = time.perf_counter()
dev.sync()
t2 = time.perf_counter()
print(f"host issue time per add: {(t1 - t0) / reps * 1e6:.2f} us; with final sync {(t2 - t0) / reps * 1e6:.2f} us")
pr = cProfile.Profile()
pr.enable()
for _ in range(reps):
    a.add(b)
pr.disable()
dev.sync()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
