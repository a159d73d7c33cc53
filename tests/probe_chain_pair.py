"""probe (not a test): the pair kernel against its parts at two sizes (CUDA events, back to back)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
import arrow_gpu_b200 as ag
from arrow_gpu_b200 import kernels as K

dev = ag.GpuDevice(0)
rng = np.random.default_rng(0)
for n in (1 << 24, 1 << 28):
    x = rng.uniform(-1, 1, n).astype(np.float32)
    valid = rng.random(n) < 0.9
    a = ag.Float32ArrayGPU.from_numpy(x, valid, dev)
    b = ag.Float32ArrayGPU.from_numpy(x[::-1].copy(), valid[::-1].copy(), dev)
    cases = {"add (12.375 B/row)": (lambda: a.add(b), 12.375), "gt (8.5)": (lambda: a.gt(b), 8.5),
             "chain [add b] (12.375)": (lambda: K.fused_chain(a, [("add", b)]), 12.375),
             "chain [gt b] (8.5)": (lambda: K.fused_chain(a, [("gt", b)]), 8.5),
             "pair add | gt (12.5)": (lambda: K.fused_chain_pair(a, [("add", b)], [("gt", b)]), 12.5),
             "pair min | gt, interpreter (12.5)": (lambda: K.fused_chain_pair(a, [("min", b)], [("gt", b)]), 12.5),
             "chain [mul b, add b, sub 1.0] (12.375)": (lambda: K.fused_chain(a, [("mul", b), ("add", b), ("sub", 1.0)]), 12.375)}
    reps = 30 if n == 1 << 24 else 8
    for name, (fn, bpr) in cases.items():
        keep = [fn(), fn(), fn()]      # three live results: the timed loop finds its blocks in the cache
        del keep
        out = fn()
        dev.sync()
        e0 = dev.record_event()
        for _ in range(reps):
            out = fn()
        e1 = dev.record_event()
        dev.sync()
        ms = e0.elapsed_ms(e1) / reps
        print(f"n=2^{n.bit_length() - 1} {name:26s} {ms * 1e3:8.1f} us  {bpr * n / ms / 1e6:7.0f} GB/s  {bpr * n / ms / 1e6 / 6541.1:.3f}")
        del out
    del a, b
