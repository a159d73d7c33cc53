"""probe: many captured pipelines in a row with allocations inside the capture (cache misses)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import arrow_gpu_b200 as ag

dev = ag.GpuDevice(0)
n = 1 << 20
rng = np.random.default_rng(1)
a_h = rng.uniform(-1, 1, n).astype(np.float32)
va = rng.random(n) < 0.9
base = ag.Float32ArrayGPU.from_numpy(a_h, va, dev)
pairs = [(base.clone_array(), base.clone_array()) for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 16)]
dev.sync()
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"
if mode == "warm":            # eager ops first: the cache then holds a few blocks of the right sizes
    for a, b in pairs[:4]:
        x = a.add(b); y = a.gt(b)
    del x, y
    dev.sync()
progs = []
for k, (a, b) in enumerate(pairs):
    try:
        p = ag.ArrowComputePipeline(dev, "p", capture=True)
        s = a.add_op(b, p)
        g = a.gt_op(b, p)
        p.finish()
        progs.append((p, s, g))
    except Exception as e:
        print("capture", k, "failed:", e)
        try:
            p.abort()
        except Exception as e2:
            print("abort:", e2)
        break
dev.sync()
print("captured", len(progs), "programs, mode", mode)
for p, s, g in progs:
    p.replay()
dev.sync()
print("replayed ok; sum0 =", float(progs[0][1].raw_values()[:4].sum()) if progs else None)
