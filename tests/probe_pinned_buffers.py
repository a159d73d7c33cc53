"""probe (not a test): D2H and H2D bandwidth of several separately allocated pinned host buffers.
Looks for the placement effect behind one slow e2e run (19 GB/s into two landing buffers while the
link probe into landing[0] alone gave 54 GB/s in the same process)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())
import arrow_gpu_b200 as ag
from arrow_gpu_b200 import _ffi

lib = _ffi.lib()
dev = ag.GpuDevice(0)
size = 1 << 30
nbuf = int(sys.argv[1]) if len(sys.argv) > 1 else 6
src = dev.create_empty_buffer(size)
bufs = [ag.GpuDevice.pinned_empty(size, np.uint8) for _ in range(nbuf)]
for b in bufs:
    b[::4096] = 1
for rnd in range(2):
    out = []
    for b in bufs:
        dev.sync()
        t0 = time.perf_counter()
        for _ in range(4):
            _ffi.check(lib.agpu_d2h_async(dev.handle, b.ctypes.data, src.ptr, size), "d2h")
        dev.sync()
        d2h = 4 * size / (time.perf_counter() - t0) / 1e9
        t0 = time.perf_counter()
        for _ in range(4):
            _ffi.check(lib.agpu_h2d(dev.handle, src.ptr, b.ctypes.data, size), "h2d")
        dev.sync()
        h2d = 4 * size / (time.perf_counter() - t0) / 1e9
        out.append((round(d2h, 1), round(h2d, 1)))
    print("round", rnd, "(d2h, h2d) GB/s per buffer:", out)
