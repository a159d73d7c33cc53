"""Quick tour of the Python host mirror on a B200: the two usage styles the reference documents in
its own example (crates/arrow/examples/simple.rs) — an eager method call and ops recorded on an
ArrowComputePipeline — plus the fusing pipeline this implementation adds.

    python examples/scalar_ops.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import arrow_gpu_b200 as ag
from arrow_gpu_b200 import kernels as K


def eager_and_dyn(dev):
    col = ag.Float32ArrayGPU.from_slice([float(i) for i in range(10)], dev)
    twenty = ag.Float32ArrayGPU.from_slice([20.0], dev)          # scalars are one-element columns
    want = [float(i) + 20.0 for i in range(10)]
    assert col.add_scalar(twenty).values() == want               # typed method (ArrowScalarAdd)
    out = K.add_scalar_dyn(col, twenty)                          # dtype-dispatching form
    assert type(out) is ag.Float32ArrayGPU and out.values() == want


def recorded(dev, fuse):
    n = 100
    col = ag.Float32ArrayGPU.from_slice([float(i) for i in range(n)], dev)
    twenty = ag.Float32ArrayGPU.from_slice([20.0], dev)
    pipe = ag.ArrowComputePipeline(dev, "scalar_ops", fuse=fuse)
    before = dev.launch_count()
    shifted = K.add_scalar_op_dyn(col, twenty, pipe)
    scaled = K.mul_scalar_op_dyn(shifted, twenty, pipe)
    pipe.finish()
    kernels = dev.launch_count() - before
    assert scaled.values() == [(float(i) + 20.0) * 20.0 for i in range(n)]
    return kernels


if __name__ == "__main__":
    device = ag.GpuDevice.new()
    eager_and_dyn(device)
    print(f"scalar_ops ok: {recorded(device, fuse=False)} kernels recorded one by one, "
          f"{recorded(device, fuse=True)} on a fusing pipeline")
