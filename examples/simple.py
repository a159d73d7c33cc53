"""The reference's example (crates/arrow/examples/simple.rs) on this implementation: same calls,
same checks.  `python examples/simple.py` on a machine with a B200."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import arrow_gpu_b200 as ag
from arrow_gpu_b200.kernels import add_scalar_dyn, add_scalar_op_dyn, mul_scalar_op_dyn


def run_basic_add():
    """create an array on the GPU and run a compute kernel (simple.rs:10-43)"""
    device = ag.GpuDevice.new()
    float_values = [float(x) for x in range(10)]
    gpu_float_array = ag.Float32ArrayGPU.from_slice(float_values, device)
    gpu_float_array_scalar = ag.Float32ArrayGPU.from_slice([20.0], device)

    add_scalar_result = gpu_float_array.add_scalar(gpu_float_array_scalar)
    for index, value in enumerate(add_scalar_result.values()):
        assert value == float_values[index] + 20.0

    # every kernel has a dyn form working on any array type
    dyn_result = add_scalar_dyn(gpu_float_array, gpu_float_array_scalar)
    assert isinstance(dyn_result, ag.Float32ArrayGPU), "Result should be float32 type"
    for index, value in enumerate(dyn_result.values()):
        assert value == float_values[index] + 20.0


def run_compute_pipeline_ops(fuse: bool = False):
    """several operations recorded on one pipeline (simple.rs:45-77); with fuse=True the two
    scalar ops run as ONE kernel, same result bit for bit"""
    device = ag.GpuDevice.new()
    pipeline = ag.ArrowComputePipeline(device, "example", fuse=fuse)
    float_values = [float(x) for x in range(100)]
    lhs = ag.Float32ArrayGPU.from_slice(float_values, device)
    rhs = ag.Float32ArrayGPU.from_slice([20.0], device)

    launches = device.launch_count()
    r1 = add_scalar_op_dyn(lhs, rhs, pipeline)
    r2 = mul_scalar_op_dyn(r1, rhs, pipeline)
    pipeline.finish()
    launches = device.launch_count() - launches

    assert isinstance(r2, ag.Float32ArrayGPU), "Result should be float32 type"
    for index, value in enumerate(r2.values()):
        assert value == (float_values[index] + 20.0) * 20.0
    return launches


def main():
    run_basic_add()
    plain = run_compute_pipeline_ops()
    fused = run_compute_pipeline_ops(fuse=True)
    print(f"example ok: recorded pipeline = {plain} kernel launches, fusing pipeline = {fused}")


if __name__ == "__main__":
    main()
