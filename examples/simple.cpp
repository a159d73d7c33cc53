// The reference's example (crates/arrow/examples/simple.rs) on the C++ host mirror: same calls,
// same checks.  Build:  g++ -std=c++17 -I../include -I../arrow_gpu_b200/cpp simple.cpp \
//                        -L../arrow_gpu_b200/lib -lagpu -Wl,-rpath,'$ORIGIN/../arrow_gpu_b200/lib' -o simple
#include <cassert>
#include <cstdio>
#include <memory>
#include <vector>

#include "arrow_gpu.hpp"

using namespace arrow_gpu;

static void run_basic_add() {
  auto device = std::make_shared<GpuDevice>();
  std::vector<float> float_values;
  for (int x = 0; x < 10; ++x) float_values.push_back((float)x);
  auto gpu_float_array = Float32ArrayGPU::from_slice(float_values, device);
  auto gpu_float_array_scalar = Float32ArrayGPU::from_slice({20.0f}, device);

  auto add_scalar_result = gpu_float_array.add_scalar(gpu_float_array_scalar);
  auto values = add_scalar_result.values();
  for (size_t i = 0; i < values.size(); ++i) assert(values[i].value() == float_values[i] + 20.0f);

  ArrowArrayGPU lhs = gpu_float_array, rhs = gpu_float_array_scalar;
  auto dyn_result = add_scalar_dyn(lhs, rhs);
  const auto& x = try_from<Float32ArrayGPU>(dyn_result);  // throws ArrowErrorGPU for another type
  auto dyn_values = x.values();
  for (size_t i = 0; i < dyn_values.size(); ++i) assert(dyn_values[i].value() == float_values[i] + 20.0f);
}

static void run_compute_pipeline_ops() {
  auto device = std::make_shared<GpuDevice>();
  ArrowComputePipeline pipeline(device, "example");
  std::vector<float> float_values;
  for (int x = 0; x < 100; ++x) float_values.push_back((float)x);
  ArrowArrayGPU lhs = Float32ArrayGPU::from_slice(float_values, device);
  ArrowArrayGPU rhs = Float32ArrayGPU::from_slice({20.0f}, device);

  auto r1 = add_scalar_op_dyn(lhs, rhs, pipeline);
  auto r2 = mul_scalar_op_dyn(r1, rhs, pipeline);
  pipeline.finish();

  auto values = try_from<Float32ArrayGPU>(r2).values();
  for (size_t i = 0; i < values.size(); ++i) assert(values[i].value() == (float_values[i] + 20.0f) * 20.0f);
}

int main() {
  run_basic_add();
  run_compute_pipeline_ops();
  std::puts("example ok");
  return 0;
}
