// Quick tour of the C++ host mirror (arrow_gpu_b200/cpp/arrow_gpu.hpp): an eager method call, the
// dtype-dispatching *_dyn form and two ops recorded on an ArrowComputePipeline — the usage styles
// of the reference's own example (crates/arrow/examples/simple.rs).
//   g++ -std=c++17 -I../include -I../arrow_gpu_b200/cpp scalar_ops.cpp -L../arrow_gpu_b200/lib -lagpu \
//       -Wl,-rpath,$PWD/../arrow_gpu_b200/lib -o scalar_ops
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "arrow_gpu.hpp"

namespace ag = arrow_gpu;

static void require(bool ok, const char* what) {
  if (!ok) { std::fprintf(stderr, "scalar_ops: %s\n", what); std::exit(1); }
}

static std::vector<float> ramp(int n) {
  std::vector<float> v(n);
  for (int i = 0; i < n; ++i) v[i] = (float)i;
  return v;
}

int main() {
  auto dev = std::make_shared<ag::GpuDevice>();
  auto twenty = ag::Float32ArrayGPU::from_slice({20.0f}, dev);  // scalars are one-element columns

  {  // eager method and its dyn form
    auto col = ag::Float32ArrayGPU::from_slice(ramp(10), dev);
    auto sum = col.add_scalar(twenty).raw_values();
    ag::ArrowArrayGPU any_col = col, any_scalar = twenty;
    auto dyn = ag::try_from<ag::Float32ArrayGPU>(ag::add_scalar_dyn(any_col, any_scalar)).raw_values();
    for (int i = 0; i < 10; ++i) require(sum[i] == i + 20.0f && dyn[i] == sum[i], "add_scalar");
  }
  {  // two ops on one pipeline
    ag::ArrowComputePipeline pipe(dev, "scalar_ops");
    ag::ArrowArrayGPU col = ag::Float32ArrayGPU::from_slice(ramp(100), dev), s = twenty;
    auto scaled = ag::mul_scalar_op_dyn(ag::add_scalar_op_dyn(col, s, pipe), s, pipe);
    pipe.finish();
    auto got = ag::try_from<ag::Float32ArrayGPU>(scaled).raw_values();
    for (int i = 0; i < 100; ++i) require(got[i] == (i + 20.0f) * 20.0f, "recorded ops");
  }
  std::puts("scalar_ops ok");
  return 0;
}
