// bench_small.cpp — per-op cost of the C++ host mirror on small/medium columns (BASELINE.json
// config 1 shape: f32 add + gt with null bitmaps), where launch latency, not HBM, is the bound.
// Prints one line per column size: microseconds per op (1000 ops back to back, one final sync),
// eager (one launch + two allocations per op) and as ONE captured submit per add + gt pair
// (ArrowComputePipeline(capture = true): compute_pipeline.rs:259-273's record-then-submit).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <random>

#include "arrow_gpu.hpp"

using namespace arrow_gpu;

int main() {
  auto device = std::make_shared<GpuDevice>(0);
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> dist(-1000.f, 1000.f);
  for (size_t n : {size_t(1) << 16, size_t(1) << 18, size_t(1) << 20, size_t(1) << 22, size_t(1) << 24}) {
    std::vector<std::optional<float>> a(n), b(n);
    for (size_t i = 0; i < n; ++i) {
      if (rng() % 10) a[i] = dist(rng);
      if (rng() % 10) b[i] = dist(rng);
    }
    auto ga = Float32ArrayGPU::from_optional_slice(a, device);
    auto gb = Float32ArrayGPU::from_optional_slice(b, device);
    for (int w = 0; w < 50; ++w) { auto s = ga.add(gb); auto g = ga.gt(gb); }
    device->sync();
    const int reps = 500;
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
      auto s = ga.add(gb);   // 12.375 B/row
      auto g = ga.gt(gb);    // 8.5 B/row
    }
    device->sync();
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / (2.0 * reps);
    const double gbs = (12.375 + 8.5) / 2.0 * double(n) / us / 1e3;
    // the same pair recorded once, submitted `reps` times
    ArrowComputePipeline p(device, "cfg1", true);
    auto s = ga.add_op(gb, p);
    auto g = ga.gt_op(gb, p);
    p.finish();
    for (int w = 0; w < 50; ++w) p.replay();
    device->sync();
    t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) p.replay();
    device->sync();
    const double us_g = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / (2.0 * reps);
    const double gbs_g = (12.375 + 8.5) / 2.0 * double(n) / us_g / 1e3;
    std::printf("rows=2^%d  eager %.2f us per op  %.0f GB/s | captured %.2f us per op  %.0f GB/s (mean of add and gt, validity "
                "included; %llu kernels per submit)\n", (int)std::log2((double)n), us, gbs, us_g, gbs_g,
                (unsigned long long)p.kernels_per_submit());
  }
  return 0;
}
