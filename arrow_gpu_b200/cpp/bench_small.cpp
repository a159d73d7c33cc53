// bench_small.cpp — per-op cost of the C++ host mirror on small/medium columns (BASELINE.json
// config 1 shape: f32 add + gt with null bitmaps), where launch latency, not HBM, is the bound.
// One line per column size: microseconds per op and algorithmic GB/s,
//   eager     one launch + two allocations per op, add and gt alternating, back to back;
//   captured  the add + gt pair recorded once per input copy on ArrowComputePipeline(capture = true)
//             (compute_pipeline.rs:259-273's record-then-submit) and submitted once per iteration.
// Inputs are COLD: every iteration works on another copy of the two columns (>= 512 MiB of copies,
// 4x the 126 MB L2, visited round-robin), so the bytes come from HBM like they do for big columns.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <memory>
#include <random>
#include <string>

#include "arrow_gpu.hpp"

using namespace arrow_gpu;

int main(int argc, char** argv) {
  const bool json = argc > 1 && std::string(argv[1]) == "--json";   // one JSON object for bench.py's per_config.cfg1
  auto device = std::make_shared<GpuDevice>(0);
  if (argc > 1 && std::string(argv[1]) == "--launch-floor") {
    // what one op costs when the column is tiny (1 Ki rows): the host + driver floor under every
    // per-op time above.  (a) the mirror's eager add/gt (2 allocations + launch + 2 frees),
    // (b) the bare agpu_binary call into preallocated buffers, (c) allocation + free alone.
    std::vector<std::optional<float>> a(1024, 1.0f);
    a[3] = std::nullopt;
    auto ga = Float32ArrayGPU::from_optional_slice(a, device);
    auto gb = Float32ArrayGPU::from_optional_slice(a, device);
    const int reps = 20000;
    for (int i = 0; i < 1000; ++i) { auto s = ga.add(gb); }
    device->sync();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) { auto s = ga.add(gb); auto g = ga.gt(gb); }
    auto t1 = std::chrono::steady_clock::now();   // enqueue only: the host cost
    device->sync();
    auto t2 = std::chrono::steady_clock::now();
    const double enq = std::chrono::duration<double, std::micro>(t1 - t0).count() / (2.0 * reps);
    const double all = std::chrono::duration<double, std::micro>(t2 - t0).count() / (2.0 * reps);
    auto out = ga.add(gb);
    device->sync();
    t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < 2 * reps; ++i)
      agpu_binary(device->handle(), AGPU_ADD, AGPU_F32, ga.data->ptr(), gb.data->ptr(), out.data->ptr(), 1024,
                  (const uint32_t*)ga.null_buffer->bit_buffer->ptr(), (const uint32_t*)gb.null_buffer->bit_buffer->ptr(),
                  (uint32_t*)out.null_buffer->bit_buffer->ptr());
    t1 = std::chrono::steady_clock::now();
    device->sync();
    t2 = std::chrono::steady_clock::now();
    const double bare_enq = std::chrono::duration<double, std::micro>(t1 - t0).count() / (2.0 * reps);
    const double bare_all = std::chrono::duration<double, std::micro>(t2 - t0).count() / (2.0 * reps);
    t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < 2 * reps; ++i) {
      void* p = nullptr;
      void* q = nullptr;
      agpu_alloc(device->handle(), 4096, &p);
      agpu_alloc(device->handle(), 128, &q);
      agpu_free(device->handle(), p);
      agpu_free(device->handle(), q);
    }
    t1 = std::chrono::steady_clock::now();
    const double alloc2 = std::chrono::duration<double, std::micro>(t1 - t0).count() / (2.0 * reps);
    std::printf("launch floor, 1 Ki-row f32 columns with validity, us per op: mirror eager %.2f enqueue / %.2f incl. drain; "
                "bare agpu_binary %.2f enqueue / %.2f incl. drain; 2 x (alloc + free) %.2f\n", enq, all, bare_enq, bare_all, alloc2);
    return 0;
  }
  if (json) std::printf("{\"what\": \"C++ host mirror (arrow_gpu.hpp), f32 add + gt with null bitmaps, cold inputs (rotating copies >= 512 MiB), "
                        "wall clock over >= 512 submissions; frac = algorithmic GB/s / 6541.1; fused_pair = add + gt as ONE kernel (agpu_fused_chain_pair), per PROGRAM\", \"sizes\": {");
  bool first_size = true;
  std::mt19937 rng(1);
  std::uniform_real_distribution<float> dist(-1000.f, 1000.f);
  std::vector<size_t> sizes = {size_t(1) << 16, size_t(1) << 18, size_t(1) << 20, size_t(1) << 22, size_t(1) << 24};
  if (json) sizes = {size_t(1) << 20, size_t(1) << 22, size_t(1) << 24};
  for (size_t n : sizes) {
    std::vector<std::optional<float>> a(n), b(n);
    for (size_t i = 0; i < n; ++i) {
      if (rng() % 10) a[i] = dist(rng);
      if (rng() % 10) b[i] = dist(rng);
    }
    auto ga = Float32ArrayGPU::from_optional_slice(a, device);
    auto gb = Float32ArrayGPU::from_optional_slice(b, device);
    const size_t copies = std::max<size_t>(4, std::min<size_t>(256, (size_t(512) << 20) / (8 * n)));
    std::vector<Float32ArrayGPU> as, bs;
    for (size_t k = 0; k < copies; ++k) { as.push_back(ga.clone_array()); bs.push_back(gb.clone_array()); }
    for (size_t k = 0; k < std::min<size_t>(copies, 8); ++k) { auto s = as[k].add(bs[k]); auto g = as[k].gt(bs[k]); }
    device->sync();
    const int rounds = std::max<int>(2, int(512 / copies));
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < rounds; ++r)
      for (size_t k = 0; k < copies; ++k) {
        auto s = as[k].add(bs[k]);   // 12.375 B/row
        auto g = as[k].gt(bs[k]);    // 8.5 B/row
      }
    device->sync();
    const double pairs = double(rounds) * double(copies);
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / (2.0 * pairs);
    const double gbs = (12.375 + 8.5) / 2.0 * double(n) / us / 1e3;
    // the same pair recorded once per copy, submitted `rounds` times
    std::vector<std::unique_ptr<ArrowComputePipeline>> progs;
    std::vector<Float32ArrayGPU> sums;
    std::vector<BooleanArrayGPU> preds;
    for (size_t k = 0; k < copies; ++k) {
      progs.push_back(std::make_unique<ArrowComputePipeline>(device, "cfg1", true));
      sums.push_back(as[k].add_op(bs[k], *progs.back()));
      preds.push_back(as[k].gt_op(bs[k], *progs.back()));
      progs.back()->finish();
    }
    device->sync();
    t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < rounds; ++r)
      for (auto& p : progs) p->replay();
    device->sync();
    const double us_g = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / (2.0 * pairs);
    const double gbs_g = (12.375 + 8.5) / 2.0 * double(n) / us_g / 1e3;
    // both results out of ONE kernel (agpu_fused_chain_pair): a and b read once, one launch per program
    for (size_t k = 0; k < std::min<size_t>(copies, 8); ++k)
      fused_chain_pair(as[k], {ChainStep::binary(AGPU_ADD, bs[k])}, {ChainStep::compare(AGPU_GT, bs[k])});
    device->sync();
    t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < rounds; ++r)
      for (size_t k = 0; k < copies; ++k)
        auto sg = fused_chain_pair(as[k], {ChainStep::binary(AGPU_ADD, bs[k])}, {ChainStep::compare(AGPU_GT, bs[k])});
    device->sync();
    const double us_p = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / pairs;  // per PROGRAM
    const double gbs_p = 12.5 * double(n) / us_p / 1e3;                   // what this kernel moves
    const double gbs_pu = (12.375 + 8.5) * double(n) / us_p / 1e3;        // what the two separate ops would have moved
    if (json) {
      std::printf("%s\"%zu rows\": {\"eager_us_per_op\": %.3f, \"eager_GBps\": %.1f, \"eager_frac_measured_peak\": %.4f, "
                  "\"captured_us_per_op\": %.3f, \"captured_GBps\": %.1f, \"captured_frac_measured_peak\": %.4f, "
                  "\"fused_pair_us_per_program\": %.3f, \"fused_pair_frac_measured_peak\": %.4f, \"fused_pair_frac_of_unfused_bytes\": %.4f, "
                  "\"input_copies\": %zu}",
                  first_size ? "" : ", ", n, us, gbs, gbs / 6541.1, us_g, gbs_g, gbs_g / 6541.1, us_p, gbs_p / 6541.1, gbs_pu / 6541.1, copies);
      first_size = false;
      continue;
    }
    std::printf("rows=2^%d  copies=%zu  eager %.2f us per op  %.0f GB/s (%.3f of 6541) | captured %.2f us per op  %.0f GB/s (%.3f of 6541) | "
                "add+gt as one kernel %.2f us per program  %.3f of 6541 on its own 12.5 B/row, %.3f counting the 20.875 B/row of the two ops  "
                "[mean of add and gt, validity included; %llu kernels per submit]\n", (int)std::log2((double)n), copies, us, gbs,
                gbs / 6541.1, us_g, gbs_g, gbs_g / 6541.1, us_p, gbs_p / 6541.1, gbs_pu / 6541.1,
                (unsigned long long)progs[0]->kernels_per_submit());
  }
  if (json) std::printf("}}\n");
  return 0;
}
