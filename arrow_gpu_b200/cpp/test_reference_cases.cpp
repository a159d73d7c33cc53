// test_reference_cases.cpp — reference unit tests transcribed against the C++ host mirror
// (arrow_gpu.hpp) so they read like the reference's own `test_*_op!` invocations.  Vectors come
// from the reference's tests (file:line cited per case).  Links only libagpu.so; needs a B200.
#include <cmath>
#include <cstdio>
#include <limits>

#include "arrow_gpu.hpp"

using namespace arrow_gpu;
template <typename T> using Opt = std::optional<T>;
static int failures = 0, checks = 0;
#define CHECK(cond)                                                                 \
  do {                                                                              \
    ++checks;                                                                       \
    if (!(cond)) { ++failures; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)

template <typename T> bool eq(const std::vector<T>& a, const std::vector<T>& b) { return a == b; }
static bool feq(const std::vector<float>& a, const std::vector<float>& b) {  // NaN == NaN, else exact
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (!((std::isnan(a[i]) && std::isnan(b[i])) || a[i] == b[i])) return false;
  return true;
}
constexpr auto None = std::nullopt;
static const float NaN = std::numeric_limits<float>::quiet_NaN(), Inf = std::numeric_limits<float>::infinity();

int main() {
  auto device = std::make_shared<GpuDevice>(0);

  {  // crates/arithmetic/src/i32.rs:124-185  test_{add,sub,mul,div,rem}_i32_scalar_i32
    auto hundred = Int32ArrayGPU::from_slice({100}, device);
    CHECK(eq(Int32ArrayGPU::from_slice({0, 1, 2, 3, 4}, device).add_scalar(hundred).raw_values(), {100, 101, 102, 103, 104}));
    CHECK(eq(Int32ArrayGPU::from_slice({0, 100, 200, 3, 104}, device).sub_scalar(hundred).raw_values(), {-100, 0, 100, -97, 4}));
    CHECK(eq(Int32ArrayGPU::from_slice({0, INT32_MAX, 2, 3, 4}, device).mul_scalar(hundred).raw_values(), {0, -100, 200, 300, 400}));
    CHECK(eq(Int32ArrayGPU::from_slice({0, 1, 100, 260, 450}, device).div_scalar(hundred).raw_values(), {0, 0, 1, 2, 4}));
    CHECK(eq(Int32ArrayGPU::from_slice({0, 1, 2, 3, 104}, device).rem_scalar(hundred).raw_values(), {0, 1, 2, 3, 4}));
    // the _dyn form (test_scalar_op! runs both)
    ArrowArrayGPU r = rem_scalar_dyn(Int32ArrayGPU::from_slice({0, 1, 2, 3, 104}, device), hundred);
    CHECK(eq(try_from<Int32ArrayGPU>(r).raw_values(), {0, 1, 2, 3, 4}));
  }
  {  // crates/arithmetic/src/i32.rs:233-243  test_add_i32_array_i32 (nulls AND-propagate)
    auto a = Int32ArrayGPU::from_optional_slice({0, 1, None, None, 4}, device);
    auto b = Int32ArrayGPU::from_optional_slice({1, 2, None, 4, None}, device);
    std::vector<Opt<int32_t>> want = {1, 3, None, None, None};
    CHECK(a.add(b).values() == want);
    CHECK(try_from<Int32ArrayGPU>(add_dyn(a, b)).values() == want);
  }
  {  // crates/arithmetic/src/u32.rs:85-95  0 - 100 wraps to u32::MAX - 99
    auto out = UInt32ArrayGPU::from_slice({0, 100, 200, 3, 104}, device).sub_scalar(UInt32ArrayGPU::from_slice({100}, device));
    CHECK(eq(out.raw_values(), {UINT32_MAX - 99, 0, 100, UINT32_MAX - 96, 4}));
  }
  {  // crates/arithmetic/src/u16.rs  test_add_u16_scalar_u16 (the reference's only sub-word arithmetic)
    auto out = UInt16ArrayGPU::from_slice({0, 1, 2, 3, 4}, device).add_scalar(UInt16ArrayGPU::from_slice({100}, device));
    CHECK(eq(out.raw_values(), {100, 101, 102, 103, 104}));
  }
  {  // crates/compare/src/f32.rs:18-66  test_gt_f32_array_f32 (NaN, +-inf, nulls)
    auto a = Float32ArrayGPU::from_optional_slice({-1.0f, 3.0f, -1.0f, None, None, NaN, Inf, -Inf, -Inf, Inf, NaN}, device);
    auto b = Float32ArrayGPU::from_optional_slice({0.0f, 2.0f, None, 3.0f, None, NaN, Inf, -Inf, Inf, -Inf, 3.0f}, device);
    std::vector<Opt<bool>> want = {false, true, None, None, None, false, false, false, false, true, false};
    CHECK(a.gt(b).values() == want);
    CHECK(try_from<BooleanArrayGPU>(gt_dyn(a, b)).values() == want);
    // crates/compare/src/f32.rs:258-305  test_max_f32_array_f32 (NaN-ignoring max)
    auto mx = a.max(b).values();
    std::vector<Opt<float>> wmx = {0.0f, 3.0f, None, None, None, NaN, Inf, -Inf, Inf, Inf, 3.0f};
    bool ok = mx.size() == wmx.size();
    for (size_t i = 0; ok && i < mx.size(); ++i)
      ok = mx[i].has_value() == wmx[i].has_value() && (!mx[i] || (std::isnan(*mx[i]) && std::isnan(*wmx[i])) || *mx[i] == *wmx[i]);
    CHECK(ok);
  }
  {  // crates/logical/src/i8.rs  and / shl / shr with negatives, counts are a UInt32 column
    auto a = Int8ArrayGPU::from_optional_slice({0, 1, 100, -100, None, 50}, device);
    auto b = Int8ArrayGPU::from_optional_slice({0, -1, 100, -101, 13, None}, device);
    std::vector<Opt<int8_t>> want_and = {0, 1, 100, int8_t(-100 & -101), None, None};
    CHECK(a.bitwise_and(b).values() == want_and);
    auto x = Int8ArrayGPU::from_slice({0, 1, -100, -100, 127, 5}, device);
    auto c = UInt32ArrayGPU::from_slice({0, 1, 3, 5, 5, 2}, device);
    CHECK(eq(x.bitwise_shr(c).raw_values(), {0, 0, -13, -4, 3, 1}));
    CHECK(eq(x.bitwise_shl(c).raw_values(), {0, 2, int8_t(-100 * 8), int8_t(-100 * 32), int8_t(127 * 32), 20}));
    CHECK(eq(UInt8ArrayGPU::from_slice({0, 1, 2, 3, 4}, device).bitwise_not().raw_values(), {255, 254, 253, 252, 251}));
  }
  {  // crates/cast/src/{i8,u8,i16,u16,f32,boolean}_cast.rs
    auto i8 = Int8ArrayGPU::from_slice({0, 1, 2, 3, -1, -2, -3, -7, 7}, device);
    CHECK(eq(i8.cast<Int32ArrayGPU>().raw_values(), {0, 1, 2, 3, -1, -2, -3, -7, 7}));
    CHECK(eq(i8.cast<UInt16ArrayGPU>().raw_values(), {0, 1, 2, 3, 65535, 65534, 65533, 65529, 7}));
    CHECK(feq(i8.cast<Float32ArrayGPU>().raw_values(), {0, 1, 2, 3, -1, -2, -3, -7, 7}));
    auto u16 = UInt16ArrayGPU::from_slice({0, 1, 2, 3, 255, 250, 7, 65535}, device);
    CHECK(eq(u16.cast<Int16ArrayGPU>().raw_values(), {0, 1, 2, 3, 255, 250, 7, -1}));
    CHECK(eq(try_from<UInt32ArrayGPU>(cast_dyn(u16, ArrowType::UInt32Type)).raw_values(), {0, 1, 2, 3, 255, 250, 7, 65535}));
    // f32_cast.rs:40-48 (ignored on Linux in the reference; passes here)
    CHECK(eq(Float32ArrayGPU::from_slice({0.0f, 1.0f, -1.0f, 5713.0f, -5713.0f, 255.0f, 256.0f}, device).cast<UInt8ArrayGPU>().raw_values(),
             {0, 1, 0, uint8_t(5713 % 256), 0, 255, 0}));
    auto bools = BooleanArrayGPU::from_slice({true, false, true, true, false, false, true, true, false}, device);
    CHECK(feq(try_from<Float32ArrayGPU>(cast_dyn(bools, ArrowType::Float32Type)).raw_values(), {1, 0, 1, 1, 0, 0, 1, 1, 0}));
    bool panicked = false;
    try { cast_dyn(Int32ArrayGPU::from_slice({1}, device), ArrowType::Float32Type); } catch (const Panic&) { panicked = true; }
    CHECK(panicked);  // not in the reference's cast matrix -> panic!
  }
  {  // crates/math/src/i32.rs:71-111  power incl. negative exponents; abs
    auto x = Int32ArrayGPU::from_slice({2, -2, 0, -1, 3, -2}, device);
    auto p = Int32ArrayGPU::from_slice({10, 3, -1, -1, 0, -2}, device);
    CHECK(eq(x.power(p).raw_values(), {1024, -8, 1, -1, 1, 0}));
    CHECK(eq(Int32ArrayGPU::from_slice({0, -1, INT32_MIN, 7}, device).abs().raw_values(), {0, 1, INT32_MIN, 7}));
    CHECK(feq(Float32ArrayGPU::from_slice({0.0f, 1.0f, 4.0f, 9.0f, -1.0f}, device).sqrt().raw_values(), {0, 1, 2, 3, NaN}));
  }
  {  // crates/routines/src/i32.rs:22-73  merge with three bitmaps
    auto a = Int32ArrayGPU::from_optional_slice({0, 1, None, None, 4, 4, 10, None, 50}, device);
    auto b = Int32ArrayGPU::from_optional_slice({1, 2, None, 4, None, None, 20, 30, None}, device);
    auto m = BooleanArrayGPU::from_optional_slice({true, true, false, false, true, false, None, None, false}, device);
    std::vector<Opt<int32_t>> want = {0, 1, None, 4, 4, None, None, None, None};
    CHECK(a.merge(b, m).values() == want);
    CHECK(try_from<Int32ArrayGPU>(merge_dyn(a, b, m)).values() == want);
  }
  {  // crates/routines/src/i32.rs:127-146  take with nulls; crates/routines/src/bool.rs:204-225 take bool
    auto a = Int32ArrayGPU::from_optional_slice({0, 1, None, 3}, device);
    auto idx = UInt32ArrayGPU::from_slice({0, 1, 2, 3, 0, 1, 2, 3}, device);
    std::vector<Opt<int32_t>> want = {0, 1, None, 3, 0, 1, None, 3};
    CHECK(a.take(idx).values() == want);
    CHECK(BooleanArrayGPU::from_slice({true}, device).take(UInt32ArrayGPU::from_slice(std::vector<uint32_t>(100, 0), device)).raw_values() ==
          std::vector<bool>(100, true));
    // put (routines/src/put.rs): dst[dst_idx[i]] = src[src_idx[i]]
    auto src = Int32ArrayGPU::from_slice({0, 1, 2, 3}, device);
    auto dst = Int32ArrayGPU::from_slice({0, 0, 0, 0, 0, 0, 0, 0}, device);
    src.put(UInt32ArrayGPU::from_slice({0, 1, 2, 3}, device), dst, UInt32ArrayGPU::from_slice({1, 3, 5, 7}, device));
    CHECK(eq(dst.raw_values(), {0, 0, 0, 1, 0, 2, 0, 3}));
  }
  {  // broadcast + sum (arithmetic/src/i32.rs:257-279, f32.rs:267-289) — f32 sum in the reference's tree order
    CHECK(eq(Int32ArrayGPU::broadcast(-5, 256 * 256, device).sum().raw_values(), {256 * 256 * -5}));
    CHECK(feq(Float32ArrayGPU::broadcast(5.0f, 4 * 1024 * 1024, device).sum().raw_values(), {4.0f * 1024 * 1024 * 5.0f}));
    CHECK(eq(Int8ArrayGPU::broadcast(-1, 100, device).raw_values(), std::vector<int8_t>(100, -1)));  // Q2: correct for negatives
  }
  {  // crates/logical/src/boolean.rs:259-319  any / all
    CHECK(BooleanArrayGPU::from_slice({true, true, false, true, false}, device).any());
    CHECK(!BooleanArrayGPU::from_slice(std::vector<bool>(16384, false), device).any());
    CHECK(BooleanArrayGPU::from_slice(std::vector<bool>(100, true), device).all());
    std::vector<bool> big(1024 * 1024 * 2 + 1, true);
    big.back() = false;
    CHECK(!BooleanArrayGPU::from_slice(big, device).all());
  }
  {  // new surface: fused chain == unfused chain, filter
    std::vector<float> a, b, c, d;
    for (int i = 0; i < 10007; ++i) { a.push_back(i * 0.37f - 1000); b.push_back(3.0f - i * 0.011f); c.push_back(i % 17 - 8.5f); d.push_back(std::sin(float(i)) * 5000); }
    auto ga = Float32ArrayGPU::from_slice(a, device), gb = Float32ArrayGPU::from_slice(b, device), gc = Float32ArrayGPU::from_slice(c, device),
         gd = Float32ArrayGPU::from_slice(d, device);
    CHECK(fused_mul_add_gt(ga, gb, gc, gd).raw_values() == ga.mul(gb).add(gc).gt(gd).raw_values());
    auto chained = fused_chain(ga, {ChainStep::binary(AGPU_MUL, gb), ChainStep::binary(AGPU_ADD, gc), ChainStep::compare(AGPU_GT, gd)});
    CHECK(std::get<BooleanArrayGPU>(chained).raw_values() == ga.mul(gb).add(gc).gt(gd).raw_values());
    auto chain_vals = fused_chain(Int8ArrayGPU::from_slice({0, 1, 4, 9, -16}, device), {ChainStep::unary(AGPU_ABS), ChainStep::unary(AGPU_SQRT), ChainStep::binary(AGPU_MUL, 2.0f)});
    CHECK(feq(std::get<Float32ArrayGPU>(chain_vals).raw_values(), {0, 2, 4, 6, 8}));
    {  // s = a + b and g = a > b out of one kernel; the interpreter path (min) and a third column too
      auto pa = Float32ArrayGPU::from_optional_slice({1.5f, None, -3.0f, 4.0f, 1e30f, -0.0f, 7.0f}, device);
      auto pb = Float32ArrayGPU::from_optional_slice({1.5f, 2.0f, None, -4.0f, 1e30f, 0.0f, 8.0f}, device);
      auto pc = Float32ArrayGPU::from_slice({0.f, 0.f, 0.f, 5.f, 0.f, 0.f, 7.f}, device);
      auto [ps, pg] = fused_chain_pair(pa, {ChainStep::binary(AGPU_ADD, pb)}, {ChainStep::compare(AGPU_GT, pb)});
      CHECK(ps.values() == pa.add(pb).values());
      CHECK(pg.values() == pa.gt(pb).values());
      auto [qs, qg] = fused_chain_pair(pa, {ChainStep::binary(AGPU_MIN, pb), ChainStep::binary(AGPU_MUL, 2.0f)}, {ChainStep::unary(AGPU_ABS), ChainStep::compare(AGPU_LTEQ, pb)});
      CHECK(qs.values() == pa.min(pb).mul_scalar(Float32ArrayGPU::from_slice({2.0f}, device)).values());
      CHECK(qg.values() == pa.abs().lteq(pb).values());
      auto [rs, rg] = fused_chain_pair(pc, {ChainStep::binary(AGPU_SUB, pc)}, {ChainStep::compare(AGPU_EQ, 7.0f)});
      CHECK(feq(rs.raw_values(), {0, 0, 0, 0, 0, 0, 0}));
      CHECK((rg.raw_values() == std::vector<bool>{false, false, false, false, false, false, true}));
      bool refused = false;   // g would depend on pb's bitmap, s would not
      try { fused_chain_pair(pc, {ChainStep::binary(AGPU_ADD, pc)}, {ChainStep::compare(AGPU_GT, pb)}); } catch (const Panic&) { refused = true; }
      CHECK(refused);
    }
    // dyn layer incl. the recording forms, broadcast, bitcast and put
    {
      ArrowComputePipeline pipe(device, "dyn");
      ArrowArrayGPU x = Int32ArrayGPU::from_slice({1, -2, 3}, device), y = Int32ArrayGPU::from_slice({10, 20, 30}, device);
      auto sum = add_op_dyn(x, y, pipe);
      auto m = gt_op_dyn(sum, y, pipe);
      pipe.finish();
      CHECK((std::get<Int32ArrayGPU>(sum).raw_values() == std::vector<int32_t>{11, 18, 33}));
      CHECK((std::get<BooleanArrayGPU>(m).raw_values() == std::vector<bool>{true, false, true}));
      auto bc = broadcast_dyn(ScalarValue{uint16_t(7)}, 5, device);
      CHECK((std::get<UInt16ArrayGPU>(bc).raw_values() == std::vector<uint16_t>(5, 7)));
      auto bb = broadcast_dyn(ScalarValue{true}, 35, device);
      CHECK(std::get<BooleanArrayGPU>(bb).all());
      ArrowArrayGPU bits = UInt32ArrayGPU::from_slice({0x3F800000u, 0xC0000000u}, device);
      CHECK(feq(std::get<Float32ArrayGPU>(bitcast_dyn(bits, ArrowType::Float32Type)).raw_values(), {1.0f, -2.0f}));
      ArrowArrayGPU dst = Int32ArrayGPU::from_slice({0, 0, 0, 0}, device);
      put_dyn(x, UInt32ArrayGPU::from_slice({0, 2}, device), dst, UInt32ArrayGPU::from_slice({3, 1}, device));
      CHECK((std::get<Int32ArrayGPU>(dst).raw_values() == std::vector<int32_t>{0, 3, 0, 1}));
      bool threw = false;
      try { bitcast_dyn(x, ArrowType::Float32Type); } catch (const Panic&) { threw = true; }
      CHECK(threw);
    }
    // integer chain: ((x + y) & 0x0F) * 3 with u8 wrap, then > y  — one kernel each
    auto ux = UInt8ArrayGPU::from_slice({250, 7, 16, 255, 0}, device);
    auto uy = UInt8ArrayGPU::from_slice({10, 9, 16, 1, 0}, device);
    auto u15 = UInt8ArrayGPU::from_slice({15}, device);
    auto u3 = UInt8ArrayGPU::from_slice({3}, device);
    using S8 = IntChainStep<uint8_t>;
    auto ichain = fused_chain_int(ux, {S8::binary(AGPU_ADD, uy, ux.len), S8::binary(AGPU_AND, u15, ux.len), S8::binary(AGPU_MUL, u3, ux.len)});
    CHECK((std::get<UInt8ArrayGPU>(ichain).raw_values() == std::vector<uint8_t>{12, 0, 0, 0, 0}));
    auto ishift = fused_chain_int(ux, {S8::shift(AGPU_SHR, UInt32ArrayGPU::from_slice({1, 0, 4, 7, 33}, device), ux.len), S8::binary(AGPU_ADD, u3, ux.len)});
    CHECK((std::get<UInt8ArrayGPU>(ishift).raw_values() == std::vector<uint8_t>{128, 10, 4, 4, 3}));
    auto ipred = fused_chain_int(ux, {S8::unary(AGPU_NOT), S8::compare(AGPU_GT, uy, ux.len)});
    CHECK((std::get<BooleanArrayGPU>(ipred).raw_values() == std::vector<bool>{false, true, true, false, true}));
    auto vals = Int32ArrayGPU::from_optional_slice({10, None, 30, 40, 50, 60}, device);
    auto keep = BooleanArrayGPU::from_optional_slice({true, true, false, None, true, false}, device);
    std::vector<Opt<int32_t>> want = {10, None, 50};
    CHECK(vals.filter(keep).values() == want);
  }
  {  // unsupported pairs panic! (arithmetic_kernels.rs:92-97)
    bool panicked = false;
    try { add_array_dyn(Int32ArrayGPU::from_slice({1}, device), Float32ArrayGPU::from_slice({1.0f}, device)); } catch (const Panic&) { panicked = true; }
    CHECK(panicked);
  }
  std::printf("%d checks, %d failures, %llu kernel launches\n", checks, failures, (unsigned long long)device->launch_count());
  return failures ? 1 : 0;
}
