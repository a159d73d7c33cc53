// arrow_gpu.hpp — C++17 host-side mirror of psvri/arrow-gpu's array and operator crates over the
// C ABI of include/agpu.h.  Header-only; links only libagpu.so (no torch, no oracle, no CPU path).
//
// The reference is Rust; no Rust toolchain exists in this image, so the host layer above the C
// ABI is written here with the reference's names, argument meaning and error behaviour:
//   crates/array        -> GpuDevice, ArrowGpuBuffer, ArrowComputePipeline, NullBitBufferGpu,
//                          BooleanBufferBuilder, PrimitiveArrayGpu<T>, BooleanArrayGPU, ArrowArrayGPU
//   crates/arithmetic   -> add/sub/mul/div[_scalar][_op], neg, sum, *_dyn
//   crates/compare      -> gt gteq lt lteq eq, min max, *_dyn
//   crates/logical      -> bitwise_{and,or,xor,not,shl,shr}, any, all, *_dyn
//   crates/cast         -> cast<T>(), bitcast<T>(), cast_dyn, bitcast_dyn
//   crates/math         -> abs sqrt cbrt exp exp2 log log2 power, *_dyn
//   crates/trigonometry -> sin cos acos sinh, *_dyn
//   crates/routines     -> merge take put (+ filter), *_dyn
// Rust `panic!` on unsupported dtype pairs (e.g. arithmetic_kernels.rs:92-97) becomes
// `throw arrow_gpu::Panic`.  Every `x_op(.., pipeline)` enqueues one kernel on the device's
// stream; `x(..)` is the reference's default_impl! (new pipeline, op, finish).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <variant>
#include <vector>

#include "../../include/agpu.h"

namespace arrow_gpu {

struct Panic : std::runtime_error { using std::runtime_error::runtime_error; };
struct ArrowErrorGPU : std::runtime_error { using std::runtime_error::runtime_error; };  // array/src/lib.rs:10-13

inline void check(int code, const char* what) {
  if (code != 0) throw Panic(std::string(what) + ": " + agpu_error_string(code));
}

// crates/array/src/array/mod.rs:40-50
enum class ArrowType { BooleanType = AGPU_BOOL, Float32Type = AGPU_F32, UInt32Type = AGPU_U32, UInt16Type = AGPU_U16,
                       UInt8Type = AGPU_U8, Int32Type = AGPU_I32, Int16Type = AGPU_I16, Int8Type = AGPU_I8,
                       Date32Type = AGPU_DATE32 };

struct Date32Type { int32_t v; };  // marker for Date32ArrayGPU (i32 storage)

template <typename T> struct ArrowPrimitiveType;  // array/mod.rs:64-85
#define AGPU_PRIM(T, N, ID) template <> struct ArrowPrimitiveType<T> { using NativeType = N; static constexpr int DTYPE = ID; };
AGPU_PRIM(float, float, AGPU_F32) AGPU_PRIM(uint32_t, uint32_t, AGPU_U32) AGPU_PRIM(uint16_t, uint16_t, AGPU_U16)
AGPU_PRIM(uint8_t, uint8_t, AGPU_U8) AGPU_PRIM(int32_t, int32_t, AGPU_I32) AGPU_PRIM(int16_t, int16_t, AGPU_I16)
AGPU_PRIM(int8_t, int8_t, AGPU_I8) AGPU_PRIM(Date32Type, int32_t, AGPU_DATE32)
#undef AGPU_PRIM

// ---------------------------------------------------------------- device & buffers
class GpuDevice {  // gpu_utils/gpu_device.rs:29-33
 public:
  explicit GpuDevice(int ordinal = 0) { check(agpu_device_create(ordinal, &h_), "GpuDevice::new"); }
  ~GpuDevice() { if (h_) agpu_device_destroy(h_); }
  GpuDevice(const GpuDevice&) = delete;
  GpuDevice& operator=(const GpuDevice&) = delete;
  agpu_device* handle() const { return h_; }
  void sync() const { check(agpu_sync(h_), "sync"); }
  int ordinal() const { return agpu_device_ordinal(h_); }
  uint64_t launch_count() const { return agpu_launch_count(h_); }
 private:
  agpu_device* h_ = nullptr;
};
using DevicePtr = std::shared_ptr<GpuDevice>;

class ArrowGpuBuffer {  // array/buffer.rs:5-7 — owns one stream-ordered allocation
 public:
  ArrowGpuBuffer(DevicePtr dev, size_t bytes) : dev_(std::move(dev)), size_(bytes) {
    check(agpu_alloc(dev_->handle(), bytes ? bytes : 16, &ptr_), "create_empty_buffer");
  }
  // device memory owned by someone else (an imported ArrowDeviceArray): `owner` is dropped with the
  // buffer, nothing is freed here
  ArrowGpuBuffer(DevicePtr dev, void* foreign, size_t bytes, std::shared_ptr<void> owner)
      : dev_(std::move(dev)), ptr_(foreign), size_(bytes), owner_(std::move(owner)), foreign_(true) {}
  ~ArrowGpuBuffer() { if (ptr_ && !foreign_) agpu_free(dev_->handle(), ptr_); }
  ArrowGpuBuffer(const ArrowGpuBuffer&) = delete;
  void* ptr() const { return ptr_; }
  // the pointer, for work about to be enqueued on ANOTHER handle of the same GPU: the allocator
  // then keeps the block out of circulation after Drop until that handle's stream got there
  void* ptr_on(const DevicePtr& user) const {
    if (user.get() != dev_.get() && !foreign_) check(agpu_buffer_record_use(user->handle(), ptr_), "record_use");
    return ptr_;
  }
  uint64_t size() const { return size_; }
  const DevicePtr& device() const { return dev_; }
  static std::shared_ptr<ArrowGpuBuffer> with_data(const DevicePtr& dev, const void* host, size_t bytes) {
    auto b = std::make_shared<ArrowGpuBuffer>(dev, bytes);  // create_gpu_buffer_with_data
    check(agpu_h2d(dev->handle(), b->ptr_, host, bytes), "h2d");
    dev->sync();
    return b;
  }
  std::shared_ptr<ArrowGpuBuffer> clone() const {  // clone_buffer
    auto b = std::make_shared<ArrowGpuBuffer>(dev_, size_);
    check(agpu_d2d(dev_->handle(), b->ptr_, ptr_, size_), "clone_buffer");
    return b;
  }
  std::vector<uint8_t> retrive_data(size_t bytes) const {  // gpu_device.rs:232-265
    std::vector<uint8_t> out(bytes);
    check(agpu_d2h(dev_->handle(), out.data(), ptr_, bytes), "retrive_data");
    return out;
  }
 private:
  DevicePtr dev_;
  void* ptr_ = nullptr;
  size_t size_;
  std::shared_ptr<void> owner_;
  bool foreign_ = false;
};
using BufferPtr = std::shared_ptr<ArrowGpuBuffer>;

// gpu_utils/compute_pipeline.rs:8-22.  By default a stream scope: every *_op enqueues at once and
// finish() has nothing left to submit.  With capture = true it is the literal record-then-submit
// object of the reference: ops recorded between construction and finish() are captured into a CUDA
// graph (agpu_graph_begin/end), finish() submits the whole program with one driver call
// (compute_pipeline.rs:259-273) and replay() submits the same recorded program again.
class ArrowComputePipeline {
 public:
  explicit ArrowComputePipeline(DevicePtr device, const char* label = nullptr, bool capture = false)
      : device(std::move(device)), label_(label ? label : ""), capture_(capture) {
    if (capture_) check(agpu_graph_begin(this->device->handle()), "ArrowComputePipeline::new (capture)");
  }
  ArrowComputePipeline(const ArrowComputePipeline&) = delete;
  ~ArrowComputePipeline() {
    if (capture_ && !finished_) {  // abandoned recording: close the capture, submit nothing
      agpu_graph* g = nullptr;
      if (agpu_graph_end(device->handle(), &g) == 0 && g) agpu_graph_destroy(g);
    }
    if (graph_) agpu_graph_destroy(graph_);
  }
  void finish() {  // never waits, like the reference
    if (capture_ && !finished_) {
      finished_ = true;
      check(agpu_graph_end(device->handle(), &graph_), "ArrowComputePipeline::finish (capture)");
      replay();
    }
    finished_ = true;
  }
  void replay() {
    if (!graph_) throw Panic("replay() needs a pipeline created with capture = true and finished");
    check(agpu_graph_launch(device->handle(), graph_), "ArrowComputePipeline::replay");
  }
  uint64_t kernels_per_submit() const { return agpu_graph_kernel_count(graph_); }
  DevicePtr device;
 private:
  std::string label_;
  bool capture_ = false;
  bool finished_ = false;
  agpu_graph* graph_ = nullptr;
};

// ---------------------------------------------------------------- bitmaps
inline size_t bitmap_words(size_t n) { return (n + 31) / 32; }

struct BooleanBufferBuilder {  // null_bit_buffer.rs:10-62
  std::vector<uint8_t> data;
  size_t len;
  bool contains_nulls = true;
  static BooleanBufferBuilder new_with_capacity(size_t size) { return {std::vector<uint8_t>(bitmap_words(size) * 4, 0), size, true}; }
  static BooleanBufferBuilder new_set_with_capacity(size_t size) {
    BooleanBufferBuilder b = new_with_capacity(size);
    for (size_t i = 0; i < size; ++i) b.set_bit(i);
    b.contains_nulls = false;
    return b;
  }
  void set_bit(size_t pos) { data[pos / 8] |= uint8_t(1u << (pos % 8)); }
  void unset_bit(size_t pos) { data[pos / 8] &= uint8_t(~(1u << (pos % 8))); }
  bool is_set(size_t pos) const { return data[pos / 8] & (1u << (pos % 8)); }
};

struct NullBitBufferGpu {  // null_bit_buffer.rs:91-96
  BufferPtr bit_buffer;
  size_t len;
  DevicePtr gpu_device;
  static std::optional<NullBitBufferGpu> from_builder(const DevicePtr& dev, const BooleanBufferBuilder& b) {
    if (!b.contains_nulls) return std::nullopt;
    return NullBitBufferGpu{ArrowGpuBuffer::with_data(dev, b.data.data(), b.data.size()), b.len, dev};
  }
  std::vector<uint8_t> raw_values() const { return bit_buffer->retrive_data((len + 7) / 8); }
  static std::optional<NullBitBufferGpu> clone_null_bit_buffer(const std::optional<NullBitBufferGpu>& d) {
    if (!d) return std::nullopt;
    return NullBitBufferGpu{d->bit_buffer->clone(), d->len, d->gpu_device};
  }
  // null_bit_buffer.rs:168-243: AND; one side -> copy; none -> None
  static std::optional<NullBitBufferGpu> merge_null_bit_buffer(const std::optional<NullBitBufferGpu>& l,
                                                               const std::optional<NullBitBufferGpu>& r) {
    if (!l && !r) return std::nullopt;
    const NullBitBufferGpu& ref = l ? *l : *r;
    auto out = std::make_shared<ArrowGpuBuffer>(ref.gpu_device, bitmap_words(ref.len) * 4);
    check(agpu_validity_and(ref.gpu_device->handle(), l ? (const uint32_t*)l->bit_buffer->ptr() : nullptr,
                            r ? (const uint32_t*)r->bit_buffer->ptr() : nullptr, (uint32_t*)out->ptr(), ref.len),
          "merge_null_bit_buffer");
    return NullBitBufferGpu{out, ref.len, ref.gpu_device};
  }
};
using Validity = std::optional<NullBitBufferGpu>;
inline const uint32_t* vptr(const Validity& v) { return v ? (const uint32_t*)v->bit_buffer->ptr() : nullptr; }
inline uint32_t* vptr_mut(Validity& v) { return v ? (uint32_t*)v->bit_buffer->ptr() : nullptr; }
template <typename... V>
Validity new_validity(const DevicePtr& dev, size_t len, const V&... in) {
  if (!(... || bool(in))) return std::nullopt;
  return NullBitBufferGpu{std::make_shared<ArrowGpuBuffer>(dev, bitmap_words(len) * 4), len, dev};
}

// ---------------------------------------------------------------- arrays
class BooleanArrayGPU;
template <typename T> class PrimitiveArrayGpu;
using Float32ArrayGPU = PrimitiveArrayGpu<float>;
using UInt32ArrayGPU = PrimitiveArrayGpu<uint32_t>;
using UInt16ArrayGPU = PrimitiveArrayGpu<uint16_t>;
using UInt8ArrayGPU = PrimitiveArrayGpu<uint8_t>;
using Int32ArrayGPU = PrimitiveArrayGpu<int32_t>;
using Int16ArrayGPU = PrimitiveArrayGpu<int16_t>;
using Int8ArrayGPU = PrimitiveArrayGpu<int8_t>;
using Date32ArrayGPU = PrimitiveArrayGpu<Date32Type>;

class BooleanArrayGPU {  // boolean_gpu.rs:15-21
 public:
  BufferPtr data;
  DevicePtr gpu_device;
  size_t len = 0;
  Validity null_buffer;
  static constexpr int DTYPE = AGPU_BOOL;

  static BooleanArrayGPU from_slice(const std::vector<bool>& v, const DevicePtr& dev) {
    auto b = BooleanBufferBuilder::new_with_capacity(v.size());
    for (size_t i = 0; i < v.size(); ++i) if (v[i]) b.set_bit(i);
    return {ArrowGpuBuffer::with_data(dev, b.data.data(), b.data.size()), dev, v.size(), std::nullopt};
  }
  static BooleanArrayGPU from_optional_slice(const std::vector<std::optional<bool>>& v, const DevicePtr& dev) {
    auto b = BooleanBufferBuilder::new_with_capacity(v.size()), nb = b;
    for (size_t i = 0; i < v.size(); ++i) if (v[i]) { nb.set_bit(i); if (*v[i]) b.set_bit(i); }
    return {ArrowGpuBuffer::with_data(dev, b.data.data(), b.data.size()), dev, v.size(), NullBitBufferGpu::from_builder(dev, nb)};
  }
  static BooleanArrayGPU empty(size_t n, const DevicePtr& dev, Validity nb) {
    return {std::make_shared<ArrowGpuBuffer>(dev, bitmap_words(n) * 4), dev, n, std::move(nb)};
  }
  std::vector<bool> raw_values() const {
    auto raw = data->retrive_data(bitmap_words(len) * 4);
    std::vector<bool> out(len);
    for (size_t i = 0; i < len; ++i) out[i] = raw[i / 8] & (1u << (i % 8));
    return out;
  }
  std::vector<std::optional<bool>> values() const {
    auto raw = raw_values();
    std::vector<std::optional<bool>> out(len);
    std::vector<uint8_t> nv = null_buffer ? null_buffer->raw_values() : std::vector<uint8_t>();
    for (size_t i = 0; i < len; ++i)
      if (!null_buffer || (nv[i / 8] & (1u << (i % 8)))) out[i] = bool(raw[i]);
    return out;
  }
  const uint32_t* bits() const { return (const uint32_t*)data->ptr(); }

  // Logical for BooleanArrayGPU (logical/src/boolean.rs:45-104), LogicalContains (:106-147)
  BooleanArrayGPU bitwise_and_op(const BooleanArrayGPU& o, ArrowComputePipeline&) const { return logical(AGPU_AND, o); }
  BooleanArrayGPU bitwise_or_op(const BooleanArrayGPU& o, ArrowComputePipeline&) const { return logical(AGPU_OR, o); }
  BooleanArrayGPU bitwise_xor_op(const BooleanArrayGPU& o, ArrowComputePipeline&) const { return logical(AGPU_XOR, o); }
  BooleanArrayGPU bitwise_not_op(ArrowComputePipeline&) const {
    auto out = empty(len, gpu_device, new_validity(gpu_device, len, null_buffer));
    check(agpu_bitmap_not(gpu_device->handle(), bits(), (uint32_t*)out.data->ptr(), len, vptr(null_buffer), vptr_mut(out.null_buffer)), "bitwise_not");
    return out;
  }
  BooleanArrayGPU bitwise_and(const BooleanArrayGPU& o) const { return logical(AGPU_AND, o); }
  BooleanArrayGPU bitwise_or(const BooleanArrayGPU& o) const { return logical(AGPU_OR, o); }
  BooleanArrayGPU bitwise_xor(const BooleanArrayGPU& o) const { return logical(AGPU_XOR, o); }
  BooleanArrayGPU bitwise_not() const { ArrowComputePipeline p(gpu_device); return bitwise_not_op(p); }
  bool any() const { return reduce(agpu_any, "any"); }
  bool all() const { return reduce(agpu_all, "all"); }
  // Swizzle for BooleanArrayGPU (routines/src/bool.rs:48-128)
  BooleanArrayGPU merge(const BooleanArrayGPU& other, const BooleanArrayGPU& mask) const {
    auto out = empty(len, gpu_device, new_validity(gpu_device, len, null_buffer, other.null_buffer, mask.null_buffer));
    check(agpu_merge(gpu_device->handle(), AGPU_BOOL, data->ptr(), other.data->ptr(), mask.bits(), out.data->ptr(), len,
                     vptr(null_buffer), vptr(other.null_buffer), vptr(mask.null_buffer), vptr_mut(out.null_buffer)), "merge");
    return out;
  }
  template <typename Idx> BooleanArrayGPU take(const Idx& indexes) const {
    auto out = empty(indexes.len, gpu_device, new_validity(gpu_device, indexes.len, null_buffer));
    check(agpu_take(gpu_device->handle(), AGPU_BOOL, data->ptr(), len, (const uint32_t*)indexes.data->ptr(), out.data->ptr(),
                    indexes.len, vptr(null_buffer), vptr_mut(out.null_buffer)), "take");
    return out;
  }

 private:
  BooleanArrayGPU logical(int op, const BooleanArrayGPU& o) const {
    auto out = empty(len, gpu_device, new_validity(gpu_device, len, null_buffer, o.null_buffer));
    check(agpu_bitmap_binary(gpu_device->handle(), op, bits(), o.bits(), (uint32_t*)out.data->ptr(), len, vptr(null_buffer),
                             vptr(o.null_buffer), vptr_mut(out.null_buffer)), "bitwise op");
    return out;
  }
  template <typename F> bool reduce(F fn, const char* what) const {
    ArrowGpuBuffer flag(gpu_device, 4);
    check(fn(gpu_device->handle(), bits(), len, (uint32_t*)flag.ptr()), what);
    uint32_t r;
    auto raw = flag.retrive_data(4);
    std::memcpy(&r, raw.data(), 4);
    return r != 0;
  }
};

template <typename T>
class PrimitiveArrayGpu {  // primitive_array_gpu.rs:12-19 (same public fields)
 public:
  using Native = typename ArrowPrimitiveType<T>::NativeType;
  static constexpr int DTYPE = ArrowPrimitiveType<T>::DTYPE;
  BufferPtr data;
  DevicePtr gpu_device;
  size_t len = 0;
  Validity null_buffer;

  static PrimitiveArrayGpu from_slice(const std::vector<Native>& v, const DevicePtr& dev) {
    return {ArrowGpuBuffer::with_data(dev, v.data(), v.size() * sizeof(Native)), dev, v.size(), std::nullopt};
  }
  static PrimitiveArrayGpu from_optional_slice(const std::vector<std::optional<Native>>& v, const DevicePtr& dev) {
    std::vector<Native> dense(v.size(), Native{});  // nulls hold T::default() (:39-41)
    auto nb = BooleanBufferBuilder::new_with_capacity(v.size());
    for (size_t i = 0; i < v.size(); ++i) if (v[i]) { dense[i] = *v[i]; nb.set_bit(i); }
    return {ArrowGpuBuffer::with_data(dev, dense.data(), dense.size() * sizeof(Native)), dev, v.size(),
            NullBitBufferGpu::from_builder(dev, nb)};
  }
  static PrimitiveArrayGpu empty(size_t n, const DevicePtr& dev, Validity nb) {
    return {std::make_shared<ArrowGpuBuffer>(dev, n * sizeof(Native)), dev, n, std::move(nb)};
  }
  static PrimitiveArrayGpu broadcast(Native value, size_t n, const DevicePtr& dev) {  // kernels/broadcast.rs:6-17
    auto out = empty(n, dev, std::nullopt);
    check(agpu_broadcast(dev->handle(), DTYPE, &value, out.data->ptr(), n), "broadcast");
    return out;
  }
  std::vector<Native> raw_values() const {
    auto raw = data->retrive_data(len * sizeof(Native));
    std::vector<Native> out(len);
    std::memcpy(out.data(), raw.data(), raw.size());
    return out;
  }
  std::vector<std::optional<Native>> values() const {
    auto raw = raw_values();
    std::vector<std::optional<Native>> out(len);
    std::vector<uint8_t> nv = null_buffer ? null_buffer->raw_values() : std::vector<uint8_t>();
    for (size_t i = 0; i < len; ++i)
      if (!null_buffer || (nv[i / 8] & (1u << (i % 8)))) out[i] = raw[i];
    return out;
  }
  PrimitiveArrayGpu clone_array() const { return {data->clone(), gpu_device, len, NullBitBufferGpu::clone_null_bit_buffer(null_buffer)}; }

  // ---- arithmetic: ArrowScalar{Add,Sub,Mul,Div,Rem}, Arrow{Add,Sub,Mul,Div}, Neg, Sum
#define AGPU_SCALAR(name, OP)                                                                                   \
  PrimitiveArrayGpu name##_scalar_op(const PrimitiveArrayGpu& value, ArrowComputePipeline&) const { return scalar(OP, value); } \
  PrimitiveArrayGpu name##_scalar(const PrimitiveArrayGpu& value) const { return scalar(OP, value); }
  AGPU_SCALAR(add, AGPU_ADD) AGPU_SCALAR(sub, AGPU_SUB) AGPU_SCALAR(mul, AGPU_MUL) AGPU_SCALAR(div, AGPU_DIV) AGPU_SCALAR(rem, AGPU_REM)
#undef AGPU_SCALAR
#define AGPU_BINARY(name, OP)                                                                                 \
  PrimitiveArrayGpu name##_op(const PrimitiveArrayGpu& value, ArrowComputePipeline&) const { return binary(OP, value); } \
  PrimitiveArrayGpu name(const PrimitiveArrayGpu& value) const { return binary(OP, value); }
  AGPU_BINARY(add, AGPU_ADD) AGPU_BINARY(sub, AGPU_SUB) AGPU_BINARY(mul, AGPU_MUL) AGPU_BINARY(div, AGPU_DIV)
  AGPU_BINARY(min, AGPU_MIN) AGPU_BINARY(max, AGPU_MAX)                       // MinMax (compare/src/lib.rs:71-83)
  AGPU_BINARY(bitwise_and, AGPU_AND) AGPU_BINARY(bitwise_or, AGPU_OR) AGPU_BINARY(bitwise_xor, AGPU_XOR)  // Logical
  AGPU_BINARY(power, AGPU_POW)                                                // MathBinary
#undef AGPU_BINARY
#define AGPU_UNARY(name, OP)                                                                   \
  PrimitiveArrayGpu name##_op(ArrowComputePipeline&) const { return unary<T>(OP); }            \
  PrimitiveArrayGpu name() const { return unary<T>(OP); }
  AGPU_UNARY(neg, AGPU_NEG) AGPU_UNARY(abs, AGPU_ABS) AGPU_UNARY(bitwise_not, AGPU_NOT) AGPU_UNARY(sqrt, AGPU_SQRT)
  AGPU_UNARY(cbrt, AGPU_CBRT) AGPU_UNARY(exp, AGPU_EXP) AGPU_UNARY(exp2, AGPU_EXP2) AGPU_UNARY(log, AGPU_LOG)
  AGPU_UNARY(log2, AGPU_LOG2) AGPU_UNARY(acos, AGPU_ACOS)
#undef AGPU_UNARY
  // Trigonometric / Hyperbolic: integer columns give Float32ArrayGPU (fused cast)
  Float32ArrayGPU sin() const { return unary<float>(AGPU_SIN); }
  Float32ArrayGPU cos() const { return unary<float>(AGPU_COS); }
  Float32ArrayGPU sinh() const { return unary<float>(AGPU_SINH); }
  PrimitiveArrayGpu sum() const {  // aggregate_kernels.rs:24-52
    auto out = empty(1, gpu_device, std::nullopt);
    check(agpu_sum(gpu_device->handle(), DTYPE, data->ptr(), len, out.data->ptr()), "sum");
    return out;
  }
  // ---- compare (compare/src/lib.rs:41-68)
#define AGPU_CMP(name, OP)                                                                              \
  BooleanArrayGPU name##_op(const PrimitiveArrayGpu& o, ArrowComputePipeline&) const { return compare(OP, o); } \
  BooleanArrayGPU name(const PrimitiveArrayGpu& o) const { return compare(OP, o); }
  AGPU_CMP(gt, AGPU_GT) AGPU_CMP(gteq, AGPU_GTEQ) AGPU_CMP(lt, AGPU_LT) AGPU_CMP(lteq, AGPU_LTEQ) AGPU_CMP(eq, AGPU_EQ)
#undef AGPU_CMP
  // ---- shifts (logical/src/lib.rs:160-186)
  PrimitiveArrayGpu bitwise_shl(const UInt32ArrayGPU& c) const { return shift(AGPU_SHL, c); }
  PrimitiveArrayGpu bitwise_shr(const UInt32ArrayGPU& c) const { return shift(AGPU_SHR, c); }
  // ---- cast (cast/src/lib.rs:15-38): Cast<Into>::cast / BitCast<Into>::bitcast
  template <typename Into> Into cast() const {
    auto out = Into::empty(len, gpu_device, new_validity(gpu_device, len, null_buffer));
    int rc = agpu_cast(gpu_device->handle(), DTYPE, Into::DTYPE, data->ptr(), out.data->ptr(), len, vptr(null_buffer), vptr_mut(out.null_buffer));
    if (rc == AGPU_EUNSUPPORTED) throw Panic("Casting not supported for this type pair");
    check(rc, "cast");
    return out;
  }
  template <typename Into> Into bitcast() const {
    static_assert(std::is_same<T, uint32_t>::value && std::is_same<Into, Float32ArrayGPU>::value, "only u32 -> f32 (cast/src/lib.rs:187-192)");
    return cast<Into>();
  }
  // ---- routines (routines/src/lib.rs:28-72)
  PrimitiveArrayGpu merge(const PrimitiveArrayGpu& other, const BooleanArrayGPU& mask) const {
    auto out = empty(len, gpu_device, new_validity(gpu_device, len, null_buffer, other.null_buffer, mask.null_buffer));
    check(agpu_merge(gpu_device->handle(), DTYPE, data->ptr(), other.data->ptr(), mask.bits(), out.data->ptr(), len,
                     vptr(null_buffer), vptr(other.null_buffer), vptr(mask.null_buffer), vptr_mut(out.null_buffer)), "merge");
    return out;
  }
  PrimitiveArrayGpu take(const UInt32ArrayGPU& indexes) const {
    auto out = empty(indexes.len, gpu_device, new_validity(gpu_device, indexes.len, null_buffer));
    check(agpu_take(gpu_device->handle(), DTYPE, data->ptr(), len, (const uint32_t*)indexes.data->ptr(), out.data->ptr(),
                    indexes.len, vptr(null_buffer), vptr_mut(out.null_buffer)), "take");
    return out;
  }
  void put(const UInt32ArrayGPU& src_indexes, PrimitiveArrayGpu& dst, const UInt32ArrayGPU& dst_indexes) const {
    if (null_buffer || dst.null_buffer) throw Panic("put with validity is todo!() in the reference (routines/src/lib.rs:164-169)");
    check(agpu_put(gpu_device->handle(), DTYPE, data->ptr(), len, (const uint32_t*)src_indexes.data->ptr(), dst.data->ptr(),
                   dst.len, (const uint32_t*)dst_indexes.data->ptr(), src_indexes.len), "put");
  }
  PrimitiveArrayGpu filter(const BooleanArrayGPU& mask) const {  // new surface (BASELINE.json config 5)
    ArrowGpuBuffer scratch(gpu_device, agpu_filter_scratch_bytes(len)), total(gpu_device, 8);
    check(agpu_filter_count(gpu_device->handle(), mask.bits(), vptr(mask.null_buffer), len, scratch.ptr(), (uint64_t*)total.ptr()), "filter_count");
    uint64_t count;
    auto raw = total.retrive_data(8);
    std::memcpy(&count, raw.data(), 8);
    Validity nb;
    if (null_buffer) nb = NullBitBufferGpu{std::make_shared<ArrowGpuBuffer>(gpu_device, bitmap_words(count) * 4), count, gpu_device};
    auto out = empty(count, gpu_device, nb);
    check(agpu_filter_scatter(gpu_device->handle(), DTYPE, data->ptr(), vptr(null_buffer), mask.bits(), vptr(mask.null_buffer), len,
                              scratch.ptr(), out.data->ptr(), vptr_mut(out.null_buffer), count), "filter_scatter");
    return out;
  }

 private:
  PrimitiveArrayGpu binary(int op, const PrimitiveArrayGpu& o) const {
    if (len != o.len) throw Panic("length mismatch");
    auto out = empty(len, gpu_device, new_validity(gpu_device, len, null_buffer, o.null_buffer));
    int rc = agpu_binary(gpu_device->handle(), op, DTYPE, data->ptr(), o.data->ptr(), out.data->ptr(), len, vptr(null_buffer),
                         vptr(o.null_buffer), vptr_mut(out.null_buffer));
    if (rc == AGPU_EUNSUPPORTED) throw Panic("Operation not supported for this type");
    check(rc, "binary op");
    return out;
  }
  PrimitiveArrayGpu scalar(int op, const PrimitiveArrayGpu& s) const {
    if (s.len != 1) throw Panic("scalar operand must have one element");
    auto out = empty(len, gpu_device, new_validity(gpu_device, len, null_buffer));
    check(agpu_scalar(gpu_device->handle(), op, DTYPE, data->ptr(), s.data->ptr(), out.data->ptr(), len, vptr(null_buffer),
                      vptr_mut(out.null_buffer)), "scalar op");
    return out;
  }
  template <typename TO> PrimitiveArrayGpu<TO> unary(int op) const {
    auto out = PrimitiveArrayGpu<TO>::empty(len, gpu_device, new_validity(gpu_device, len, null_buffer));
    int rc = agpu_unary(gpu_device->handle(), op, DTYPE, data->ptr(), out.data->ptr(), len, vptr(null_buffer), vptr_mut(out.null_buffer));
    if (rc == AGPU_EUNSUPPORTED) throw Panic("Operation not supported for this type");
    check(rc, "unary op");
    return out;
  }
  BooleanArrayGPU compare(int op, const PrimitiveArrayGpu& o) const {
    if (len != o.len) throw Panic("length mismatch");
    auto out = BooleanArrayGPU::empty(len, gpu_device, new_validity(gpu_device, len, null_buffer, o.null_buffer));
    check(agpu_compare(gpu_device->handle(), op, DTYPE, data->ptr(), o.data->ptr(), (uint32_t*)out.data->ptr(), len,
                       vptr(null_buffer), vptr(o.null_buffer), vptr_mut(out.null_buffer)), "compare");
    return out;
  }
  PrimitiveArrayGpu shift(int op, const UInt32ArrayGPU& c) const {
    auto out = empty(len, gpu_device, new_validity(gpu_device, len, null_buffer, c.null_buffer));
    int rc = agpu_shift(gpu_device->handle(), op, DTYPE, data->ptr(), (const uint32_t*)c.data->ptr(), out.data->ptr(), len,
                        vptr(null_buffer), vptr(c.null_buffer), vptr_mut(out.null_buffer));
    if (rc == AGPU_EUNSUPPORTED) throw Panic("Operation not supported for this type");
    check(rc, "shift");
    return out;
  }
};

// ---------------------------------------------------------------- enum ArrowArrayGPU + *_dyn
using ArrowArrayGPU = std::variant<Float32ArrayGPU, UInt32ArrayGPU, UInt16ArrayGPU, UInt8ArrayGPU, Int32ArrayGPU,
                                   Int16ArrayGPU, Int8ArrayGPU, Date32ArrayGPU, BooleanArrayGPU>;  // array/mod.rs:104-114

template <typename A> const A& try_from(const ArrowArrayGPU& v) {  // TryFrom<ArrowArrayGPU>
  if (auto p = std::get_if<A>(&v)) return *p;
  throw ArrowErrorGPU("CastingNotSupported");
}

namespace detail {
template <typename F> ArrowArrayGPU same_type(const ArrowArrayGPU& a, const ArrowArrayGPU& b, const char* what, F f) {
  return std::visit([&](const auto& x) -> ArrowArrayGPU {
    using A = std::decay_t<decltype(x)>;
    if (auto y = std::get_if<A>(&b)) return f(x, *y);
    throw Panic(std::string("Operation ") + what + " not supported for this type pair");
  }, a);
}
template <typename A> constexpr bool is_bool = std::is_same<A, BooleanArrayGPU>::value;
}  // namespace detail

#define AGPU_DYN_PRIM2(fn, method)                                                                                \
  inline ArrowArrayGPU fn(const ArrowArrayGPU& a, const ArrowArrayGPU& b) {                                       \
    return detail::same_type(a, b, #fn, [](const auto& x, const auto& y) -> ArrowArrayGPU {                       \
      if constexpr (detail::is_bool<std::decay_t<decltype(x)>>) throw Panic("Operation " #fn " not supported for BooleanType"); \
      else return x.method(y);                                                                                    \
    });                                                                                                           \
  }
AGPU_DYN_PRIM2(add_array_dyn, add) AGPU_DYN_PRIM2(sub_array_dyn, sub) AGPU_DYN_PRIM2(mul_array_dyn, mul) AGPU_DYN_PRIM2(div_array_dyn, div)
AGPU_DYN_PRIM2(add_scalar_dyn, add_scalar) AGPU_DYN_PRIM2(sub_scalar_dyn, sub_scalar) AGPU_DYN_PRIM2(mul_scalar_dyn, mul_scalar)
AGPU_DYN_PRIM2(div_scalar_dyn, div_scalar) AGPU_DYN_PRIM2(rem_scalar_dyn, rem_scalar)
AGPU_DYN_PRIM2(min_dyn, min) AGPU_DYN_PRIM2(max_dyn, max) AGPU_DYN_PRIM2(power_dyn, power)
AGPU_DYN_PRIM2(gt_dyn, gt) AGPU_DYN_PRIM2(gteq_dyn, gteq) AGPU_DYN_PRIM2(lt_dyn, lt) AGPU_DYN_PRIM2(lteq_dyn, lteq) AGPU_DYN_PRIM2(eq_dyn, eq)
#undef AGPU_DYN_PRIM2
#define AGPU_DYN_ANY2(fn, method)                                                          \
  inline ArrowArrayGPU fn(const ArrowArrayGPU& a, const ArrowArrayGPU& b) {                \
    return detail::same_type(a, b, #fn, [](const auto& x, const auto& y) -> ArrowArrayGPU { return x.method(y); }); \
  }
AGPU_DYN_ANY2(bitwise_and_dyn, bitwise_and) AGPU_DYN_ANY2(bitwise_or_dyn, bitwise_or) AGPU_DYN_ANY2(bitwise_xor_dyn, bitwise_xor)
#undef AGPU_DYN_ANY2

inline size_t len_of(const ArrowArrayGPU& a) { return std::visit([](const auto& x) { return x.len; }, a); }

// add_dyn & co: operand-length routing of arithmetic_kernels.rs:101-119
#define AGPU_DYN_ROUTED(fn, array_fn, scalar_fn)                                           \
  inline ArrowArrayGPU fn(const ArrowArrayGPU& a, const ArrowArrayGPU& b) {                \
    size_t x = len_of(a), y = len_of(b);                                                   \
    if ((x == 1 && y == 1) || (x != 1 && y != 1)) return array_fn(a, b);                   \
    if (y == 1) return scalar_fn(a, b);                                                    \
    return scalar_fn(b, a);                                                                \
  }
AGPU_DYN_ROUTED(add_dyn, add_array_dyn, add_scalar_dyn) AGPU_DYN_ROUTED(sub_dyn, sub_array_dyn, sub_scalar_dyn)
AGPU_DYN_ROUTED(mul_dyn, mul_array_dyn, mul_scalar_dyn) AGPU_DYN_ROUTED(div_dyn, div_array_dyn, div_scalar_dyn)
#undef AGPU_DYN_ROUTED

#define AGPU_DYN_UNARY(fn, method)                                                         \
  inline ArrowArrayGPU fn(const ArrowArrayGPU& a) {                                        \
    return std::visit([](const auto& x) -> ArrowArrayGPU {                                 \
      if constexpr (detail::is_bool<std::decay_t<decltype(x)>>) throw Panic("Operation " #fn " not supported for BooleanType"); \
      else return x.method();                                                              \
    }, a);                                                                                 \
  }
AGPU_DYN_UNARY(neg_dyn, neg) AGPU_DYN_UNARY(abs_dyn, abs) AGPU_DYN_UNARY(sqrt_dyn, sqrt) AGPU_DYN_UNARY(cbrt_dyn, cbrt)
AGPU_DYN_UNARY(exp_dyn, exp) AGPU_DYN_UNARY(exp2_dyn, exp2) AGPU_DYN_UNARY(log_dyn, log) AGPU_DYN_UNARY(log2_dyn, log2)
AGPU_DYN_UNARY(sin_dyn, sin) AGPU_DYN_UNARY(cos_dyn, cos) AGPU_DYN_UNARY(acos_dyn, acos) AGPU_DYN_UNARY(sinh_dyn, sinh)
#undef AGPU_DYN_UNARY
inline ArrowArrayGPU bitwise_not_dyn(const ArrowArrayGPU& a) {
  return std::visit([](const auto& x) -> ArrowArrayGPU { return x.bitwise_not(); }, a);
}
inline ArrowArrayGPU bitwise_shl_dyn(const ArrowArrayGPU& a, const ArrowArrayGPU& c) {
  const auto& counts = try_from<UInt32ArrayGPU>(c);
  return std::visit([&](const auto& x) -> ArrowArrayGPU {
    if constexpr (detail::is_bool<std::decay_t<decltype(x)>>) throw Panic("shift not supported for BooleanType");
    else return x.bitwise_shl(counts);
  }, a);
}
inline ArrowArrayGPU bitwise_shr_dyn(const ArrowArrayGPU& a, const ArrowArrayGPU& c) {
  const auto& counts = try_from<UInt32ArrayGPU>(c);
  return std::visit([&](const auto& x) -> ArrowArrayGPU {
    if constexpr (detail::is_bool<std::decay_t<decltype(x)>>) throw Panic("shift not supported for BooleanType");
    else return x.bitwise_shr(counts);
  }, a);
}
inline ArrowArrayGPU merge_dyn(const ArrowArrayGPU& a, const ArrowArrayGPU& b, const BooleanArrayGPU& mask) {
  return detail::same_type(a, b, "merge", [&](const auto& x, const auto& y) -> ArrowArrayGPU { return x.merge(y, mask); });
}
inline ArrowArrayGPU take_dyn(const ArrowArrayGPU& a, const UInt32ArrayGPU& idx) {
  return std::visit([&](const auto& x) -> ArrowArrayGPU { return x.take(idx); }, a);
}
// cast_dyn (cast/src/lib.rs:135-161): the C ABI rejects pairs outside the reference matrix
inline ArrowArrayGPU cast_dyn(const ArrowArrayGPU& from, ArrowType into) {
  return std::visit([&](const auto& x) -> ArrowArrayGPU {
    using A = std::decay_t<decltype(x)>;
    if constexpr (detail::is_bool<A>) {
      if (into != ArrowType::Float32Type) throw Panic("Casting not supported");
      auto out = Float32ArrayGPU::empty(x.len, x.gpu_device, new_validity(x.gpu_device, x.len, x.null_buffer));
      check(agpu_cast(x.gpu_device->handle(), AGPU_BOOL, AGPU_F32, x.data->ptr(), out.data->ptr(), x.len, vptr(x.null_buffer),
                      vptr_mut(out.null_buffer)), "cast");
      return out;
    } else {
      switch (into) {
        case ArrowType::Float32Type: return x.template cast<Float32ArrayGPU>();
        case ArrowType::UInt32Type: return x.template cast<UInt32ArrayGPU>();
        case ArrowType::UInt16Type: return x.template cast<UInt16ArrayGPU>();
        case ArrowType::UInt8Type: return x.template cast<UInt8ArrayGPU>();
        case ArrowType::Int32Type: return x.template cast<Int32ArrayGPU>();
        case ArrowType::Int16Type: return x.template cast<Int16ArrayGPU>();
        case ArrowType::Int8Type: return x.template cast<Int8ArrayGPU>();
        default: throw Panic("Casting not supported");
      }
    }
  }, from);
}

// bitcast_dyn (cast/src/lib.rs:187-192): only u32 -> f32 exists
inline ArrowArrayGPU bitcast_dyn(const ArrowArrayGPU& from, ArrowType into) {
  if (auto u = std::get_if<UInt32ArrayGPU>(&from); u && into == ArrowType::Float32Type) return u->bitcast<Float32ArrayGPU>();
  throw Panic("Casting not supported");
}
// put_dyn (routines/src/put.rs:59-108): same-typed 32-bit columns
inline void put_dyn(const ArrowArrayGPU& src, const UInt32ArrayGPU& src_indexes, ArrowArrayGPU& dst, const UInt32ArrayGPU& dst_indexes) {
  std::visit([&](const auto& x) {
    using A = std::decay_t<decltype(x)>;
    if constexpr (detail::is_bool<A>) {
      throw Panic("Put Operation on BooleanType: use the C ABI (agpu_put with AGPU_BOOL)");
    } else if constexpr (sizeof(typename A::Native) != 4) {
      throw Panic("Put Operation not supported for this type");
    } else {
      if (auto y = std::get_if<A>(&dst)) x.put(src_indexes, *y, dst_indexes);
      else throw Panic("Put Operation not supported for this type pair");
    }
  }, src);
}
// broadcast_dyn (array/mod.rs:189-200): ScalarValue::F32(x) ... -> the variant alternative
using ScalarValue = std::variant<float, uint32_t, uint16_t, uint8_t, int32_t, int16_t, int8_t, bool>;
inline ArrowArrayGPU broadcast_dyn(const ScalarValue& value, size_t len, const DevicePtr& device) {
  return std::visit([&](auto v) -> ArrowArrayGPU {
    using V = decltype(v);
    if constexpr (std::is_same<V, bool>::value) return BooleanArrayGPU::from_slice(std::vector<bool>(len, v), device);
    else return PrimitiveArrayGpu<V>::broadcast(v, len, device);
  }, value);
}

// the recording forms (`*_op_dyn`, pipeline last): a pipeline is a stream scope here, so they
// enqueue exactly what the eager forms enqueue
#define AGPU_OP_DYN1(n) inline ArrowArrayGPU n##_op_dyn(const ArrowArrayGPU& a, ArrowComputePipeline&) { return n##_dyn(a); }
#define AGPU_OP_DYN2(n) \
  inline ArrowArrayGPU n##_op_dyn(const ArrowArrayGPU& a, const ArrowArrayGPU& b, ArrowComputePipeline&) { return n##_dyn(a, b); }
AGPU_OP_DYN1(neg) AGPU_OP_DYN1(abs) AGPU_OP_DYN1(sqrt) AGPU_OP_DYN1(cbrt) AGPU_OP_DYN1(exp) AGPU_OP_DYN1(exp2) AGPU_OP_DYN1(log)
AGPU_OP_DYN1(log2) AGPU_OP_DYN1(sin) AGPU_OP_DYN1(cos) AGPU_OP_DYN1(acos) AGPU_OP_DYN1(sinh) AGPU_OP_DYN1(bitwise_not)
AGPU_OP_DYN2(add) AGPU_OP_DYN2(sub) AGPU_OP_DYN2(mul) AGPU_OP_DYN2(div)
AGPU_OP_DYN2(add_array) AGPU_OP_DYN2(sub_array) AGPU_OP_DYN2(mul_array) AGPU_OP_DYN2(div_array)
AGPU_OP_DYN2(add_scalar) AGPU_OP_DYN2(sub_scalar) AGPU_OP_DYN2(mul_scalar) AGPU_OP_DYN2(div_scalar) AGPU_OP_DYN2(rem_scalar)
AGPU_OP_DYN2(min) AGPU_OP_DYN2(max) AGPU_OP_DYN2(power)
AGPU_OP_DYN2(gt) AGPU_OP_DYN2(gteq) AGPU_OP_DYN2(lt) AGPU_OP_DYN2(lteq) AGPU_OP_DYN2(eq)
AGPU_OP_DYN2(bitwise_and) AGPU_OP_DYN2(bitwise_or) AGPU_OP_DYN2(bitwise_xor) AGPU_OP_DYN2(bitwise_shl) AGPU_OP_DYN2(bitwise_shr)
#undef AGPU_OP_DYN1
#undef AGPU_OP_DYN2
inline ArrowArrayGPU merge_op_dyn(const ArrowArrayGPU& a, const ArrowArrayGPU& b, const BooleanArrayGPU& mask, ArrowComputePipeline&) { return merge_dyn(a, b, mask); }
inline ArrowArrayGPU take_op_dyn(const ArrowArrayGPU& a, const UInt32ArrayGPU& idx, ArrowComputePipeline&) { return take_dyn(a, idx); }
inline ArrowArrayGPU cast_op_dyn(const ArrowArrayGPU& from, ArrowType into, ArrowComputePipeline&) { return cast_dyn(from, into); }
inline ArrowArrayGPU bitcast_op_dyn(const ArrowArrayGPU& from, ArrowType into, ArrowComputePipeline&) { return bitcast_dyn(from, into); }
inline void put_op_dyn(const ArrowArrayGPU& src, const UInt32ArrayGPU& si, ArrowArrayGPU& dst, const UInt32ArrayGPU& di, ArrowComputePipeline&) { put_dyn(src, si, dst, di); }
inline ArrowArrayGPU broadcast_op_dyn(const ScalarValue& value, size_t len, ArrowComputePipeline& pipeline) { return broadcast_dyn(value, len, pipeline.device); }

// fused expression of BASELINE.json config 3: ((a*b)+c) > d, bit-identical to the unfused chain
inline BooleanArrayGPU fused_mul_add_gt(const Float32ArrayGPU& a, const Float32ArrayGPU& b, const Float32ArrayGPU& c,
                                        const Float32ArrayGPU& d) {
  auto out = BooleanArrayGPU::empty(a.len, a.gpu_device, new_validity(a.gpu_device, a.len, a.null_buffer, b.null_buffer, c.null_buffer, d.null_buffer));
  check(agpu_fused_mul_add_gt(a.gpu_device->handle(), (const float*)a.data->ptr(), (const float*)b.data->ptr(), (const float*)c.data->ptr(),
                              (const float*)d.data->ptr(), (uint32_t*)out.data->ptr(), a.len, vptr(a.null_buffer), vptr(b.null_buffer),
                              vptr(c.null_buffer), vptr(d.null_buffer), vptr_mut(out.null_buffer)), "fused_mul_add_gt");
  return out;
}

// general fused linear chain (agpu_fused_chain): acc = f32(a); each step is unary(op), binary(op,
// column | scalar) or a final compare -> BooleanArrayGPU.  Bit-identical to the unfused ops.
struct ChainStep {
  agpu_chain_step raw{};
  static ChainStep unary(agpu_unop op) { ChainStep s; s.raw.kind = AGPU_STEP_UNARY; s.raw.op = op; return s; }
  static ChainStep binary(agpu_binop op, const Float32ArrayGPU& col) {
    ChainStep s; s.raw.kind = AGPU_STEP_BINARY_COLUMN; s.raw.op = op; s.raw.operand = (const float*)col.data->ptr();
    s.raw.validity = vptr(col.null_buffer); s.has_validity = bool(col.null_buffer); return s;
  }
  static ChainStep binary(agpu_binop op, float scalar) { ChainStep s; s.raw.kind = AGPU_STEP_BINARY_SCALAR; s.raw.op = op; s.raw.scalar = scalar; return s; }
  static ChainStep compare(agpu_cmpop op, const Float32ArrayGPU& col) {
    ChainStep s; s.raw.kind = AGPU_STEP_COMPARE_COLUMN; s.raw.op = op; s.raw.operand = (const float*)col.data->ptr();
    s.raw.validity = vptr(col.null_buffer); s.has_validity = bool(col.null_buffer); return s;
  }
  static ChainStep compare(agpu_cmpop op, float scalar) { ChainStep s; s.raw.kind = AGPU_STEP_COMPARE_SCALAR; s.raw.op = op; s.raw.scalar = scalar; return s; }
  bool has_validity = false;
};

template <typename T>
std::variant<Float32ArrayGPU, BooleanArrayGPU> fused_chain(const PrimitiveArrayGpu<T>& a, const std::vector<ChainStep>& steps) {
  std::vector<agpu_chain_step> raw;
  bool any_validity = bool(a.null_buffer), pred = false;
  for (const auto& s : steps) {
    raw.push_back(s.raw);
    any_validity = any_validity || s.has_validity;
    pred = s.raw.kind == AGPU_STEP_COMPARE_COLUMN || s.raw.kind == AGPU_STEP_COMPARE_SCALAR;
  }
  Validity nb;
  if (any_validity) nb = NullBitBufferGpu{std::make_shared<ArrowGpuBuffer>(a.gpu_device, bitmap_words(a.len) * 4), a.len, a.gpu_device};
  if (pred) {
    auto out = BooleanArrayGPU::empty(a.len, a.gpu_device, nb);
    int rc = agpu_fused_chain(a.gpu_device->handle(), PrimitiveArrayGpu<T>::DTYPE, a.data->ptr(), vptr(a.null_buffer), raw.data(),
                              (int)raw.size(), out.data->ptr(), a.len, vptr_mut(out.null_buffer));
    if (rc == AGPU_EUNSUPPORTED || rc == AGPU_EINVAL) throw Panic("fused_chain: unsupported chain");
    check(rc, "fused_chain");
    return out;
  }
  auto out = Float32ArrayGPU::empty(a.len, a.gpu_device, nb);
  int rc = agpu_fused_chain(a.gpu_device->handle(), PrimitiveArrayGpu<T>::DTYPE, a.data->ptr(), vptr(a.null_buffer), raw.data(),
                            (int)raw.size(), out.data->ptr(), a.len, vptr_mut(out.null_buffer));
  if (rc == AGPU_EUNSUPPORTED || rc == AGPU_EINVAL) throw Panic("fused_chain: unsupported chain");
  check(rc, "fused_chain");
  return out;
}

// A value chain and a predicate chain of ONE source column in one kernel (agpu_fused_chain_pair),
// e.g. s = a + b; g = a > b:  fused_chain_pair(a, {binary(AGPU_ADD, b)}, {compare(AGPU_GT, b)}).
// Both chains must depend on the same validity bitmaps; the two results share one bitmap buffer.
inline std::pair<Float32ArrayGPU, BooleanArrayGPU> fused_chain_pair(const Float32ArrayGPU& a, const std::vector<ChainStep>& value_steps,
                                                                   const std::vector<ChainStep>& pred_steps) {
  auto bitmaps = [&](const std::vector<ChainStep>& steps) {
    std::vector<const uint32_t*> v;
    if (a.null_buffer) v.push_back(vptr(a.null_buffer));
    for (const auto& s : steps)
      if (s.has_validity) v.push_back(s.raw.validity);
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    return v;
  };
  const auto vb = bitmaps(value_steps);
  if (vb != bitmaps(pred_steps)) throw Panic("fused_chain_pair: the two chains depend on different validity bitmaps");
  std::vector<agpu_chain_step> raw;
  for (const auto& s : value_steps) raw.push_back(s.raw);
  agpu_chain_step mark{};
  mark.kind = AGPU_STEP_STORE;
  raw.push_back(mark);
  mark.kind = AGPU_STEP_RESET;
  raw.push_back(mark);
  for (const auto& s : pred_steps) raw.push_back(s.raw);
  Validity nb;
  if (!vb.empty()) nb = NullBitBufferGpu{std::make_shared<ArrowGpuBuffer>(a.gpu_device, bitmap_words(a.len) * 4), a.len, a.gpu_device};
  auto value = Float32ArrayGPU::empty(a.len, a.gpu_device, nb);
  auto pred = BooleanArrayGPU::empty(a.len, a.gpu_device, nb);
  int rc = agpu_fused_chain_pair(a.gpu_device->handle(), AGPU_F32, a.data->ptr(), vptr(a.null_buffer), raw.data(), (int)raw.size(),
                                 (float*)value.data->ptr(), (uint32_t*)pred.data->ptr(), a.len, vptr_mut(value.null_buffer));
  if (rc == AGPU_EUNSUPPORTED || rc == AGPU_EINVAL) throw Panic("fused_chain_pair: unsupported pair of chains");
  check(rc, "fused_chain_pair");
  return {value, pred};
}

// the same chain on an INTEGER column (agpu_fused_chain_int): operands are columns or one-element
// device arrays of the column's own type; no immediates
template <typename T>
struct IntChainStep {
  agpu_chain_step raw{};
  bool has_validity = false;
  static IntChainStep unary(agpu_unop op) { IntChainStep s; s.raw.kind = AGPU_STEP_UNARY; s.raw.op = op; return s; }
  static IntChainStep binary(agpu_binop op, const PrimitiveArrayGpu<T>& o, size_t len) { return with(op, o, len, AGPU_STEP_BINARY_COLUMN, AGPU_STEP_BINARY_DEVSCALAR); }
  static IntChainStep compare(agpu_cmpop op, const PrimitiveArrayGpu<T>& o, size_t len) { return with(op, o, len, AGPU_STEP_COMPARE_COLUMN, AGPU_STEP_COMPARE_DEVSCALAR); }
  static IntChainStep shift(agpu_shiftop op, const UInt32ArrayGPU& counts, size_t len) {  // one per chain
    if (counts.len != len) throw Panic("fused_chain_int: length mismatch");
    IntChainStep s;
    s.raw.kind = AGPU_STEP_SHIFT_COLUMN;
    s.raw.op = op;
    s.raw.operand = counts.data->ptr();
    s.raw.validity = vptr(counts.null_buffer);
    s.has_validity = bool(counts.null_buffer);
    return s;
  }
 private:
  static IntChainStep with(int op, const PrimitiveArrayGpu<T>& o, size_t len, int col_kind, int scalar_kind) {
    IntChainStep s;
    const bool scalar = o.len == 1 && len != 1;  // len-1 operands are scalars, as in add_dyn (arithmetic_kernels.rs:110-117)
    if (!scalar && o.len != len) throw Panic("fused_chain_int: length mismatch");
    s.raw.kind = scalar ? scalar_kind : col_kind;
    s.raw.op = op;
    s.raw.operand = o.data->ptr();
    if (!scalar) { s.raw.validity = vptr(o.null_buffer); s.has_validity = bool(o.null_buffer); }
    return s;
  }
};

template <typename T>
std::variant<PrimitiveArrayGpu<T>, BooleanArrayGPU> fused_chain_int(const PrimitiveArrayGpu<T>& a, const std::vector<IntChainStep<T>>& steps) {
  std::vector<agpu_chain_step> raw;
  bool any_validity = bool(a.null_buffer), pred = false;
  for (const auto& s : steps) {
    raw.push_back(s.raw);
    any_validity = any_validity || s.has_validity;
    pred = s.raw.kind == AGPU_STEP_COMPARE_COLUMN || s.raw.kind == AGPU_STEP_COMPARE_DEVSCALAR;
  }
  Validity nb;
  if (any_validity) nb = NullBitBufferGpu{std::make_shared<ArrowGpuBuffer>(a.gpu_device, bitmap_words(a.len) * 4), a.len, a.gpu_device};
  auto launch = [&](void* out, Validity& v) {
    int rc = agpu_fused_chain_int(a.gpu_device->handle(), PrimitiveArrayGpu<T>::DTYPE, a.data->ptr(), vptr(a.null_buffer), raw.data(),
                                  (int)raw.size(), out, a.len, vptr_mut(v));
    if (rc == AGPU_EUNSUPPORTED || rc == AGPU_EINVAL) throw Panic("fused_chain_int: unsupported chain");
    check(rc, "fused_chain_int");
  };
  if (pred) {
    auto out = BooleanArrayGPU::empty(a.len, a.gpu_device, nb);
    launch(out.data->ptr(), out.null_buffer);
    return out;
  }
  auto out = PrimitiveArrayGpu<T>::empty(a.len, a.gpu_device, nb);
  launch(out.data->ptr(), out.null_buffer);
  return out;
}

}  // namespace arrow_gpu
