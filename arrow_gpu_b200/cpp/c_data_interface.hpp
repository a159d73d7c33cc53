// c_data_interface.hpp — Arrow C Data Interface import/export for the device arrays
// (the in-memory format adjacent to the path; replaces the reference's
// python_wgarrow stub, crates/python_wgarrow/src/lib.rs:7-11, which exposes dtypes only).
//
// The structs are the ABI-stable ones of the Arrow specification ("The Arrow C data interface").
// import: host Arrow buffers -> H2D (the device layout IS the Arrow layout: dense little-endian
// values, LSB-first validity bitmap); export: D2H into buffers owned by the exported ArrowArray's
// private_data and freed by its release callback.
#pragma once
#include <cstdlib>
#include <cstring>

#include "arrow_gpu.hpp"

extern "C" {
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif
#ifndef ARROW_C_DEVICE_DATA_INTERFACE
#define ARROW_C_DEVICE_DATA_INTERFACE
typedef int32_t ArrowDeviceType;
#define ARROW_DEVICE_CPU 1
#define ARROW_DEVICE_CUDA 2
#define ARROW_DEVICE_CUDA_HOST 3
struct ArrowDeviceArray {
  struct ArrowArray array;
  int64_t device_id;
  ArrowDeviceType device_type;
  void* sync_event;  // CUDA: cudaEvent_t* (an agpu_event* points to exactly that)
  int64_t reserved[3];
};
#endif
}

namespace arrow_gpu {

namespace cdata {
// Arrow format strings of the supported types: c C s S i I f tdD (date32 days) b (bool)
inline const char* format_of(const ArrowArrayGPU& a) {
  static const char* f[] = {"f", "I", "S", "C", "i", "s", "c", "tdD", "b"};  // order of the variant
  return f[a.index()];
}
inline size_t width_of(const char* fmt) {
  const std::string s(fmt);
  if (s == "c" || s == "C") return 1;
  if (s == "s" || s == "S") return 2;
  if (s == "i" || s == "I" || s == "f" || s == "tdD") return 4;
  return 0;
}
// copies `n_bits` bits starting at bit `offset` of `src` into a word-padded bitmap
inline std::vector<uint8_t> rebased_bits(const uint8_t* src, int64_t offset, int64_t n_bits) {
  std::vector<uint8_t> out(bitmap_words((size_t)n_bits) * 4, 0);
  for (int64_t i = 0; i < n_bits; ++i) {
    const int64_t j = offset + i;
    if (src[j / 8] & (1u << (j % 8))) out[i / 8] |= uint8_t(1u << (i % 8));
  }
  return out;
}
struct ExportHolder {
  std::vector<uint8_t> validity, data;
  const void* buffers[2];
};
inline void release_array(ArrowArray* a) {
  delete static_cast<ExportHolder*>(a->private_data);
  a->release = nullptr;
}
inline void release_schema(ArrowSchema* s) {
  std::free(const_cast<char*>(s->format));
  s->release = nullptr;
}
}  // namespace cdata

// Takes ownership semantics of the C data interface: the caller keeps `array`/`schema` alive
// during the call and releases them afterwards (the data is copied to the device).
inline ArrowArrayGPU import_arrow(const ArrowSchema* schema, const ArrowArray* array, const DevicePtr& dev) {
  if (!schema || !array || array->n_children != 0 || array->dictionary) throw ArrowErrorGPU("unsupported Arrow array");
  const std::string fmt(schema->format);
  const size_t n = (size_t)array->length;
  Validity nb;
  if (array->n_buffers >= 1 && array->buffers[0] && array->null_count != 0) {
    auto bits = cdata::rebased_bits(static_cast<const uint8_t*>(array->buffers[0]), array->offset, array->length);
    nb = NullBitBufferGpu{ArrowGpuBuffer::with_data(dev, bits.data(), bits.size()), n, dev};
  }
  if (fmt == "b") {
    auto bits = cdata::rebased_bits(static_cast<const uint8_t*>(array->buffers[1]), array->offset, array->length);
    return BooleanArrayGPU{ArrowGpuBuffer::with_data(dev, bits.data(), bits.size()), dev, n, nb};
  }
  const size_t w = cdata::width_of(schema->format);
  if (!w) throw ArrowErrorGPU("unsupported Arrow format " + fmt);
  const uint8_t* values = static_cast<const uint8_t*>(array->buffers[1]) + (size_t)array->offset * w;
  auto buf = ArrowGpuBuffer::with_data(dev, values, n * w);
  if (fmt == "f") return Float32ArrayGPU{buf, dev, n, nb};
  if (fmt == "I") return UInt32ArrayGPU{buf, dev, n, nb};
  if (fmt == "S") return UInt16ArrayGPU{buf, dev, n, nb};
  if (fmt == "C") return UInt8ArrayGPU{buf, dev, n, nb};
  if (fmt == "i") return Int32ArrayGPU{buf, dev, n, nb};
  if (fmt == "s") return Int16ArrayGPU{buf, dev, n, nb};
  if (fmt == "c") return Int8ArrayGPU{buf, dev, n, nb};
  return Date32ArrayGPU{buf, dev, n, nb};
}

inline void export_arrow(const ArrowArrayGPU& a, ArrowSchema* schema, ArrowArray* array) {
  auto* h = new cdata::ExportHolder();
  size_t n = 0;
  int64_t nulls = 0;
  std::visit([&](const auto& x) {
    using A = std::decay_t<decltype(x)>;
    n = x.len;
    if (x.null_buffer) {
      h->validity = x.null_buffer->bit_buffer->retrive_data(bitmap_words(n) * 4);
      for (size_t i = 0; i < n; ++i) nulls += !(h->validity[i / 8] & (1u << (i % 8)));
    }
    if constexpr (std::is_same<A, BooleanArrayGPU>::value) h->data = x.data->retrive_data(bitmap_words(n) * 4);
    else h->data = x.data->retrive_data(n * sizeof(typename A::Native));
  }, a);
  h->buffers[0] = h->validity.empty() ? nullptr : h->validity.data();
  h->buffers[1] = h->data.data();
  *array = ArrowArray{(int64_t)n, nulls, 0, 2, 0, h->buffers, nullptr, nullptr, cdata::release_array, h};
  const char* fmt = cdata::format_of(a);
  char* owned = static_cast<char*>(std::malloc(std::strlen(fmt) + 1));
  std::strcpy(owned, fmt);
  *schema = ArrowSchema{owned, "", nullptr, 2 /* ARROW_FLAG_NULLABLE */, 0, nullptr, nullptr, cdata::release_schema, nullptr};
}

// ------------------------------------------------------------------------------------------
// Arrow C Device Data Interface: ZERO-COPY exchange of device-resident columns (the Python side is
// arrow_gpu_b200/c_device.py; both follow the specification's ArrowDeviceArray).
//   export: the array's own device pointers; the structure keeps the buffers (and the event
//           recorded on the producing stream) alive until the consumer calls release;
//   import: the producer's pointers wrapped in non-owning buffers; the consuming stream waits for
//           sync_event; release is called when the last imported buffer goes, after a sync of the
//           consuming handle (its kernels may still be reading).
namespace cdata {
struct DeviceExportHolder {
  BufferPtr validity, data;
  agpu_event* event = nullptr;
  const void* buffers[2];
};
inline void release_device_array(ArrowArray* a) {
  auto* h = static_cast<DeviceExportHolder*>(a->private_data);
  if (h->event) agpu_event_destroy(h->event);
  delete h;
  a->release = nullptr;
}
// the moved structure of an imported array; dropped with the last buffer that refers to it
struct ImportedDeviceArray {
  ArrowDeviceArray moved;
  DevicePtr dev;
  ~ImportedDeviceArray() {
    if (moved.array.release) {
      agpu_sync(dev->handle());
      moved.array.release(&moved.array);
    }
  }
};
}  // namespace cdata

inline void export_arrow_device(const ArrowArrayGPU& a, ArrowSchema* schema, ArrowDeviceArray* out) {
  auto* h = new cdata::DeviceExportHolder();
  size_t n = 0;
  DevicePtr dev;
  std::visit([&](const auto& x) {
    n = x.len;
    dev = x.gpu_device;
    h->data = x.data;
    if (x.null_buffer) h->validity = x.null_buffer->bit_buffer;
  }, a);
  check(agpu_event_create(&h->event), "export_arrow_device");
  check(agpu_event_record(dev->handle(), h->event), "export_arrow_device");
  h->buffers[0] = h->validity ? h->validity->ptr() : nullptr;
  h->buffers[1] = h->data->ptr();
  std::memset(out, 0, sizeof(*out));
  out->array = ArrowArray{(int64_t)n, h->validity ? -1 : 0, 0, 2, 0, h->buffers, nullptr, nullptr, cdata::release_device_array, h};
  out->device_id = dev->ordinal();
  out->device_type = ARROW_DEVICE_CUDA;
  out->sync_event = h->event;
  const char* fmt = cdata::format_of(a);
  char* owned = static_cast<char*>(std::malloc(std::strlen(fmt) + 1));
  std::strcpy(owned, fmt);
  *schema = ArrowSchema{owned, "", nullptr, 2 /* ARROW_FLAG_NULLABLE */, 0, nullptr, nullptr, cdata::release_schema, nullptr};
}

// MOVES `in` (its release pointer is cleared, as the specification describes for moving an array)
inline ArrowArrayGPU import_arrow_device(const ArrowSchema* schema, ArrowDeviceArray* in, const DevicePtr& dev) {
  if (!schema || !in || !in->array.release) throw ArrowErrorGPU("released or missing ArrowDeviceArray");
  if (in->array.n_children != 0 || in->array.dictionary || in->array.n_buffers != 2) throw ArrowErrorGPU("unsupported Arrow array");
  if (in->device_type != ARROW_DEVICE_CUDA) throw ArrowErrorGPU("import_arrow_device takes CUDA memory; use import_arrow for host memory");
  if (in->device_id != dev->ordinal()) throw ArrowErrorGPU("the column lives on another CUDA device");
  const std::string fmt(schema->format);
  const size_t n = (size_t)in->array.length;
  const int64_t off = in->array.offset;
  const bool is_bool = fmt == "b";
  const size_t w = cdata::width_of(schema->format);
  if (!is_bool && !w) throw ArrowErrorGPU("unsupported Arrow format " + fmt);
  const bool has_nulls = in->array.buffers[0] && in->array.null_count != 0;
  if ((has_nulls || is_bool) && off % 32) throw ArrowErrorGPU("zero-copy import needs a bitmap offset that is a multiple of 32 rows");
  auto owner = std::make_shared<cdata::ImportedDeviceArray>();
  owner->moved = *in;
  owner->dev = dev;
  in->array.release = nullptr;
  if (in->sync_event) check(agpu_stream_wait_event(dev->handle(), static_cast<agpu_event*>(in->sync_event)), "import_arrow_device");
  auto wrap = [&](const void* base, size_t byte_off, size_t bytes) {
    return std::make_shared<ArrowGpuBuffer>(dev, const_cast<uint8_t*>(static_cast<const uint8_t*>(base)) + byte_off, bytes, owner);
  };
  Validity nb;
  if (has_nulls) nb = NullBitBufferGpu{wrap(owner->moved.array.buffers[0], (size_t)off / 8, bitmap_words(n) * 4), n, dev};
  if (is_bool) return BooleanArrayGPU{wrap(owner->moved.array.buffers[1], (size_t)off / 8, bitmap_words(n) * 4), dev, n, nb};
  auto buf = wrap(owner->moved.array.buffers[1], (size_t)off * w, n * w);
  if (fmt == "f") return Float32ArrayGPU{buf, dev, n, nb};
  if (fmt == "I") return UInt32ArrayGPU{buf, dev, n, nb};
  if (fmt == "S") return UInt16ArrayGPU{buf, dev, n, nb};
  if (fmt == "C") return UInt8ArrayGPU{buf, dev, n, nb};
  if (fmt == "i") return Int32ArrayGPU{buf, dev, n, nb};
  if (fmt == "s") return Int16ArrayGPU{buf, dev, n, nb};
  if (fmt == "c") return Int8ArrayGPU{buf, dev, n, nb};
  return Date32ArrayGPU{buf, dev, n, nb};
}

}  // namespace arrow_gpu
