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

}  // namespace arrow_gpu
