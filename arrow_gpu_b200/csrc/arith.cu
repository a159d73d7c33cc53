// arith.cu — agpu_binary / agpu_scalar: arithmetic, min/max, bitwise logical and power.
#include <stdlib.h>

#include "elementwise.cuh"
#include "ops.cuh"

namespace {

template <template <typename> class F, typename T>
int run_binary(agpu_device* dev, const void* a, const void* b, void* out, size_t n, const BmAnd& bm) {
  BinaryOp<T, T, T, F<T>> op{(const T*)a, (const T*)b, (T*)out, F<T>{}};
  return launch_ew(dev, op, n, bm, aligned16(a) && aligned16(b) && aligned16(out));
}

template <template <typename> class F, typename T>
int run_scalar(agpu_device* dev, const void* a, const void* s, void* out, size_t n, const BmAnd& bm) {
  ScalarOp<T, F<T>> op{(const T*)a, (const T*)s, (T*)out, F<T>{}};
  // 2 B/row ops (1-byte column, scalar rhs): 2 granules per thread measured 1 % faster than 4 and 2 %
  // faster than 8 (84 us kernels; profiles/r02_persistent_ab.md)
  constexpr int UNROLL = sizeof(T) == 1 ? 2 : 4;
  return launch_ew<ScalarOp<T, F<T>>, UNROLL>(dev, op, n, bm, aligned16(a) && aligned16(out));
}

template <typename T>
int binary_int(agpu_device* dev, int op, const void* a, const void* b, void* out, size_t n, const BmAnd& bm) {
  switch (op) {
    case AGPU_ADD: return run_binary<OpAdd, T>(dev, a, b, out, n, bm);
    case AGPU_SUB: return run_binary<OpSub, T>(dev, a, b, out, n, bm);
    case AGPU_MUL: return run_binary<OpMul, T>(dev, a, b, out, n, bm);
    case AGPU_DIV: return run_binary<OpDiv, T>(dev, a, b, out, n, bm);
    case AGPU_REM: return run_binary<OpRem, T>(dev, a, b, out, n, bm);
    case AGPU_MIN: return run_binary<OpMin, T>(dev, a, b, out, n, bm);
    case AGPU_MAX: return run_binary<OpMax, T>(dev, a, b, out, n, bm);
    case AGPU_AND: return run_binary<OpAnd, T>(dev, a, b, out, n, bm);
    case AGPU_OR: return run_binary<OpOr, T>(dev, a, b, out, n, bm);
    case AGPU_XOR: return run_binary<OpXor, T>(dev, a, b, out, n, bm);
    case AGPU_POW:
      if constexpr (std::is_same<T, int32_t>::value) return run_binary<OpPow, T>(dev, a, b, out, n, bm);
      return AGPU_EUNSUPPORTED;
    default: return AGPU_EUNSUPPORTED;
  }
}

int binary_f32(agpu_device* dev, int op, const void* a, const void* b, void* out, size_t n, const BmAnd& bm) {
  switch (op) {
    case AGPU_ADD: return run_binary<OpAdd, float>(dev, a, b, out, n, bm);
    case AGPU_SUB: return run_binary<OpSub, float>(dev, a, b, out, n, bm);
    case AGPU_MUL: return run_binary<OpMul, float>(dev, a, b, out, n, bm);
    case AGPU_DIV: return run_binary<OpDiv, float>(dev, a, b, out, n, bm);
    case AGPU_REM: return run_binary<OpRem, float>(dev, a, b, out, n, bm);
    case AGPU_MIN: return run_binary<OpMin, float>(dev, a, b, out, n, bm);
    case AGPU_MAX: return run_binary<OpMax, float>(dev, a, b, out, n, bm);
    case AGPU_POW: return run_binary<OpPow, float>(dev, a, b, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

template <typename T>
int scalar_any(agpu_device* dev, int op, const void* a, const void* s, void* out, size_t n, const BmAnd& bm) {
  switch (op) {
    case AGPU_ADD: return run_scalar<OpAdd, T>(dev, a, s, out, n, bm);
    case AGPU_SUB: return run_scalar<OpSub, T>(dev, a, s, out, n, bm);
    case AGPU_MUL: return run_scalar<OpMul, T>(dev, a, s, out, n, bm);
    case AGPU_DIV: return run_scalar<OpDiv, T>(dev, a, s, out, n, bm);
    case AGPU_REM: return run_scalar<OpRem, T>(dev, a, s, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

}  // namespace

extern "C" int agpu_binary(agpu_device* dev, int op, int dtype, const void* a, const void* b, void* out,
                           size_t n, const uint32_t* va, const uint32_t* vb, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n && (!a || !b || !out)) return AGPU_EINVAL;
  if (vout && !va && !vb) return AGPU_EINVAL;
  const BmAnd bm = make_bm(va, vb, nullptr, nullptr, vout);
  switch (dtype) {
    case AGPU_F32: return binary_f32(dev, op, a, b, out, n, bm);
    case AGPU_I32: case AGPU_DATE32: return binary_int<int32_t>(dev, op, a, b, out, n, bm);
    case AGPU_U32: return binary_int<uint32_t>(dev, op, a, b, out, n, bm);
    case AGPU_I16: return binary_int<int16_t>(dev, op, a, b, out, n, bm);
    case AGPU_U16: return binary_int<uint16_t>(dev, op, a, b, out, n, bm);
    case AGPU_I8: return binary_int<int8_t>(dev, op, a, b, out, n, bm);
    case AGPU_U8: return binary_int<uint8_t>(dev, op, a, b, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

extern "C" int agpu_scalar(agpu_device* dev, int op, int dtype, const void* a, const void* scalar_dev,
                           void* out, size_t n, const uint32_t* va, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (!scalar_dev || (n && (!a || !out))) return AGPU_EINVAL;
  if (vout && !va) return AGPU_EINVAL;
  const BmAnd bm = make_bm(va, nullptr, nullptr, nullptr, vout);
  switch (dtype) {
    case AGPU_F32: return scalar_any<float>(dev, op, a, scalar_dev, out, n, bm);
    case AGPU_I32: case AGPU_DATE32: return scalar_any<int32_t>(dev, op, a, scalar_dev, out, n, bm);
    case AGPU_U32: return scalar_any<uint32_t>(dev, op, a, scalar_dev, out, n, bm);
    case AGPU_I16: return scalar_any<int16_t>(dev, op, a, scalar_dev, out, n, bm);
    case AGPU_U16: return scalar_any<uint16_t>(dev, op, a, scalar_dev, out, n, bm);
    case AGPU_I8: return scalar_any<int8_t>(dev, op, a, scalar_dev, out, n, bm);
    case AGPU_U8: return scalar_any<uint8_t>(dev, op, a, scalar_dev, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}
