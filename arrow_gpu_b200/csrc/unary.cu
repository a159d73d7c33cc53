// unary.cu — agpu_unary: neg / abs / not, f32 math, trig (with the int->f32 cast fused).
#include <stdlib.h>

#include "elementwise.cuh"
#include "ops.cuh"

namespace {

template <typename TI, typename TO, class F>
int run_unary(agpu_device* dev, const void* a, void* out, size_t n, const BmAnd& bm) {
  UnaryOp<TI, TO, F> op{(const TI*)a, (TO*)out, F{}};
  constexpr int UNROLL = (sizeof(TI) == 1 && sizeof(TO) == 1) ? 2 : 4;  // 2 B/row ops: see arith.cu run_scalar
  return launch_ew<UnaryOp<TI, TO, F>, UNROLL>(dev, op, n, bm, aligned16(a) && aligned16(out));
}

// 8-bit column -> f32 function value.  An 8-bit input has 256 possible values, so each CTA
// evaluates the function once per value into a 1 KiB shared-memory table (one entry per thread)
// and the rows become table look-ups: identical bits to evaluating F per row, at a fraction of
// the instructions (sinf per row made the i8 path issue-bound at 0.81 of the roofline).
template <typename TI, class F, int UNROLL>
__global__ void __launch_bounds__(kBlock) lut8_kernel(const TI* __restrict__ a, float* __restrict__ out, const size_t n,
                                                      const BmAnd bm) {
  static_assert(sizeof(TI) == 1 && kBlock == 256, "one table entry per thread");
  __shared__ float lut[256];
  lut[threadIdx.x] = F{}((TI)(uint8_t)threadIdx.x);  // indexed by the raw byte
  __syncthreads();
  constexpr int G = 4;  // 4 rows: one 4-byte chunk in, one 16-byte chunk out
  const size_t n_gran = n / G;
  const size_t tile_gran = (size_t)kBlock * UNROLL;
  const size_t g0 = (size_t)blockIdx.x * tile_gran + threadIdx.x;
  auto emit = [&](size_t g, const Vec<TI, G>& v) {
    Vec<float, G> o;
#pragma unroll
    for (int k = 0; k < G; ++k) o.e[k] = lut[(uint8_t)v.e[k]];
    st_vec<float, G>(out, g, o);
  };
  if (((size_t)blockIdx.x + 1) * tile_gran <= n_gran) {
    Vec<TI, G> in[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) in[j] = ld_vec<TI, G>(a, g0 + (size_t)j * kBlock);
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) emit(g0 + (size_t)j * kBlock, in[j]);
  } else {
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const size_t g = g0 + (size_t)j * kBlock;
      if (g < n_gran) emit(g, ld_vec<TI, G>(a, g));
    }
    if (blockIdx.x == gridDim.x - 1) {
      const size_t i = n_gran * G + threadIdx.x;
      if (i < n) out[i] = lut[(uint8_t)a[i]];
    }
  }
  constexpr int tile_words = kBlock * UNROLL * G / 32;
  bm.tile((size_t)blockIdx.x * tile_words, tile_words, (n + 31) / 32);
}

template <typename TI, class F>
int run_unary8(agpu_device* dev, const void* a, void* out, size_t n, const BmAnd& bm) {
  if (n == 0) return 0;
  if (!(aligned16(a) && aligned16(out))) return run_unary<TI, float, F>(dev, a, out, n, bm);
  constexpr int UNROLL = 8;
  const size_t grid = ceil_div(n, (size_t)kBlock * UNROLL * 4);
  if (grid > 0x7FFFFFFFull) return AGPU_EINVAL;
  AGPU_LAUNCH(dev, (lut8_kernel<TI, F, UNROLL>), (unsigned)grid, kBlock, 0, (const TI*)a, (float*)out, n, bm);
  return agpu_finish_launch();
}

// sin / cos / sinh on an integer column -> f32 column
template <typename TI>
int trig_int(agpu_device* dev, int op, const void* a, void* out, size_t n, const BmAnd& bm) {
  if constexpr (sizeof(TI) == 1) {
    switch (op) {
      case AGPU_SIN: return run_unary8<TI, FSin<TI>>(dev, a, out, n, bm);
      case AGPU_COS: return run_unary8<TI, FCos<TI>>(dev, a, out, n, bm);
      case AGPU_SINH: return run_unary8<TI, FSinh<TI>>(dev, a, out, n, bm);
      default: return AGPU_EUNSUPPORTED;
    }
  }
  switch (op) {
    case AGPU_SIN: return run_unary<TI, float, FSin<TI>>(dev, a, out, n, bm);
    case AGPU_COS: return run_unary<TI, float, FCos<TI>>(dev, a, out, n, bm);
    case AGPU_SINH: return run_unary<TI, float, FSinh<TI>>(dev, a, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

template <typename T>
int not_int(agpu_device* dev, const void* a, void* out, size_t n, const BmAnd& bm) {
  return run_unary<T, T, OpNot<T>>(dev, a, out, n, bm);
}

}  // namespace

extern "C" int agpu_unary(agpu_device* dev, int op, int dtype, const void* a, void* out, size_t n,
                          const uint32_t* va, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n && (!a || !out)) return AGPU_EINVAL;
  if (vout && !va) return AGPU_EINVAL;
  const BmAnd bm = make_bm(va, nullptr, nullptr, nullptr, vout);
  if (op == AGPU_NOT) {
    switch (dtype) {
      case AGPU_I32: return not_int<int32_t>(dev, a, out, n, bm);
      case AGPU_U32: return not_int<uint32_t>(dev, a, out, n, bm);
      case AGPU_I16: return not_int<int16_t>(dev, a, out, n, bm);
      case AGPU_U16: return not_int<uint16_t>(dev, a, out, n, bm);
      case AGPU_I8: return not_int<int8_t>(dev, a, out, n, bm);
      case AGPU_U8: return not_int<uint8_t>(dev, a, out, n, bm);
      default: return AGPU_EUNSUPPORTED;
    }
  }
  switch (dtype) {
    case AGPU_F32:
      switch (op) {
        case AGPU_NEG: return run_unary<float, float, OpNeg<float>>(dev, a, out, n, bm);
        case AGPU_ABS: return run_unary<float, float, OpAbs<float>>(dev, a, out, n, bm);
        case AGPU_SQRT: return run_unary<float, float, FSqrt<float>>(dev, a, out, n, bm);
        case AGPU_CBRT: return run_unary<float, float, FCbrt<float>>(dev, a, out, n, bm);
        case AGPU_EXP: return run_unary<float, float, FExp<float>>(dev, a, out, n, bm);
        case AGPU_EXP2: return run_unary<float, float, FExp2<float>>(dev, a, out, n, bm);
        case AGPU_LOG: return run_unary<float, float, FLog<float>>(dev, a, out, n, bm);
        case AGPU_LOG2: return run_unary<float, float, FLog2<float>>(dev, a, out, n, bm);
        case AGPU_SIN: return run_unary<float, float, FSin<float>>(dev, a, out, n, bm);
        case AGPU_COS: return run_unary<float, float, FCos<float>>(dev, a, out, n, bm);
        case AGPU_ACOS: return run_unary<float, float, FAcos<float>>(dev, a, out, n, bm);
        case AGPU_SINH: return run_unary<float, float, FSinh<float>>(dev, a, out, n, bm);
        default: return AGPU_EUNSUPPORTED;
      }
    case AGPU_I32:
      if (op == AGPU_ABS) return run_unary<int32_t, int32_t, OpAbs<int32_t>>(dev, a, out, n, bm);
      return AGPU_EUNSUPPORTED;
    case AGPU_I16: return trig_int<int16_t>(dev, op, a, out, n, bm);
    case AGPU_U16: return trig_int<uint16_t>(dev, op, a, out, n, bm);
    case AGPU_I8: return trig_int<int8_t>(dev, op, a, out, n, bm);
    case AGPU_U8: return trig_int<uint8_t>(dev, op, a, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}
