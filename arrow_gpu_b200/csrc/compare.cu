// compare.cu — agpu_compare (compare/src/lib.rs:85-111,142-162) and the fused f32 expression
// ((a*b)+c) > d of BASELINE.json config 3.
#include <type_traits>

#include "bits.cuh"
#include "ops.cuh"

namespace {


// f32 compares are IEEE (any NaN -> false, -0 == +0); integers compare with their own
// signedness — native sub-word lanes instead of the reference's get_*_byte/get_*_half helpers.
template <typename T, class P>
struct CmpOp {
  static constexpr int G = 16 / sizeof(T);
  const T* a;
  const T* b;
  struct In { Vec<T, G> a, b; };
  __device__ __forceinline__ In load(size_t g) const { return In{ld_vec<T, G>(a, g), ld_vec<T, G>(b, g)}; }
  __device__ __forceinline__ uint32_t bits(size_t, const In& in) const {
    uint32_t m = 0;
    if constexpr (sizeof(T) < 4) {
      // one lane mask per packed word, then the mask's lane bits are gathered with a multiply:
      // bytes: (m & 0x08040201) * 0x01010101 >> 24 = 4 bits; halves: (m & 0x00020001) * 0x00010001 >> 16 = 2 bits
      uint32_t wa[4], wb[4];
      memcpy(wa, &in.a, 16);
      memcpy(wb, &in.b, 16);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t lanes = P::template lanes<T>(wa[k], wb[k]);
        if constexpr (sizeof(T) == 1) m |= (((lanes & 0x08040201u) * 0x01010101u) >> 24) << (4 * k);
        else m |= ((((lanes & 0x00020001u) * 0x00010001u) >> 16) & 3u) << (2 * k);
      }
    } else {
#pragma unroll
      for (int k = 0; k < G; ++k) m |= (uint32_t)P{}(in.a.e[k], in.b.e[k]) << k;
    }
    return m;
  }
  __device__ __forceinline__ bool bit_at(size_t i) const { return P{}(a[i], b[i]); }
};

template <typename T>
int compare_for(agpu_device* dev, int op, const void* a, const void* b, uint32_t* out, size_t n, const BmAnd& bm) {
  const bool al = aligned16(a) && aligned16(b);
  const T* x = (const T*)a;
  const T* y = (const T*)b;
  switch (op) {
    case AGPU_GT: return launch_bits(dev, CmpOp<T, PGt>{x, y}, out, n, bm, al);
    case AGPU_GTEQ: return launch_bits(dev, CmpOp<T, PGe>{x, y}, out, n, bm, al);
    case AGPU_LT: return launch_bits(dev, CmpOp<T, PLt>{x, y}, out, n, bm, al);
    case AGPU_LTEQ: return launch_bits(dev, CmpOp<T, PLe>{x, y}, out, n, bm, al);
    case AGPU_EQ: return launch_bits(dev, CmpOp<T, PEq>{x, y}, out, n, bm, al);
    default: return AGPU_EUNSUPPORTED;
  }
}

// ((a*b)+c) > d with the two roundings of the unfused mul_op -> add_op chain (no FMA).
struct FusedMulAddGt {
  static constexpr int G = 4;
  const float *a, *b, *c, *d;
  struct In { Vec<float, 4> a, b, c, d; };
  __device__ __forceinline__ In load(size_t g) const {
    return In{ld_vec<float, 4>(a, g), ld_vec<float, 4>(b, g), ld_vec<float, 4>(c, g), ld_vec<float, 4>(d, g)};
  }
  __device__ __forceinline__ uint32_t bits(size_t, const In& in) const {
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      m |= (uint32_t)(__fadd_rn(__fmul_rn(in.a.e[k], in.b.e[k]), in.c.e[k]) > in.d.e[k]) << k;
    return m;
  }
  __device__ __forceinline__ bool bit_at(size_t i) const { return __fadd_rn(__fmul_rn(a[i], b[i]), c[i]) > d[i]; }
};

}  // namespace

extern "C" int agpu_compare(agpu_device* dev, int op, int dtype, const void* a, const void* b,
                            uint32_t* out_bits, size_t n, const uint32_t* va, const uint32_t* vb,
                            uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n && (!a || !b || !out_bits)) return AGPU_EINVAL;
  if (vout && !va && !vb) return AGPU_EINVAL;
  const BmAnd bm = make_bm(va, vb, nullptr, nullptr, vout);
  switch (dtype) {
    case AGPU_F32: return compare_for<float>(dev, op, a, b, out_bits, n, bm);
    case AGPU_I32: case AGPU_DATE32: return compare_for<int32_t>(dev, op, a, b, out_bits, n, bm);
    case AGPU_U32: return compare_for<uint32_t>(dev, op, a, b, out_bits, n, bm);
    case AGPU_I16: return compare_for<int16_t>(dev, op, a, b, out_bits, n, bm);
    case AGPU_U16: return compare_for<uint16_t>(dev, op, a, b, out_bits, n, bm);
    case AGPU_I8: return compare_for<int8_t>(dev, op, a, b, out_bits, n, bm);
    case AGPU_U8: return compare_for<uint8_t>(dev, op, a, b, out_bits, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

extern "C" int agpu_fused_mul_add_gt(agpu_device* dev, const float* a, const float* b, const float* c,
                                     const float* d, uint32_t* out_bits, size_t n, const uint32_t* va,
                                     const uint32_t* vb, const uint32_t* vc, const uint32_t* vd,
                                     uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n && (!a || !b || !c || !d || !out_bits)) return AGPU_EINVAL;
  if (vout && !va && !vb && !vc && !vd) return AGPU_EINVAL;
  const BmAnd bm = make_bm(va, vb, vc, vd, vout);
  const bool al = aligned16(a) && aligned16(b) && aligned16(c) && aligned16(d);
  return launch_bits<FusedMulAddGt, 2>(dev, FusedMulAddGt{a, b, c, d}, out_bits, n, bm, al);
}
