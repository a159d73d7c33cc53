// cast.cu — agpu_cast: the reference's cast matrix (cast/src/lib.rs:135-161).
#include "elementwise.cuh"
#include "ops.cuh"

namespace {

template <typename TI, typename TO>
int run_cast(agpu_device* dev, const void* a, void* out, size_t n, const BmAnd& bm) {
  // granule = 16 / max(sizeof TI, sizeof TO) rows: the wide side moves as 16-byte chunks, the
  // narrow side as 4- or 8-byte chunks of the same rows — both fully coalesced.  More granules
  // per thread than usual because the narrow loads carry few bytes each.
  UnaryOp<TI, TO, OpCast<TI, TO>> op{(const TI*)a, (TO*)out, OpCast<TI, TO>{}};
  return launch_ew<decltype(op), 8>(dev, op, n, bm, aligned16(a) && aligned16(out));
}

// bool -> f32: cast/compute_shaders/boolean/cast_f32.wgsl:9-20.  Granule = 4 rows (one 16-byte
// f32 chunk); the 8 lanes that share a bitmap word read it through the read-only cache.
struct BoolToF32 {
  static constexpr int G = 4;
  const uint32_t* bits;
  float* out;
  struct In { uint32_t nib; };
  __device__ __forceinline__ In load(size_t g) const {
    return In{(__ldg(bits + (g >> 3)) >> ((g & 7) * 4)) & 0xFu};
  }
  __device__ __forceinline__ void run(size_t g, const In& in) const {
    Vec<float, 4> o;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.e[k] = (in.nib >> k) & 1u ? 1.0f : 0.0f;
    st_vec<float, 4>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const { out[i] = (bits[i >> 5] >> (i & 31)) & 1u ? 1.0f : 0.0f; }
};

int copy_cast(agpu_device* dev, const void* a, void* out, size_t bytes, const BmAnd& bm, size_t n) {
  // same-width signed<->unsigned: the reference clones the buffer (cast/src/lib.rs:69-86)
  if (bytes) AGPU_CUDA(cudaMemcpyAsync(out, a, bytes, cudaMemcpyDeviceToDevice, dev->stream));
  if (bm.nin) return agpu_launch_bitmap_and(dev, bm, n);
  return 0;
}

}  // namespace

extern "C" int agpu_cast(agpu_device* dev, int src, int dst, const void* a, void* out, size_t n,
                         const uint32_t* va, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n && (!a || !out)) return AGPU_EINVAL;
  if (vout && !va) return AGPU_EINVAL;
  const BmAnd bm = make_bm(va, nullptr, nullptr, nullptr, vout);
#define CASE(S, D, TS, TD) \
  if (src == S && dst == D) return run_cast<TS, TD>(dev, a, out, n, bm);
  CASE(AGPU_I8, AGPU_I16, int8_t, int16_t)
  CASE(AGPU_I8, AGPU_I32, int8_t, int32_t)
  CASE(AGPU_I8, AGPU_U16, int8_t, uint16_t)
  CASE(AGPU_I8, AGPU_U32, int8_t, uint32_t)
  CASE(AGPU_I8, AGPU_F32, int8_t, float)
  CASE(AGPU_I16, AGPU_I32, int16_t, int32_t)
  CASE(AGPU_I16, AGPU_U32, int16_t, uint32_t)
  CASE(AGPU_I16, AGPU_F32, int16_t, float)
  CASE(AGPU_U8, AGPU_U16, uint8_t, uint16_t)
  CASE(AGPU_U8, AGPU_U32, uint8_t, uint32_t)
  CASE(AGPU_U8, AGPU_I16, uint8_t, int16_t)
  CASE(AGPU_U8, AGPU_I32, uint8_t, int32_t)
  CASE(AGPU_U8, AGPU_F32, uint8_t, float)
  CASE(AGPU_U16, AGPU_U32, uint16_t, uint32_t)
  CASE(AGPU_U16, AGPU_I32, uint16_t, int32_t)
  CASE(AGPU_U16, AGPU_F32, uint16_t, float)
  CASE(AGPU_F32, AGPU_U8, float, uint8_t)
#undef CASE
  if ((src == AGPU_I8 && dst == AGPU_U8) || (src == AGPU_U8 && dst == AGPU_I8)) return copy_cast(dev, a, out, n, bm, n);
  if ((src == AGPU_I16 && dst == AGPU_U16) || (src == AGPU_U16 && dst == AGPU_I16)) return copy_cast(dev, a, out, n * 2, bm, n);
  if (src == AGPU_U32 && dst == AGPU_F32) return copy_cast(dev, a, out, n * 4, bm, n);  // bitcast (lib.rs:90-108)
  if (src == AGPU_BOOL && dst == AGPU_F32) {
    BoolToF32 op{(const uint32_t*)a, (float*)out};
    return launch_ew<BoolToF32, 8>(dev, op, n, bm, aligned16(out));
  }
  return AGPU_EUNSUPPORTED;
}
