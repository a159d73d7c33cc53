// logical.cu — shifts (logical/src/lib.rs:160-186), BooleanArrayGPU bitwise ops
// (logical/src/boolean.rs:45-75) and the stand-alone validity AND
// (array/src/array/null_bit_buffer.rs:168-243).
#include "elementwise.cuh"
#include "ops.cuh"

namespace {

template <template <typename> class F, typename T>
int run_shift(agpu_device* dev, const void* a, const uint32_t* counts, void* out, size_t n, const BmAnd& bm) {
  // granule = 4 rows: counts move as one 16-byte chunk, the shifted column as 4*sizeof(T) bytes
  BinaryOp<T, uint32_t, T, F<T>> op{(const T*)a, counts, (T*)out, F<T>{}};
  return launch_ew(dev, op, n, bm, aligned16(a) && aligned16(counts) && aligned16(out));
}

template <typename T>
int shift_for(agpu_device* dev, int op, const void* a, const uint32_t* counts, void* out, size_t n, const BmAnd& bm) {
  if (op == AGPU_SHL) return run_shift<OpShl, T>(dev, a, counts, out, n, bm);
  if (op == AGPU_SHR) return run_shift<OpShr, T>(dev, a, counts, out, n, bm);
  return AGPU_EUNSUPPORTED;
}

// word-wise bitmap op; the word holding bit n_bits-1 gets its padding bits cleared (Q5)
template <int OP>  // 0 and, 1 or, 2 xor, 3 not
struct BitmapOp {
  static constexpr int G = 4;
  const uint32_t* a;
  const uint32_t* b;
  uint32_t* out;
  size_t last_word;
  uint32_t last_mask;
  struct In { Vec<uint32_t, 4> a, b; };
  __device__ __forceinline__ uint32_t f(uint32_t x, uint32_t y) const {
    return OP == 0 ? (x & y) : OP == 1 ? (x | y) : OP == 2 ? (x ^ y) : ~x;
  }
  __device__ __forceinline__ In load(size_t g) const {
    In in;
    in.a = ld_vec<uint32_t, 4>(a, g);
    if (OP != 3) in.b = ld_vec<uint32_t, 4>(b, g);
    return in;
  }
  __device__ __forceinline__ void run(size_t g, const In& in) const {
    Vec<uint32_t, 4> o;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o.e[k] = f(in.a.e[k], OP != 3 ? in.b.e[k] : 0u);
      if (g * 4 + k == last_word) o.e[k] &= last_mask;
    }
    st_vec<uint32_t, 4>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const {
    uint32_t r = f(a[i], OP != 3 ? b[i] : 0u);
    if (i == last_word) r &= last_mask;
    out[i] = r;
  }
};

template <int OP>
int run_bitmap(agpu_device* dev, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n_bits, const BmAnd& bm_in) {
  const size_t nwords = (n_bits + 31) / 32;
  if (nwords == 0) return 0;
  BitmapOp<OP> op{a, b, out, nwords - 1, (n_bits & 31) ? ((1u << (n_bits & 31)) - 1u) : 0xFFFFFFFFu};
  // the value "rows" of this launch are bitmap words, so the validity words (one per 32 bits)
  // do not line up with the tile; validity is handled by a second tiny launch below
  BmAnd none{};
  int rc = launch_ew(dev, op, nwords, none, aligned16(a) && (OP == 3 || aligned16(b)) && aligned16(out));
  if (rc) return rc;
  if (bm_in.nin) return agpu_launch_bitmap_and(dev, bm_in, n_bits);
  return 0;
}

}  // namespace

// AND (or copy) of validity bitmaps as its own launch
int agpu_launch_bitmap_and(agpu_device* dev, const BmAnd& bm, size_t n_bits) {
  const size_t nwords = (n_bits + 31) / 32;
  if (nwords == 0 || bm.nin == 0) return 0;
  if (bm.nin == 1) {
    AGPU_CUDA(cudaMemcpyAsync(bm.out, bm.in[0], nwords * 4, cudaMemcpyDeviceToDevice, dev->stream));
    return 0;
  }
  int rc = 0;
  const uint32_t* acc = bm.in[0];
  for (int k = 1; k < bm.nin && !rc; ++k) {
    BitmapOp<0> op{acc, bm.in[k], bm.out, (size_t)-1, 0xFFFFFFFFu};
    BmAnd none{};
    rc = launch_ew(dev, op, nwords, none, aligned16(acc) && aligned16(bm.in[k]) && aligned16(bm.out));
    acc = bm.out;
  }
  return rc;
}

extern "C" int agpu_validity_and(agpu_device* dev, const uint32_t* va, const uint32_t* vb, uint32_t* vout,
                                 size_t n_bits) {
  if (!dev) return AGPU_ENODEVICE;
  if (!vout || (!va && !vb)) return AGPU_EINVAL;
  return agpu_launch_bitmap_and(dev, make_bm(va, vb, nullptr, nullptr, vout), n_bits);
}

extern "C" int agpu_shift(agpu_device* dev, int op, int dtype, const void* a, const uint32_t* counts,
                          void* out, size_t n, const uint32_t* va, const uint32_t* vcounts, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n && (!a || !counts || !out)) return AGPU_EINVAL;
  if (vout && !va && !vcounts) return AGPU_EINVAL;
  const BmAnd bm = make_bm(va, vcounts, nullptr, nullptr, vout);
  switch (dtype) {
    case AGPU_I32: return shift_for<int32_t>(dev, op, a, counts, out, n, bm);
    case AGPU_U32: return shift_for<uint32_t>(dev, op, a, counts, out, n, bm);
    case AGPU_I16: return shift_for<int16_t>(dev, op, a, counts, out, n, bm);
    case AGPU_U16: return shift_for<uint16_t>(dev, op, a, counts, out, n, bm);
    case AGPU_I8: return shift_for<int8_t>(dev, op, a, counts, out, n, bm);
    case AGPU_U8: return shift_for<uint8_t>(dev, op, a, counts, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

extern "C" int agpu_bitmap_binary(agpu_device* dev, int op, const uint32_t* a, const uint32_t* b,
                                  uint32_t* out, size_t n_bits, const uint32_t* va, const uint32_t* vb,
                                  uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n_bits && (!a || !b || !out)) return AGPU_EINVAL;
  if (vout && !va && !vb) return AGPU_EINVAL;
  const BmAnd bm = make_bm(va, vb, nullptr, nullptr, vout);
  switch (op) {
    case AGPU_AND: return run_bitmap<0>(dev, a, b, out, n_bits, bm);
    case AGPU_OR: return run_bitmap<1>(dev, a, b, out, n_bits, bm);
    case AGPU_XOR: return run_bitmap<2>(dev, a, b, out, n_bits, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

extern "C" int agpu_bitmap_not(agpu_device* dev, const uint32_t* a, uint32_t* out, size_t n_bits,
                               const uint32_t* va, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n_bits && (!a || !out)) return AGPU_EINVAL;
  if (vout && !va) return AGPU_EINVAL;
  return run_bitmap<3>(dev, a, nullptr, out, n_bits, make_bm(va, nullptr, nullptr, nullptr, vout));
}
