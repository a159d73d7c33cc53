// chain_int.cu — agpu_fused_chain_int: the chain interpreter of chain.cu on integer columns.
//
// Every step applies the same functor as the stand-alone integer kernel of that op (ops.cuh:
// two's-complement wrap in the column's own width, the WGSL divide/remainder-by-zero rules,
// signedness of min/max/compare), so a fused chain is bit-identical to the ops run one by one.
// Rows move as 16-byte granules of the column type (16 x i8, 8 x i16, 4 x i32) and STAY packed:
// the running value of a granule is four 32-bit words, narrow lanes are processed SIMD-in-word
// where the op allows it.  Steps consume the operand columns in order, so the "next column" is
// always slot 0 and the slots rotate.
#include "bits.cuh"
#include "elementwise.cuh"
#include "ops.cuh"

namespace {

constexpr int kMaxCols = 3;

struct IntChainProgram {
  int n_steps;
  int n_cols;
  int kind[AGPU_CHAIN_MAX_STEPS];
  int op[AGPU_CHAIN_MAX_STEPS];
  const void* dscalar[AGPU_CHAIN_MAX_STEPS];  // one-element device arrays of the column type
  const void* cols[kMaxCols];
  const uint32_t* counts;                     // the u32 per-row counts of the chain's shift step, or NULL
};

// ---- packed words: a 16-byte granule is four 32-bit words of L = 4/sizeof(T) lanes each --------
// add/sub/min/max/compare run on whole words (SIMD-in-word), and/or/xor/not are plain word ops;
// mul/div/rem/pow unpack the lanes, apply the stand-alone functor and repack.
template <typename T, template <typename> class F>
__device__ __forceinline__ uint32_t lanewise(uint32_t a, uint32_t b) {
  constexpr int L = 4 / sizeof(T), B = 8 * sizeof(T);
  if constexpr (L == 1) {
    return (uint32_t)F<T>{}((T)a, (T)b);
  } else {
    using UT = typename std::make_unsigned<T>::type;
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < L; ++k)
      r |= (uint32_t)(UT)F<T>{}((T)(UT)(a >> (B * k)), (T)(UT)(b >> (B * k))) << (B * k);
    return r;
  }
}
template <typename T> __device__ __forceinline__ uint32_t w_add(uint32_t a, uint32_t b) {
  if constexpr (sizeof(T) == 1) return __vadd4(a, b);
  else if constexpr (sizeof(T) == 2) return __vadd2(a, b);
  else return a + b;
}
template <typename T> __device__ __forceinline__ uint32_t w_sub(uint32_t a, uint32_t b) {
  if constexpr (sizeof(T) == 1) return __vsub4(a, b);
  else if constexpr (sizeof(T) == 2) return __vsub2(a, b);
  else return a - b;
}
template <typename T> __device__ __forceinline__ uint32_t w_min(uint32_t a, uint32_t b) {
  if constexpr (sizeof(T) < 4) return OpMin<T>::word(a, b);
  else return (uint32_t)OpMin<T>{}((T)a, (T)b);
}
template <typename T> __device__ __forceinline__ uint32_t w_max(uint32_t a, uint32_t b) {
  if constexpr (sizeof(T) < 4) return OpMax<T>::word(a, b);
  else return (uint32_t)OpMax<T>{}((T)a, (T)b);
}
template <typename T> __device__ __forceinline__ uint32_t w_splat(T v) {  // the scalar in every lane
  using UT = typename std::make_unsigned<T>::type;
  if constexpr (sizeof(T) == 1) return (uint32_t)(UT)v * 0x01010101u;
  else if constexpr (sizeof(T) == 2) return (uint32_t)(UT)v * 0x00010001u;
  else return (uint32_t)v;
}

// shift the L lanes of one word by their own counts (logical/compute_shaders/*/shift.wgsl: widen,
// shift by count & 31, truncate) with the stand-alone functors
template <typename T, bool LEFT>
__device__ __forceinline__ uint32_t shift_lanes(uint32_t a, const uint32_t* c) {
  constexpr int L = 4 / sizeof(T), B = 8 * sizeof(T);
  using UT = typename std::make_unsigned<T>::type;
  uint32_t r = 0;
#pragma unroll
  for (int k = 0; k < L; ++k) {
    const T x = (T)(UT)(a >> (B * k));
    const T y = LEFT ? OpShl<T>{}(x, c[k]) : OpShr<T>{}(x, c[k]);
    r |= (uint32_t)(UT)y << (B * k);
  }
  return r;
}

template <typename T, int N>
__device__ __forceinline__ void words_unary(int op, uint32_t (&a)[N]) {
  if (op == AGPU_NOT) {
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] = ~a[k];
  } else if constexpr (std::is_same<T, int32_t>::value) {  // AGPU_ABS (int32 only, checked on the host)
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] = (uint32_t)OpAbs<T>{}((T)a[k]);
  }
}

template <typename T, int N>
__device__ __forceinline__ void words_binary(int op, uint32_t (&a)[N], const uint32_t (&b)[N]) {
#define BW(EXPR)                                                \
  _Pragma("unroll") for (int k = 0; k < N; ++k) a[k] = EXPR; \
  break;
  switch (op) {
    case AGPU_ADD: BW(w_add<T>(a[k], b[k]))
    case AGPU_SUB: BW(w_sub<T>(a[k], b[k]))
    case AGPU_MUL: BW((lanewise<T, OpMul>(a[k], b[k])))
    case AGPU_DIV: BW((lanewise<T, OpDiv>(a[k], b[k])))
    case AGPU_REM: BW((lanewise<T, OpRem>(a[k], b[k])))
    case AGPU_MIN: BW(w_min<T>(a[k], b[k]))
    case AGPU_MAX: BW(w_max<T>(a[k], b[k]))
    case AGPU_AND: BW(a[k] & b[k])
    case AGPU_OR: BW(a[k] | b[k])
    case AGPU_XOR: BW(a[k] ^ b[k])
    case AGPU_POW:
      if constexpr (std::is_same<T, int32_t>::value) {
        BW((lanewise<T, OpPow>(a[k], b[k])))
      }
      break;
    default: break;
  }
#undef BW
}

// predicate bits of NG granules (4 words each): G bits per granule, granule j at bit j*G
template <typename T, class P, int N>
__device__ __forceinline__ uint32_t words_pred(const uint32_t (&a)[N], const uint32_t (&b)[N]) {
  constexpr int L = 4 / sizeof(T);
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    if constexpr (sizeof(T) == 1) m |= (((P::template lanes<T>(a[k], b[k]) & 0x08040201u) * 0x01010101u) >> 24) << (L * k);
    else if constexpr (sizeof(T) == 2) m |= ((((P::template lanes<T>(a[k], b[k]) & 0x00020001u) * 0x00010001u) >> 16) & 3u) << (L * k);
    else m |= (uint32_t)P{}((T)a[k], (T)b[k]) << k;
  }
  return m;
}
template <typename T, int N>
__device__ __forceinline__ uint32_t words_compare(int op, const uint32_t (&a)[N], const uint32_t (&b)[N]) {
  switch (op) {
    case AGPU_GT: return words_pred<T, PGt, N>(a, b);
    case AGPU_GTEQ: return words_pred<T, PGe, N>(a, b);
    case AGPU_LT: return words_pred<T, PLt, N>(a, b);
    case AGPU_LTEQ: return words_pred<T, PLe, N>(a, b);
    default: return words_pred<T, PEq, N>(a, b);
  }
}

// NC = number of operand columns of the chain: the granules of unused column slots would
// otherwise still occupy registers (16 per slot with two granules in flight)
// SH = the chain has a shift step: its u32 counts are G/4 more 16-byte chunks per granule
template <typename T, int NC, bool SH>
struct IntChainOp {
  static constexpr int G = 16 / sizeof(T);
  static constexpr int NCA = NC ? NC : 1;
  static constexpr int CQ = SH ? G / 4 : 1;  // 16-byte chunks of counts per granule
  IntChainProgram p;
  const T* in;
  T* out;  // value chains only
  struct In { uint4 a; uint4 c[NCA]; uint4 cnt[CQ]; };

  static __device__ __forceinline__ uint4 ld16(const T* base, size_t g) {
    return __ldcs(reinterpret_cast<const uint4*>(base) + g);
  }
  __device__ __forceinline__ In load(size_t g) const {
    In r;
    r.a = ld16(in, g);
#pragma unroll
    for (int k = 0; k < NC; ++k)
      if (k < p.n_cols) r.c[k] = ld16((const T*)p.cols[k], g);  // NC may exceed n_cols in the shift variants
    if constexpr (SH) {
#pragma unroll
      for (int q = 0; q < CQ; ++q) r.cnt[q] = __ldcs(reinterpret_cast<const uint4*>(p.counts) + g * CQ + q);
    }
    return r;
  }
  static constexpr bool JOINT = true;  // all granules of a full tile share one pass over the steps

  // acc / rhs: 4 words per granule
  template <int U>
  __device__ __forceinline__ void eval(const In (&inu)[U], uint32_t (&acc)[4 * U], uint32_t (&rhs)[4 * U], int& cmp_op) const {
    uint4 cc[U][NCA];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      acc[4 * j] = inu[j].a.x; acc[4 * j + 1] = inu[j].a.y; acc[4 * j + 2] = inu[j].a.z; acc[4 * j + 3] = inu[j].a.w;
#pragma unroll
      for (int c = 0; c < NC; ++c) cc[j][c] = inu[j].c[c];
    }
    cmp_op = -1;
#pragma unroll 1
    for (int s = 0; s < p.n_steps; ++s) {
      const int kind = p.kind[s];
      if (kind == AGPU_STEP_UNARY) {
        words_unary<T, 4 * U>(p.op[s], acc);
        continue;
      }
      if constexpr (SH) {
        if (kind == AGPU_STEP_SHIFT_COLUMN) {
          constexpr int L = 4 / sizeof(T);
          const bool left = p.op[s] == AGPU_SHL;
#pragma unroll
          for (int j = 0; j < U; ++j) {
            uint32_t c[G];  // the granule's counts, row order
            memcpy(c, inu[j].cnt, sizeof(c));
#pragma unroll
            for (int w = 0; w < 4; ++w)
              acc[4 * j + w] = left ? shift_lanes<T, true>(acc[4 * j + w], c + w * L) : shift_lanes<T, false>(acc[4 * j + w], c + w * L);
          }
          continue;
        }
      }
      if (kind == AGPU_STEP_BINARY_DEVSCALAR || kind == AGPU_STEP_COMPARE_DEVSCALAR) {
        const uint32_t v = w_splat<T>(__ldg((const T*)p.dscalar[s]));
#pragma unroll
        for (int k = 0; k < 4 * U; ++k) rhs[k] = v;
      } else if constexpr (NC > 0) {  // next operand column: slot 0, then the slots move up
#pragma unroll
        for (int j = 0; j < U; ++j) {
          rhs[4 * j] = cc[j][0].x; rhs[4 * j + 1] = cc[j][0].y; rhs[4 * j + 2] = cc[j][0].z; rhs[4 * j + 3] = cc[j][0].w;
#pragma unroll
          for (int c = 0; c + 1 < NC; ++c) cc[j][c] = cc[j][c + 1];
        }
      }
      if (kind == AGPU_STEP_BINARY_COLUMN || kind == AGPU_STEP_BINARY_DEVSCALAR)
        words_binary<T, 4 * U>(p.op[s], acc, rhs);
      else cmp_op = p.op[s];  // compare is the last step (checked on the host)
    }
  }
  template <int U>
  __device__ __forceinline__ void run_joint(size_t g0, const In (&inu)[U]) const {
    uint32_t acc[4 * U], rhs[4 * U];
    int cmp;
    eval<U>(inu, acc, rhs, cmp);
#pragma unroll
    for (int j = 0; j < U; ++j)
      __stcs(reinterpret_cast<uint4*>(out) + g0 + (size_t)j * kBlock,
             make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
  }
  template <int U>
  __device__ __forceinline__ void bits_joint(size_t, const In (&inu)[U], uint32_t (&b)[U]) const {
    uint32_t acc[4 * U], rhs[4 * U];
    int cmp;
    eval<U>(inu, acc, rhs, cmp);
    // words_compare packs L bits per word, 4 words per granule: granule j at bit j*G
    if constexpr (G * U <= 32) {
      const uint32_t m = words_compare<T, 4 * U>(cmp, acc, rhs);
#pragma unroll
      for (int j = 0; j < U; ++j) b[j] = (m >> (G * j)) & (G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u));
    } else {
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const uint32_t a4[4] = {acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]};
        const uint32_t r4[4] = {rhs[4 * j], rhs[4 * j + 1], rhs[4 * j + 2], rhs[4 * j + 3]};
        b[j] = words_compare<T, 4>(cmp, a4, r4);
      }
    }
  }
  __device__ __forceinline__ void eval1(const In& in1, uint32_t (&acc)[4], uint32_t (&rhs)[4], int& cmp_op) const {
    const In one[1] = {in1};
    eval<1>(one, acc, rhs, cmp_op);
  }
  __device__ __forceinline__ In load_row(size_t i) const {  // one row replicated over a granule
    In r;
    const uint32_t v = w_splat<T>(in[i]);
    r.a = make_uint4(v, v, v, v);
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c < p.n_cols) {
        const uint32_t w = w_splat<T>(((const T*)p.cols[c])[i]);
        r.c[c] = make_uint4(w, w, w, w);
      }
    if constexpr (SH) {
      const uint32_t k = p.counts[i];
#pragma unroll
      for (int q = 0; q < CQ; ++q) r.cnt[q] = make_uint4(k, k, k, k);
    }
    return r;
  }
  // ---- value chain: elementwise Op interface
  __device__ __forceinline__ void run(size_t g, const In& in1) const {
    uint32_t acc[4], rhs[4];
    int cmp;
    eval1(in1, acc, rhs, cmp);
    __stcs(reinterpret_cast<uint4*>(out) + g, make_uint4(acc[0], acc[1], acc[2], acc[3]));
  }
  __device__ __forceinline__ void tail(size_t i) const {
    uint32_t acc[4], rhs[4];
    int cmp;
    eval1(load_row(i), acc, rhs, cmp);
    out[i] = (T)acc[0];
  }
  // ---- predicate chain: BitsOp interface
  __device__ __forceinline__ uint32_t bits(size_t, const In& in1) const {
    uint32_t acc[4], rhs[4];
    int cmp;
    eval1(in1, acc, rhs, cmp);
    return words_compare<T, 4>(cmp, acc, rhs);
  }
  __device__ __forceinline__ bool bit_at(size_t i) const {
    uint32_t acc[4], rhs[4];
    int cmp;
    eval1(load_row(i), acc, rhs, cmp);
    return words_compare<T, 4>(cmp, acc, rhs) & 1u;
  }
};

template <typename T, int NC, bool SH>
int run_int_chain_as(agpu_device* dev, const IntChainProgram& p, const void* in, void* out, size_t n, const BmAnd& bm,
                     bool is_pred) {
  using Op = IntChainOp<T, NC, SH>;
  Op op{p, (const T*)in, (T*)out};
  bool al = aligned16(in) && aligned16(out) && (!SH || aligned16(p.counts));
  for (int k = 0; k < p.n_cols; ++k) al = al && aligned16(p.cols[k]);
  // 1-byte rows with a shift carry 64 bytes of counts per granule: one granule per thread
  constexpr int UNROLL = (SH && sizeof(T) == 1) ? 1 : 2;
  if (is_pred) return launch_bits<Op, UNROLL>(dev, op, (uint32_t*)out, n, bm, al);
  return launch_ew<Op, UNROLL>(dev, op, n, bm, al);
}

template <typename T>
int run_int_chain(agpu_device* dev, const IntChainProgram& p, const void* in, void* out, size_t n, const BmAnd& bm,
                  bool is_pred) {
  if (p.counts) {  // shift variants: column-free or generic, to bound the number of kernels
    if (p.n_cols == 0) return run_int_chain_as<T, 0, true>(dev, p, in, out, n, bm, is_pred);
    return run_int_chain_as<T, kMaxCols - 1, true>(dev, p, in, out, n, bm, is_pred);
  }
  switch (p.n_cols) {
    case 0: return run_int_chain_as<T, 0, false>(dev, p, in, out, n, bm, is_pred);
    case 1: return run_int_chain_as<T, 1, false>(dev, p, in, out, n, bm, is_pred);
    case 2: return run_int_chain_as<T, 2, false>(dev, p, in, out, n, bm, is_pred);
    default: return run_int_chain_as<T, 3, false>(dev, p, in, out, n, bm, is_pred);
  }
}

}  // namespace

extern "C" int agpu_fused_chain_int(agpu_device* dev, int dtype, const void* in, const uint32_t* vin,
                                    const agpu_chain_step* steps, int n_steps, void* out, size_t n, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (!steps || n_steps < 1 || n_steps > AGPU_CHAIN_MAX_STEPS) return AGPU_EINVAL;
  if (n && (!in || !out)) return AGPU_EINVAL;
  const bool is_i32 = dtype == AGPU_I32 || dtype == AGPU_DATE32;
  IntChainProgram p{};
  p.n_steps = n_steps;
  const uint32_t* vals[4] = {vin, nullptr, nullptr, nullptr};
  const uint32_t* col_validity[kMaxCols] = {nullptr, nullptr, nullptr};
  const uint32_t* counts_validity = nullptr;
  bool is_pred = false;
  for (int s = 0; s < n_steps; ++s) {
    const agpu_chain_step& st = steps[s];
    p.kind[s] = st.kind;
    p.op[s] = st.op;
    p.dscalar[s] = nullptr;
    switch (st.kind) {
      case AGPU_STEP_UNARY:
        if (st.op != AGPU_NOT && !(st.op == AGPU_ABS && is_i32)) return AGPU_EUNSUPPORTED;
        break;
      case AGPU_STEP_BINARY_COLUMN:
      case AGPU_STEP_BINARY_DEVSCALAR:
        if (st.op < AGPU_ADD || st.op > AGPU_POW || (st.op == AGPU_POW && !is_i32)) return AGPU_EUNSUPPORTED;
        break;
      case AGPU_STEP_COMPARE_COLUMN:
      case AGPU_STEP_COMPARE_DEVSCALAR:
        if (st.op < AGPU_GT || st.op > AGPU_EQ) return AGPU_EUNSUPPORTED;
        if (s != n_steps - 1) return AGPU_EINVAL;  // a predicate ends the chain
        is_pred = true;
        break;
      case AGPU_STEP_SHIFT_COLUMN:
        if (st.op != AGPU_SHL && st.op != AGPU_SHR) return AGPU_EUNSUPPORTED;
        if (p.counts || dtype == AGPU_DATE32) return AGPU_EUNSUPPORTED;  // one shift step per chain
        break;
      case AGPU_STEP_BINARY_SCALAR:
      case AGPU_STEP_COMPARE_SCALAR:
        return AGPU_EUNSUPPORTED;  // the float immediate cannot hold every i32/u32: use a device scalar
      default: return AGPU_EINVAL;
    }
    if (!st.operand && st.kind != AGPU_STEP_UNARY) return AGPU_EINVAL;
    if (st.kind == AGPU_STEP_BINARY_DEVSCALAR || st.kind == AGPU_STEP_COMPARE_DEVSCALAR) p.dscalar[s] = st.operand;
    if (st.kind == AGPU_STEP_BINARY_COLUMN || st.kind == AGPU_STEP_COMPARE_COLUMN) {
      if (p.n_cols + (p.counts ? 1 : 0) == kMaxCols) return AGPU_EUNSUPPORTED;
      p.cols[p.n_cols] = st.operand;
      col_validity[p.n_cols] = st.validity;
      ++p.n_cols;
    }
    if (st.kind == AGPU_STEP_SHIFT_COLUMN) {
      if (p.n_cols == kMaxCols) return AGPU_EUNSUPPORTED;  // columns + counts share the three bitmap slots
      p.counts = (const uint32_t*)st.operand;
      counts_validity = st.validity;
    }
  }
  for (int k = 0; k < p.n_cols; ++k) vals[1 + k] = col_validity[k];
  if (p.counts) vals[1 + p.n_cols] = counts_validity;
  if (vout && !vals[0] && !vals[1] && !vals[2] && !vals[3]) return AGPU_EINVAL;
  const BmAnd bm = make_bm(vals[0], vals[1], vals[2], vals[3], vout);
  switch (dtype) {
    case AGPU_I32: case AGPU_DATE32: return run_int_chain<int32_t>(dev, p, in, out, n, bm, is_pred);
    case AGPU_U32: return run_int_chain<uint32_t>(dev, p, in, out, n, bm, is_pred);
    case AGPU_I16: return run_int_chain<int16_t>(dev, p, in, out, n, bm, is_pred);
    case AGPU_U16: return run_int_chain<uint16_t>(dev, p, in, out, n, bm, is_pred);
    case AGPU_I8: return run_int_chain<int8_t>(dev, p, in, out, n, bm, is_pred);
    case AGPU_U8: return run_int_chain<uint8_t>(dev, p, in, out, n, bm, is_pred);
    default: return AGPU_EUNSUPPORTED;
  }
}
