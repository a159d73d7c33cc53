// chain_int.cu — agpu_fused_chain_int: the chain interpreter of chain.cu on integer columns.
//
// Every step applies the same functor as the stand-alone integer kernel of that op (ops.cuh:
// two's-complement wrap in the column's own width, the WGSL divide/remainder-by-zero rules,
// signedness of min/max/compare), so a fused chain is bit-identical to the ops run one by one.
// Rows move as 16-byte granules of the column type (16 x i8, 8 x i16, 4 x i32); operand columns
// stay packed in their load registers until the step that consumes them (steps consume the
// columns in order, so the "next column" is always slot 0 and the slots rotate).
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
};

template <typename T, int N>
__device__ __forceinline__ void int_unary(int op, T (&a)[N]) {
  if (op == AGPU_NOT) {
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] = OpNot<T>{}(a[k]);
  } else if constexpr (std::is_same<T, int32_t>::value) {  // AGPU_ABS (int32 only, checked on the host)
#pragma unroll
    for (int k = 0; k < N; ++k) a[k] = OpAbs<T>{}(a[k]);
  }
}

template <typename T, int N>
__device__ __forceinline__ void int_binary(int op, T (&a)[N], const T (&b)[N]) {
#define BN(F)                                                                  \
  _Pragma("unroll") for (int k = 0; k < N; ++k) a[k] = F<T>{}(a[k], b[k]); \
  break;
  switch (op) {
    case AGPU_ADD: BN(OpAdd)
    case AGPU_SUB: BN(OpSub)
    case AGPU_MUL: BN(OpMul)
    case AGPU_DIV: BN(OpDiv)
    case AGPU_REM: BN(OpRem)
    case AGPU_MIN: BN(OpMin)
    case AGPU_MAX: BN(OpMax)
    case AGPU_AND: BN(OpAnd)
    case AGPU_OR: BN(OpOr)
    case AGPU_XOR: BN(OpXor)
    case AGPU_POW:
      if constexpr (std::is_same<T, int32_t>::value) {
        BN(OpPow)
      }
      break;
    default: break;
  }
#undef BN
}

template <typename T, int N>
__device__ __forceinline__ uint32_t int_compare(int op, const T (&a)[N], const T (&b)[N]) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    bool p;
    switch (op) {
      case AGPU_GT: p = a[k] > b[k]; break;
      case AGPU_GTEQ: p = a[k] >= b[k]; break;
      case AGPU_LT: p = a[k] < b[k]; break;
      case AGPU_LTEQ: p = a[k] <= b[k]; break;
      default: p = a[k] == b[k]; break;
    }
    m |= (uint32_t)p << k;
  }
  return m;
}

template <typename T>
struct IntChainOp {
  static constexpr int G = 16 / sizeof(T);
  IntChainProgram p;
  const T* in;
  T* out;  // value chains only
  struct In { Vec<T, G> a; Vec<T, G> c[kMaxCols]; };

  __device__ __forceinline__ In load(size_t g) const {
    In r;
    r.a = ld_vec<T, G>(in, g);
#pragma unroll
    for (int k = 0; k < kMaxCols; ++k)
      if (k < p.n_cols) r.c[k] = ld_vec<T, G>((const T*)p.cols[k], g);
    return r;
  }
  // 32-bit rows: two granules (8 accumulators) share one pass over the steps; narrower rows
  // already have 8 or 16 accumulators per granule
  static constexpr bool JOINT = sizeof(T) == 4;

  template <int U>
  __device__ __forceinline__ void eval(const In (&inu)[U], T (&acc)[G * U], T (&rhs)[G * U], int& cmp_op) const {
    Vec<T, G> cc[U][kMaxCols];
#pragma unroll
    for (int j = 0; j < U; ++j) {
#pragma unroll
      for (int k = 0; k < G; ++k) acc[j * G + k] = inu[j].a.e[k];
#pragma unroll
      for (int c = 0; c < kMaxCols; ++c) cc[j][c] = inu[j].c[c];
    }
    cmp_op = -1;
#pragma unroll 1
    for (int s = 0; s < p.n_steps; ++s) {
      const int kind = p.kind[s];
      if (kind == AGPU_STEP_UNARY) {
        int_unary<T, G * U>(p.op[s], acc);
        continue;
      }
      if (kind == AGPU_STEP_BINARY_DEVSCALAR || kind == AGPU_STEP_COMPARE_DEVSCALAR) {
        const T v = __ldg((const T*)p.dscalar[s]);
#pragma unroll
        for (int k = 0; k < G * U; ++k) rhs[k] = v;
      } else {  // next operand column: slot 0, then the slots move up
#pragma unroll
        for (int j = 0; j < U; ++j) {
#pragma unroll
          for (int k = 0; k < G; ++k) rhs[j * G + k] = cc[j][0].e[k];
          cc[j][0] = cc[j][1];
          cc[j][1] = cc[j][2];
        }
      }
      if (kind == AGPU_STEP_BINARY_COLUMN || kind == AGPU_STEP_BINARY_DEVSCALAR)
        int_binary<T, G * U>(p.op[s], acc, rhs);
      else cmp_op = p.op[s];  // compare is the last step (checked on the host)
    }
  }
  template <int U>
  __device__ __forceinline__ void run_joint(size_t g0, const In (&inu)[U]) const {
    T acc[G * U], rhs[G * U];
    int cmp;
    eval<U>(inu, acc, rhs, cmp);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      Vec<T, G> o;
#pragma unroll
      for (int k = 0; k < G; ++k) o.e[k] = acc[j * G + k];
      st_vec<T, G>(out, g0 + (size_t)j * kBlock, o);
    }
  }
  template <int U>
  __device__ __forceinline__ void bits_joint(size_t, const In (&inu)[U], uint32_t (&b)[U]) const {
    T acc[G * U], rhs[G * U];
    int cmp;
    eval<U>(inu, acc, rhs, cmp);
    const uint32_t m = int_compare<T, G * U>(cmp, acc, rhs);
#pragma unroll
    for (int j = 0; j < U; ++j) b[j] = (m >> (G * j)) & ((1u << G) - 1u);
  }
  __device__ __forceinline__ void eval1(const In& in1, T (&acc)[G], T (&rhs)[G], int& cmp_op) const {
    const In one[1] = {in1};
    eval<1>(one, acc, rhs, cmp_op);
  }
  __device__ __forceinline__ In load_row(size_t i) const {  // one row replicated over a granule
    In r;
    const T v = in[i];
#pragma unroll
    for (int k = 0; k < G; ++k) r.a.e[k] = v;
#pragma unroll
    for (int c = 0; c < kMaxCols; ++c)
      if (c < p.n_cols) {
        const T w = ((const T*)p.cols[c])[i];
#pragma unroll
        for (int k = 0; k < G; ++k) r.c[c].e[k] = w;
      }
    return r;
  }
  // ---- value chain: elementwise Op interface
  __device__ __forceinline__ void run(size_t g, const In& in1) const {
    T acc[G], rhs[G];
    int cmp;
    eval1(in1, acc, rhs, cmp);
    Vec<T, G> o;
#pragma unroll
    for (int k = 0; k < G; ++k) o.e[k] = acc[k];
    st_vec<T, G>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const {
    T acc[G], rhs[G];
    int cmp;
    eval1(load_row(i), acc, rhs, cmp);
    out[i] = acc[0];
  }
  // ---- predicate chain: BitsOp interface
  __device__ __forceinline__ uint32_t bits(size_t, const In& in1) const {
    T acc[G], rhs[G];
    int cmp;
    eval1(in1, acc, rhs, cmp);
    return int_compare<T, G>(cmp, acc, rhs);
  }
  __device__ __forceinline__ bool bit_at(size_t i) const {
    T acc[G], rhs[G];
    int cmp;
    eval1(load_row(i), acc, rhs, cmp);
    return int_compare<T, G>(cmp, acc, rhs) & 1u;
  }
};

template <typename T>
int run_int_chain(agpu_device* dev, const IntChainProgram& p, const void* in, void* out, size_t n, const BmAnd& bm,
                  bool is_pred) {
  IntChainOp<T> op{p, (const T*)in, (T*)out};
  bool al = aligned16(in) && aligned16(out);
  for (int k = 0; k < p.n_cols; ++k) al = al && aligned16(p.cols[k]);
  if (is_pred) return launch_bits<IntChainOp<T>, 2>(dev, op, (uint32_t*)out, n, bm, al);
  return launch_ew<IntChainOp<T>, 2>(dev, op, n, bm, al);
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
      case AGPU_STEP_BINARY_SCALAR:
      case AGPU_STEP_COMPARE_SCALAR:
        return AGPU_EUNSUPPORTED;  // the float immediate cannot hold every i32/u32: use a device scalar
      default: return AGPU_EINVAL;
    }
    if (!st.operand && st.kind != AGPU_STEP_UNARY) return AGPU_EINVAL;
    if (st.kind == AGPU_STEP_BINARY_DEVSCALAR || st.kind == AGPU_STEP_COMPARE_DEVSCALAR) p.dscalar[s] = st.operand;
    if (st.kind == AGPU_STEP_BINARY_COLUMN || st.kind == AGPU_STEP_COMPARE_COLUMN) {
      if (p.n_cols == kMaxCols) return AGPU_EUNSUPPORTED;
      p.cols[p.n_cols] = st.operand;
      vals[1 + p.n_cols] = st.validity;
      ++p.n_cols;
    }
  }
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
