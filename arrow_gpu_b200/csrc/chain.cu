// chain.cu — agpu_fused_chain: a linear chain of f32 ops evaluated in registers, one pass.
//
// The reference records `mul_op -> add_op -> gt_op ...` as separate dispatches, each a full HBM
// round trip plus a fresh buffer.  Here the chain is data: up to 8 steps with
// runtime opcodes.  The accumulator and the operand chunks are statically named registers, the
// opcode switch is evaluated once per step per 4-row granule — warp-uniform, so no divergence —
// and every step calls the same
// device functions as the stand-alone kernels (ops.cuh), which makes the fused result
// bit-identical to the unfused chain.
#include "bits.cuh"
#include "elementwise.cuh"
#include "ops.cuh"

namespace {

constexpr int kMaxCols = 3;

struct ChainProgram {
  int n_steps;
  int n_cols;
  int kind[AGPU_CHAIN_MAX_STEPS];
  int op[AGPU_CHAIN_MAX_STEPS];
  int col[AGPU_CHAIN_MAX_STEPS];  // operand column slot for *_COLUMN steps
  float scalar[AGPU_CHAIN_MAX_STEPS];
  const float* dscalar[AGPU_CHAIN_MAX_STEPS];  // one-element device arrays for *_DEVSCALAR steps
  const float* cols[kMaxCols];
};

// HEAVY = the chain contains a transcendental / pow step.  The arithmetic-only interpreter
// (neg abs sqrt, + - * / % min max, compares) is a separate, much smaller kernel: the register
// allocation of the full one is set by powf / the sinf slow path whatever the chain executes.
template <int N, bool HEAVY>
__device__ __forceinline__ void apply_unary(int op, float (&a)[N]) {
#define U4(F)                          \
  _Pragma("unroll") for (int k = 0; k < N; ++k) a[k] = F<float>{}(a[k]); \
  break;
  if constexpr (!HEAVY) {
    switch (op) {
      case AGPU_NEG: U4(OpNeg)
      case AGPU_ABS: U4(OpAbs)
      default: U4(FSqrt)
    }
    return;
  }
  switch (op) {
    case AGPU_NEG: U4(OpNeg)
    case AGPU_ABS: U4(OpAbs)
    case AGPU_SQRT: U4(FSqrt)
    case AGPU_CBRT: U4(FCbrt)
    case AGPU_EXP: U4(FExp)
    case AGPU_EXP2: U4(FExp2)
    case AGPU_LOG: U4(FLog)
    case AGPU_LOG2: U4(FLog2)
    case AGPU_SIN: U4(FSin)
    case AGPU_COS: U4(FCos)
    case AGPU_ACOS: U4(FAcos)
    case AGPU_SINH: U4(FSinh)
    default: break;
  }
#undef U4
}

template <int N, bool HEAVY>
__device__ __forceinline__ void apply_binary(int op, float (&a)[N], const float (&b)[N]) {
#define B4(F)                          \
  _Pragma("unroll") for (int k = 0; k < N; ++k) a[k] = F<float>{}(a[k], b[k]); \
  break;
  if constexpr (!HEAVY) {
    switch (op) {
      case AGPU_ADD: B4(OpAdd)
      case AGPU_SUB: B4(OpSub)
      case AGPU_MUL: B4(OpMul)
      case AGPU_DIV: B4(OpDiv)
      case AGPU_REM: B4(OpRem)
      case AGPU_MIN: B4(OpMin)
      default: B4(OpMax)
    }
    return;
  }
  switch (op) {
    case AGPU_ADD: B4(OpAdd)
    case AGPU_SUB: B4(OpSub)
    case AGPU_MUL: B4(OpMul)
    case AGPU_DIV: B4(OpDiv)
    case AGPU_REM: B4(OpRem)
    case AGPU_MIN: B4(OpMin)
    case AGPU_MAX: B4(OpMax)
    case AGPU_POW: B4(OpPow)
    default: break;
  }
#undef B4
}

template <int N>
__device__ __forceinline__ uint32_t apply_compare(int op, const float (&a)[N], const float (&b)[N]) {
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

// NC = operand-column slots held in registers (8 registers per slot with two granules in
// flight); the arithmetic-only kernels are specialised on the exact column count
template <typename TI, int NC, bool HEAVY>
struct ChainOp {
  static constexpr int G = 4;
  static constexpr int NCA = NC ? NC : 1;
  ChainProgram p;
  const TI* in;
  float* out;  // value chains only
  struct In { Vec<TI, 4> a; Vec<float, 4> c[NCA]; };

  __device__ __forceinline__ In load(size_t g) const {
    In r;
    r.a = ld_vec<TI, 4>(in, g);
#pragma unroll
    for (int k = 0; k < NC; ++k)
      if (k < p.n_cols) r.c[k] = ld_vec<float, 4>(p.cols[k], g);
    return r;
  }
  static constexpr bool JOINT = true;  // full tiles: all UNROLL granules go through the step loop together

  // runs every step but a trailing compare on U granules at once (4*U accumulators): the opcode
  // dispatch of a step is paid once per 4*U rows
  template <int U>
  __device__ __forceinline__ void eval(const In (&in4)[U], float (&acc)[4 * U], float (&rhs)[4 * U], int& cmp_op) const {
#pragma unroll
    for (int j = 0; j < U; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[j * 4 + k] = (float)in4[j].a.e[k];
    cmp_op = -1;
    // runtime loop over the steps: the opcodes live in the kernel's constant bank, the
    // accumulator / operand chunks are statically named registers, so nothing is indexed
    // dynamically and the body (with every transcendental inlined) exists once in the code
#pragma unroll 1
    for (int s = 0; s < p.n_steps; ++s) {
      const int kind = p.kind[s];
      if (kind == AGPU_STEP_UNARY) {
        apply_unary<4 * U, HEAVY>(p.op[s], acc);
      } else {
        if (kind == AGPU_STEP_BINARY_SCALAR || kind == AGPU_STEP_COMPARE_SCALAR) {
#pragma unroll
          for (int k = 0; k < 4 * U; ++k) rhs[k] = p.scalar[s];
        } else if (kind == AGPU_STEP_BINARY_DEVSCALAR || kind == AGPU_STEP_COMPARE_DEVSCALAR) {
          const float v = __ldg(p.dscalar[s]);
#pragma unroll
          for (int k = 0; k < 4 * U; ++k) rhs[k] = v;
        } else if constexpr (NC > 0) {
          const int c = p.col[s];  // warp-uniform select between the statically named column chunks
#pragma unroll
          for (int j = 0; j < U; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float v = in4[j].c[0].e[k];
              if constexpr (NC > 1) v = c == 1 ? in4[j].c[1].e[k] : v;
              if constexpr (NC > 2) v = c == 2 ? in4[j].c[2].e[k] : v;
              rhs[j * 4 + k] = v;
            }
        }
        if (kind == AGPU_STEP_BINARY_COLUMN || kind == AGPU_STEP_BINARY_SCALAR || kind == AGPU_STEP_BINARY_DEVSCALAR)
          apply_binary<4 * U, HEAVY>(p.op[s], acc, rhs);
        else cmp_op = p.op[s];  // compare is the last step (checked on the host)
      }
    }
  }
  template <int U>
  __device__ __forceinline__ void run_joint(size_t g0, const In (&in4)[U]) const {
    float acc[4 * U], rhs[4 * U];
    int cmp;
    eval<U>(in4, acc, rhs, cmp);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      Vec<float, 4> o;
#pragma unroll
      for (int k = 0; k < 4; ++k) o.e[k] = acc[j * 4 + k];
      st_vec<float, 4>(out, g0 + (size_t)j * kBlock, o);
    }
  }
  template <int U>
  __device__ __forceinline__ void bits_joint(size_t, const In (&in4)[U], uint32_t (&b)[U]) const {
    float acc[4 * U], rhs[4 * U];
    int cmp;
    eval<U>(in4, acc, rhs, cmp);
    const uint32_t m = apply_compare<4 * U>(cmp, acc, rhs);
#pragma unroll
    for (int j = 0; j < U; ++j) b[j] = (m >> (4 * j)) & 0xFu;
  }
  __device__ __forceinline__ void eval1(const In& in1, float (&acc)[4], float (&rhs)[4], int& cmp_op) const {
    const In one[1] = {in1};
    eval<1>(one, acc, rhs, cmp_op);
  }
  __device__ __forceinline__ In load_row(size_t i) const {  // one row replicated into lane 0 of a chunk
    In r;
    r.a.e[0] = in[i];
#pragma unroll
    for (int k = 1; k < 4; ++k) r.a.e[k] = r.a.e[0];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c < p.n_cols) {
        const float v = p.cols[c][i];
#pragma unroll
        for (int k = 0; k < 4; ++k) r.c[c].e[k] = v;
      }
    return r;
  }
  // ---- value chain: elementwise Op interface
  __device__ __forceinline__ void run(size_t g, const In& in4) const {
    float acc[4], rhs[4];
    int cmp;
    eval1(in4, acc, rhs, cmp);
    Vec<float, 4> o;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.e[k] = acc[k];
    st_vec<float, 4>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const {
    float acc[4], rhs[4];
    int cmp;
    eval1(load_row(i), acc, rhs, cmp);
    out[i] = acc[0];
  }
  // ---- predicate chain: BitsOp interface
  __device__ __forceinline__ uint32_t bits(size_t, const In& in4) const {
    float acc[4], rhs[4];
    int cmp;
    eval1(in4, acc, rhs, cmp);
    return apply_compare<4>(cmp, acc, rhs);
  }
  __device__ __forceinline__ bool bit_at(size_t i) const {
    float acc[4], rhs[4];
    int cmp;
    eval1(load_row(i), acc, rhs, cmp);
    return apply_compare<4>(cmp, acc, rhs) & 1u;
  }
};

template <typename TI, int NC, bool HEAVY>
int run_chain_as(agpu_device* dev, const ChainProgram& p, const void* in, void* out, size_t n, const BmAnd& bm, bool is_pred) {
  using Op = ChainOp<TI, NC, HEAVY>;
  Op op{p, (const TI*)in, (float*)out};
  bool al = aligned16(in) && aligned16(out);
  for (int k = 0; k < p.n_cols; ++k) al = al && aligned16(p.cols[k]);
  // two granules per thread evaluated jointly; four (126 registers in the full interpreter)
  // measured 30 % slower
  if (is_pred) return launch_bits<Op, 2>(dev, op, (uint32_t*)out, n, bm, al);
  return launch_ew<Op, 2>(dev, op, n, bm, al);
}

template <typename TI>
int run_chain(agpu_device* dev, const ChainProgram& p, const void* in, void* out, size_t n, const BmAnd& bm, bool is_pred,
              bool heavy) {
  if (!heavy) {
    if constexpr (std::is_same<TI, float>::value) {
      switch (p.n_cols) {
        case 0: return run_chain_as<TI, 0, false>(dev, p, in, out, n, bm, is_pred);
        case 1: return run_chain_as<TI, 1, false>(dev, p, in, out, n, bm, is_pred);
        case 2: return run_chain_as<TI, 2, false>(dev, p, in, out, n, bm, is_pred);
        default: return run_chain_as<TI, 3, false>(dev, p, in, out, n, bm, is_pred);
      }
    } else {  // fused int -> f32 cast + arithmetic (e.g. u8 * scale + offset): two variants
      if (p.n_cols == 0) return run_chain_as<TI, 0, false>(dev, p, in, out, n, bm, is_pred);
      return run_chain_as<TI, kMaxCols, false>(dev, p, in, out, n, bm, is_pred);
    }
  }
  return run_chain_as<TI, kMaxCols, true>(dev, p, in, out, n, bm, is_pred);
}

}  // namespace

extern "C" int agpu_fused_chain(agpu_device* dev, int in_dtype, const void* in, const uint32_t* vin,
                                const agpu_chain_step* steps, int n_steps, void* out, size_t n, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (!steps || n_steps < 1 || n_steps > AGPU_CHAIN_MAX_STEPS) return AGPU_EINVAL;
  if (n && (!in || !out)) return AGPU_EINVAL;
  ChainProgram p{};
  p.n_steps = n_steps;
  const uint32_t* vals[4] = {vin, nullptr, nullptr, nullptr};
  bool is_pred = false, heavy = false;
  for (int s = 0; s < n_steps; ++s) {
    const agpu_chain_step& st = steps[s];
    if (st.kind == AGPU_STEP_UNARY) heavy = heavy || (st.op != AGPU_NEG && st.op != AGPU_ABS && st.op != AGPU_SQRT);
    else if (st.kind == AGPU_STEP_BINARY_COLUMN || st.kind == AGPU_STEP_BINARY_SCALAR || st.kind == AGPU_STEP_BINARY_DEVSCALAR)
      heavy = heavy || st.op == AGPU_POW;
    p.kind[s] = st.kind;
    p.op[s] = st.op;
    p.scalar[s] = st.scalar;
    p.dscalar[s] = nullptr;
    p.col[s] = 0;
    switch (st.kind) {
      case AGPU_STEP_UNARY:
        if (st.op != AGPU_NEG && st.op != AGPU_ABS && (st.op < AGPU_SQRT || st.op > AGPU_SINH)) return AGPU_EUNSUPPORTED;
        break;
      case AGPU_STEP_BINARY_COLUMN:
      case AGPU_STEP_BINARY_SCALAR:
      case AGPU_STEP_BINARY_DEVSCALAR:
        if (st.op < AGPU_ADD || st.op > AGPU_POW || (st.op >= AGPU_AND && st.op <= AGPU_XOR)) return AGPU_EUNSUPPORTED;
        break;
      case AGPU_STEP_COMPARE_COLUMN:
      case AGPU_STEP_COMPARE_SCALAR:
      case AGPU_STEP_COMPARE_DEVSCALAR:
        if (st.op < AGPU_GT || st.op > AGPU_EQ) return AGPU_EUNSUPPORTED;
        if (s != n_steps - 1) return AGPU_EINVAL;  // a predicate ends the chain
        is_pred = true;
        break;
      default: return AGPU_EINVAL;
    }
    if (st.kind == AGPU_STEP_BINARY_DEVSCALAR || st.kind == AGPU_STEP_COMPARE_DEVSCALAR) {
      if (!st.operand) return AGPU_EINVAL;
      p.dscalar[s] = (const float*)st.operand;
    }
    if (st.kind == AGPU_STEP_BINARY_COLUMN || st.kind == AGPU_STEP_COMPARE_COLUMN) {
      if (!st.operand) return AGPU_EINVAL;
      if (p.n_cols == kMaxCols) return AGPU_EUNSUPPORTED;
      p.col[s] = p.n_cols;
      p.cols[p.n_cols] = (const float*)st.operand;
      vals[1 + p.n_cols] = st.validity;
      ++p.n_cols;
    }
  }
  if (vout && !vals[0] && !vals[1] && !vals[2] && !vals[3]) return AGPU_EINVAL;
  // the config-3 expression ((a*b)+c) > d has a dedicated kernel with the same roundings
  // (compare.cu): a recorded mul -> add -> gt chain over three columns is routed to it
  if (in_dtype == AGPU_F32 && n_steps == 3 && steps[0].kind == AGPU_STEP_BINARY_COLUMN && steps[0].op == AGPU_MUL &&
      steps[1].kind == AGPU_STEP_BINARY_COLUMN && steps[1].op == AGPU_ADD && steps[2].kind == AGPU_STEP_COMPARE_COLUMN &&
      steps[2].op == AGPU_GT)
    return agpu_fused_mul_add_gt(dev, (const float*)in, p.cols[0], p.cols[1], p.cols[2], (uint32_t*)out, n, vals[0], vals[1],
                                 vals[2], vals[3], vout);
  const BmAnd bm = make_bm(vals[0], vals[1], vals[2], vals[3], vout);
  switch (in_dtype) {
    case AGPU_F32: return run_chain<float>(dev, p, in, out, n, bm, is_pred, heavy);
    case AGPU_I8: return run_chain<int8_t>(dev, p, in, out, n, bm, is_pred, heavy);
    case AGPU_U8: return run_chain<uint8_t>(dev, p, in, out, n, bm, is_pred, heavy);
    case AGPU_I16: return run_chain<int16_t>(dev, p, in, out, n, bm, is_pred, heavy);
    case AGPU_U16: return run_chain<uint16_t>(dev, p, in, out, n, bm, is_pred, heavy);
    default: return AGPU_EUNSUPPORTED;
  }
}
