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
struct NoStore {
  template <int N> __device__ __forceinline__ void operator()(const float (&)[N]) const {}
};

// DUAL = agpu_fused_chain_pair: a predicate kernel whose program also stores the running value to
// `out2` at its AGPU_STEP_STORE step and restarts from the source at AGPU_STEP_RESET
template <typename TI, int NC, bool HEAVY, bool DUAL = false>
struct ChainOp {
  static constexpr int G = 4;
  static constexpr int NCA = NC ? NC : 1;
  ChainProgram p;
  const TI* in;
  float* out;   // value chains only
  float* out2;  // DUAL only: the value result
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
  template <int U, class Store>
  __device__ __forceinline__ void eval(const In (&in4)[U], float (&acc)[4 * U], float (&rhs)[4 * U], int& cmp_op,
                                       const Store& store) const {
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
      if constexpr (DUAL) {
        if (kind == AGPU_STEP_STORE) {
          store(acc);
          continue;
        }
        if (kind == AGPU_STEP_RESET) {
#pragma unroll
          for (int j = 0; j < U; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[j * 4 + k] = (float)in4[j].a.e[k];
          continue;
        }
      }
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
    eval<U>(in4, acc, rhs, cmp, NoStore{});
#pragma unroll
    for (int j = 0; j < U; ++j) {
      Vec<float, 4> o;
#pragma unroll
      for (int k = 0; k < 4; ++k) o.e[k] = acc[j * 4 + k];
      st_vec<float, 4>(out, g0 + (size_t)j * kBlock, o);
    }
  }
  template <int U>
  __device__ __forceinline__ void bits_joint(size_t g0, const In (&in4)[U], uint32_t (&b)[U]) const {
    float acc[4 * U], rhs[4 * U];
    int cmp;
    float* const value_out = out2;
    eval<U>(in4, acc, rhs, cmp, [value_out, g0](const float (&v)[4 * U]) {
#pragma unroll
      for (int j = 0; j < U; ++j) {
        Vec<float, 4> o;
#pragma unroll
        for (int k = 0; k < 4; ++k) o.e[k] = v[j * 4 + k];
        st_vec<float, 4>(value_out, g0 + (size_t)j * kBlock, o);
      }
    });
    const uint32_t m = apply_compare<4 * U>(cmp, acc, rhs);
#pragma unroll
    for (int j = 0; j < U; ++j) b[j] = (m >> (4 * j)) & 0xFu;
  }
  template <class Store>
  __device__ __forceinline__ void eval1(const In& in1, float (&acc)[4], float (&rhs)[4], int& cmp_op, const Store& store) const {
    const In one[1] = {in1};
    eval<1>(one, acc, rhs, cmp_op, store);
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
    eval1(in4, acc, rhs, cmp, NoStore{});
    Vec<float, 4> o;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.e[k] = acc[k];
    st_vec<float, 4>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const {
    float acc[4], rhs[4];
    int cmp;
    eval1(load_row(i), acc, rhs, cmp, NoStore{});
    out[i] = acc[0];
  }
  // ---- predicate chain: BitsOp interface
  __device__ __forceinline__ uint32_t bits(size_t g, const In& in4) const {
    float acc[4], rhs[4];
    int cmp;
    float* const value_out = out2;
    eval1(in4, acc, rhs, cmp, [value_out, g](const float (&v)[4]) {
      Vec<float, 4> o;
#pragma unroll
      for (int k = 0; k < 4; ++k) o.e[k] = v[k];
      st_vec<float, 4>(value_out, g, o);
    });
    return apply_compare<4>(cmp, acc, rhs);
  }
  __device__ __forceinline__ bool bit_at(size_t i) const {
    float acc[4], rhs[4];
    int cmp;
    float* const value_out = out2;
    eval1(load_row(i), acc, rhs, cmp, [value_out, i](const float (&v)[4]) { value_out[i] = v[0]; });
    return apply_compare<4>(cmp, acc, rhs) & 1u;
  }
};

template <typename TI, int NC, bool HEAVY>
int run_chain_as(agpu_device* dev, const ChainProgram& p, const void* in, void* out, size_t n, const BmAnd& bm, bool is_pred) {
  using Op = ChainOp<TI, NC, HEAVY>;
  Op op{p, (const TI*)in, (float*)out, nullptr};
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

// ---- value chain + predicate chain of one source column in one pass (agpu.h) ----
namespace {

// The shape the pair exists for — value = a binop b, predicate = a cmp c (c is usually b: the
// reference's first benchmark program s = a + b; g = a > b) — as a dedicated streaming kernel.
// The interpreter below spends ~45 instructions per row on a four-step pair program and is then
// issue-bound (0.73 of the copy peak at 256 Mi rows, ncu: 47 % issue slots at 39 % occupancy);
// this one is a plain load-compute-store loop like compare.cu's ((a*b)+c) > d.
template <class F, bool SAME>
struct PairBinCmp {
  static constexpr int G = 4;
  const float *a, *b, *c;
  float* out_value;
  int cmp;
  struct In { Vec<float, 4> a, b, c; };
  __device__ __forceinline__ In load(size_t g) const {
    In r;
    r.a = ld_vec<float, 4>(a, g);
    r.b = ld_vec<float, 4>(b, g);
    if constexpr (!SAME) r.c = ld_vec<float, 4>(c, g);
    return r;
  }
  __device__ __forceinline__ uint32_t test(const float (&x)[4], const float (&y)[4]) const {
    uint32_t m = 0;
    switch (cmp) {  // warp-uniform, once per granule
#define P4(EXPR) _Pragma("unroll") for (int k = 0; k < 4; ++k) m |= (uint32_t)(EXPR) << k; break;
      case AGPU_GT: P4(x[k] > y[k])
      case AGPU_GTEQ: P4(x[k] >= y[k])
      case AGPU_LT: P4(x[k] < y[k])
      case AGPU_LTEQ: P4(x[k] <= y[k])
      default: P4(x[k] == y[k])
#undef P4
    }
    return m;
  }
  __device__ __forceinline__ uint32_t bits(size_t g, const In& in) const {
    Vec<float, 4> o;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.e[k] = F{}(in.a.e[k], in.b.e[k]);
    st_vec<float, 4>(out_value, g, o);
    return test(in.a.e, SAME ? in.b.e : in.c.e);
  }
  __device__ __forceinline__ bool bit_at(size_t i) const {
    const float x = a[i], y = b[i];
    out_value[i] = F{}(x, y);
    const float xs[4] = {x, x, x, x};
    const float z = SAME ? y : c[i];
    const float zs[4] = {z, z, z, z};
    return test(xs, zs) & 1u;
  }
};

template <class F>
int run_bin_cmp(agpu_device* dev, const float* a, const float* b, const float* c, int cmp, float* out_value,
                uint32_t* out_bits, size_t n, const BmAnd& bm) {
  const bool al = aligned16(a) && aligned16(b) && aligned16(c) && aligned16(out_value) && aligned16(out_bits);
  if (b == c) return launch_bits(dev, PairBinCmp<F, true>{a, b, c, out_value, cmp}, out_bits, n, bm, al);
  return launch_bits(dev, PairBinCmp<F, false>{a, b, c, out_value, cmp}, out_bits, n, bm, al);
}

template <int NC>
int run_pair_as(agpu_device* dev, const ChainProgram& p, const float* in, float* out_value, uint32_t* out_bits, size_t n,
                const BmAnd& bm) {
  using Op = ChainOp<float, NC, false, true>;
  Op op{p, in, nullptr, out_value};
  bool al = aligned16(in) && aligned16(out_value) && aligned16(out_bits);
  for (int k = 0; k < p.n_cols; ++k) al = al && aligned16(p.cols[k]);
  return launch_bits<Op, 2>(dev, op, out_bits, n, bm, al);
}

}  // namespace

extern "C" int agpu_fused_chain_pair(agpu_device* dev, int in_dtype, const void* in, const uint32_t* vin,
                                     const agpu_chain_step* steps, int n_steps, float* out_value, uint32_t* out_bits,
                                     size_t n, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (!steps || n_steps < 4 || n_steps > AGPU_CHAIN_MAX_STEPS) return AGPU_EINVAL;  // value step, STORE, RESET, compare
  if (n && (!in || !out_value || !out_bits)) return AGPU_EINVAL;
  if (in_dtype != AGPU_F32) return AGPU_EUNSUPPORTED;
  ChainProgram p{};
  p.n_steps = n_steps;
  const uint32_t* vals[4] = {vin, nullptr, nullptr, nullptr};
  int store_at = -1, reset_at = -1;
  for (int s = 0; s < n_steps; ++s) {
    const agpu_chain_step& st = steps[s];
    p.kind[s] = st.kind;
    p.op[s] = st.op;
    p.scalar[s] = st.scalar;
    p.dscalar[s] = nullptr;
    p.col[s] = 0;
    switch (st.kind) {
      case AGPU_STEP_STORE:
        if (store_at >= 0 || s == 0) return AGPU_EINVAL;  // exactly one, after at least one value step
        store_at = s;
        continue;
      case AGPU_STEP_RESET:
        if (s != store_at + 1 || store_at < 0) return AGPU_EINVAL;  // directly after the store
        reset_at = s;
        continue;
      case AGPU_STEP_UNARY:
        if (st.op != AGPU_NEG && st.op != AGPU_ABS && st.op != AGPU_SQRT) return AGPU_EUNSUPPORTED;
        break;
      case AGPU_STEP_BINARY_COLUMN:
      case AGPU_STEP_BINARY_SCALAR:
      case AGPU_STEP_BINARY_DEVSCALAR:
        if (st.op < AGPU_ADD || st.op > AGPU_MAX || (st.op >= AGPU_AND && st.op <= AGPU_XOR)) return AGPU_EUNSUPPORTED;
        break;
      case AGPU_STEP_COMPARE_COLUMN:
      case AGPU_STEP_COMPARE_SCALAR:
      case AGPU_STEP_COMPARE_DEVSCALAR:
        if (st.op < AGPU_GT || st.op > AGPU_EQ) return AGPU_EUNSUPPORTED;
        if (s != n_steps - 1 || reset_at < 0) return AGPU_EINVAL;  // the predicate chain ends the program
        break;
      default: return AGPU_EINVAL;
    }
    if (st.kind == AGPU_STEP_BINARY_DEVSCALAR || st.kind == AGPU_STEP_COMPARE_DEVSCALAR) {
      if (!st.operand) return AGPU_EINVAL;
      p.dscalar[s] = (const float*)st.operand;
    }
    if (st.kind == AGPU_STEP_BINARY_COLUMN || st.kind == AGPU_STEP_COMPARE_COLUMN) {
      if (!st.operand) return AGPU_EINVAL;
      int slot = -1;
      for (int c = 0; c < p.n_cols; ++c)
        if (p.cols[c] == (const float*)st.operand) slot = c;  // a column both chains use is loaded once
      if (slot < 0) {
        if (p.n_cols == kMaxCols) return AGPU_EUNSUPPORTED;
        slot = p.n_cols++;
        p.cols[slot] = (const float*)st.operand;
        vals[1 + slot] = st.validity;
      }
      p.col[s] = slot;
    }
  }
  const int last = steps[n_steps - 1].kind;
  if (store_at < 0 || reset_at < 0 ||
      (last != AGPU_STEP_COMPARE_COLUMN && last != AGPU_STEP_COMPARE_SCALAR && last != AGPU_STEP_COMPARE_DEVSCALAR))
    return AGPU_EINVAL;
  if (vout && !vals[0] && !vals[1] && !vals[2] && !vals[3]) return AGPU_EINVAL;
  const BmAnd bm = make_bm(vals[0], vals[1], vals[2], vals[3], vout);
  if (n_steps == 4 && steps[0].kind == AGPU_STEP_BINARY_COLUMN && steps[3].kind == AGPU_STEP_COMPARE_COLUMN) {
    const float* b = (const float*)steps[0].operand;
    const float* c = (const float*)steps[3].operand;
    const int cmp = steps[3].op;
    switch (steps[0].op) {
      case AGPU_ADD: return run_bin_cmp<OpAdd<float>>(dev, (const float*)in, b, c, cmp, out_value, out_bits, n, bm);
      case AGPU_SUB: return run_bin_cmp<OpSub<float>>(dev, (const float*)in, b, c, cmp, out_value, out_bits, n, bm);
      case AGPU_MUL: return run_bin_cmp<OpMul<float>>(dev, (const float*)in, b, c, cmp, out_value, out_bits, n, bm);
      case AGPU_DIV: return run_bin_cmp<OpDiv<float>>(dev, (const float*)in, b, c, cmp, out_value, out_bits, n, bm);
      default: break;  // rem / min / max: the interpreter
    }
  }
  switch (p.n_cols) {
    case 0: return run_pair_as<0>(dev, p, (const float*)in, out_value, out_bits, n, bm);
    case 1: return run_pair_as<1>(dev, p, (const float*)in, out_value, out_bits, n, bm);
    case 2: return run_pair_as<2>(dev, p, (const float*)in, out_value, out_bits, n, bm);
    default: return run_pair_as<3>(dev, p, (const float*)in, out_value, out_bits, n, bm);
  }
}
