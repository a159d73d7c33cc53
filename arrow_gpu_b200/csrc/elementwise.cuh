// elementwise.cuh — the one streaming kernel shape every row-wise op uses.
//
// Work unit = "granule": G consecutive rows, chosen so the WIDEST operand of the op moves as a
// 16-byte chunk per lane (G = 16 / widest element size); narrower operands of the same rows
// move as 8- or 4-byte chunks.  Lane l of a warp always owns granule (base + l), so every
// load/store instruction of a warp touches one contiguous 512/256/128-byte span: all HBM
// traffic is whole 32-byte sectors, each read or written exactly once.
//
// A CTA of 256 threads owns a tile of 256*UNROLL granules.  For full tiles all UNROLL loads of
// every input are issued before the first use (UNROLL*inputs independent 16-byte requests per
// thread in flight), then computed, then stored.  The ragged last tile takes a bounds-checked
// path and its CTA also finishes the < G leftover rows one element per thread.  The op's
// validity bitmaps for the same row range are AND-ed by the same CTA (common.cuh).
//
// An Op provides:
//   static constexpr int G;          rows per granule
//   struct In;                       registers holding one granule of every input
//   In   load(size_t g) const;       streaming loads of granule g
//   void run(size_t g, const In&) const;   compute + streaming store of granule g
//   void tail(size_t i) const;       one row, element-wise (leftover rows / unaligned buffers)
#pragma once
#include <type_traits>

#include "common.cuh"

// Ops whose per-granule work has a fixed dispatch cost (the chain interpreter) set
// `static constexpr bool JOINT = true` and get all UNROLL granules of a full tile in one call.
template <class Op, class = void> struct IsJointOp : std::false_type {};
template <class Op> struct IsJointOp<Op, std::void_t<decltype(Op::JOINT)>> : std::bool_constant<Op::JOINT> {};

// one tile (kBlock * UNROLL granules + its bitmap words) of an element-wise op
template <class Op, int UNROLL, class Bm>
__device__ __forceinline__ void ew_tile(const Op& op, const size_t n, const Bm& bm, const size_t tile, const bool last_tile) {
  constexpr int G = Op::G;
  const size_t n_gran = n / G;
  const size_t tile_gran = (size_t)kBlock * UNROLL;
  const size_t g0 = tile * tile_gran + threadIdx.x;
  if ((tile + 1) * tile_gran <= n_gran) {
    typename Op::In in[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) in[j] = op.load(g0 + (size_t)j * kBlock);
    if constexpr (IsJointOp<Op>::value) {
      op.template run_joint<UNROLL>(g0, in);  // granule j sits at g0 + j*kBlock
    } else {
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) op.run(g0 + (size_t)j * kBlock, in[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const size_t g = g0 + (size_t)j * kBlock;
      if (g < n_gran) op.run(g, op.load(g));
    }
    if (last_tile) {
      const size_t i = n_gran * G + threadIdx.x;  // n - n_gran*G < G <= 16 < kBlock
      if (i < n) op.tail(i);
    }
  }
  constexpr int tile_words = kBlock * UNROLL * G / 32;
  bm.tile(tile * tile_words, tile_words, (n + 31) / 32);
}

template <class Op, int UNROLL, class Bm>
__global__ void __launch_bounds__(kBlock) ew_kernel(const Op op, const size_t n, const Bm bm) {
  pdl_wait();               // launched with programmatic stream serialization: see AGPU_LAUNCH_PDL
  pdl_launch_dependents();  // the next kernel may start launching once every CTA of this grid is resident
  ew_tile<Op, UNROLL, Bm>(op, n, bm, blockIdx.x, blockIdx.x == gridDim.x - 1);
}

// (A persistent grid-stride form of this kernel — sm_count x resident CTAs walking the tiles — was
// measured on B200 and is 12 % SLOWER on the config-2 step (11.21 vs 10.02 ms): without software
// pipelining across tiles every CTA alternates load-wait-store, while one tile per CTA lets the
// block scheduler keep all load phases overlapped.  profiles/r02_persistent_ab.md.)

// Fallback for buffers that are not 16-byte aligned: one row per thread, same results.
template <class Op, class Bm>
__global__ void __launch_bounds__(kBlock) ew_kernel_unaligned(const Op op, const size_t n, const Bm bm) {
  pdl_wait();
  pdl_launch_dependents();
  const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) op.tail(i);
  // 256 rows = 8 bitmap words per CTA (the host clears bm.vec for this kernel)
  bm.tile((size_t)blockIdx.x * (kBlock / 32), kBlock / 32, (n + 31) / 32);
}

template <class Op, int UNROLL = 4, class Bm = BmAnd>
static int launch_ew(agpu_device* dev, const Op& op, size_t n, const Bm& bm, bool aligned) {
  if (n == 0) return 0;
  if (aligned) {
    const size_t tile_rows = (size_t)kBlock * UNROLL * Op::G;
    const size_t grid = ceil_div(n, tile_rows);
    if (grid > 0x7FFFFFFFull) return AGPU_EINVAL;
    // Small columns (the reference's own sizes: <= 16 Mi rows per op, config 1 = 1 Mi): with UNROLL
    // granules per thread a 1 Mi-row f32 op is only 256 CTAs = 1.7 per SM, so half the SMs do twice
    // the work of the others and few loads are in flight.  One granule per thread gives 4x the CTAs,
    // all resident at once: every load of the column is issued in the first microsecond.
    if constexpr (UNROLL > 1 && !IsJointOp<Op>::value) {
      if (grid < (size_t)4 * dev->sm_count) return launch_ew<Op, 1, Bm>(dev, op, n, bm, aligned);
    }
    AGPU_LAUNCH_PDL(dev, (ew_kernel<Op, UNROLL, Bm>), (unsigned)grid, kBlock, 0, op, n, bm);
  } else {
    const size_t grid = ceil_div(n, (size_t)kBlock);
    if (grid > 0x7FFFFFFFull) return AGPU_EINVAL;
    Bm scalar_bm = bm;
    scalar_bm.vec = 0;
    AGPU_LAUNCH_PDL(dev, (ew_kernel_unaligned<Op, Bm>), (unsigned)grid, kBlock, 0, op, n, scalar_bm);
  }
  return agpu_finish_launch();
}

// ---------------------------------------------------------------------------------------------
// Generic Op builders: out[i] = f(a[i]) / f(a[i], b[i]) / f(a[i], *scalar)
// ---------------------------------------------------------------------------------------------
template <int A, int B> struct MaxOf { static constexpr int v = A > B ? A : B; };

template <typename TA, typename TO, class F>
struct UnaryOp {
  static constexpr int G = 16 / MaxOf<sizeof(TA), sizeof(TO)>::v;
  const TA* a;
  TO* out;
  F f;
  struct In { Vec<TA, G> a; };
  __device__ __forceinline__ In load(size_t g) const { return In{ld_vec<TA, G>(a, g)}; }
  __device__ __forceinline__ void run(size_t g, const In& in) const {
    Vec<TO, G> o;
#pragma unroll
    for (int k = 0; k < G; ++k) o.e[k] = f(in.a.e[k]);
    st_vec<TO, G>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const { out[i] = f(a[i]); }
};

// functors that can process a packed 32-bit word of sub-word lanes at once expose
// `static constexpr bool kWord = true` and `static uint32_t word(uint32_t, uint32_t)`
template <class F, class = void> struct HasWordOp : std::false_type {};
template <class F> struct HasWordOp<F, std::void_t<decltype(F::kWord)>> : std::bool_constant<F::kWord> {};

template <typename TA, typename TB, typename TO, class F>
struct BinaryOp {
  static constexpr int G = 16 / MaxOf<MaxOf<sizeof(TA), sizeof(TB)>::v, sizeof(TO)>::v;
  const TA* a;
  const TB* b;
  TO* out;
  F f;
  struct In { Vec<TA, G> a; Vec<TB, G> b; };
  __device__ __forceinline__ In load(size_t g) const { return In{ld_vec<TA, G>(a, g), ld_vec<TB, G>(b, g)}; }
  __device__ __forceinline__ void run(size_t g, const In& in) const {
    Vec<TO, G> o;
    if constexpr (HasWordOp<F>::value && std::is_same<TA, TB>::value && std::is_same<TA, TO>::value && sizeof(TA) * G == 16) {
      uint32_t wa[4], wb[4], wo[4];
      memcpy(wa, &in.a, 16);
      memcpy(wb, &in.b, 16);
#pragma unroll
      for (int k = 0; k < 4; ++k) wo[k] = F::word(wa[k], wb[k]);
      memcpy(&o, wo, 16);
    } else {
#pragma unroll
      for (int k = 0; k < G; ++k) o.e[k] = f(in.a.e[k], in.b.e[k]);
    }
    st_vec<TO, G>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const { out[i] = f(a[i], b[i]); }
};

// rhs is a one-element device array (the reference passes scalars as length-1 arrays)
template <typename T, class F>
struct ScalarOp {
  static constexpr int G = 16 / sizeof(T);
  const T* a;
  const T* s;
  T* out;
  F f;
  struct In { Vec<T, G> a; T s; };
  __device__ __forceinline__ In load(size_t g) const { return In{ld_vec<T, G>(a, g), __ldg(s)}; }
  __device__ __forceinline__ void run(size_t g, const In& in) const {
    Vec<T, G> o;
#pragma unroll
    for (int k = 0; k < G; ++k) o.e[k] = f(in.a.e[k], in.s);
    st_vec<T, G>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const { out[i] = f(a[i], __ldg(s)); }
};
