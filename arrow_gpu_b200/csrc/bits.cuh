// bits.cuh — row-wise predicates whose output is a packed LSB-first bitmap (compare, fused
// expression).  Same granule/tile scheme as elementwise.cuh: lane l loads granule (base + l) as
// one 16-byte chunk per input, evaluates G predicate bits, and the S = 32/G lanes that share an
// output word merge their bit groups with log2(S) xor-shuffles; lane (l % S == 0) stores the
// word.  The reference builds the same words with workgroup-shared atomicOr + a barrier
// (compare/compute_shaders/*/cmp.wgsl:13-31).  Bits >= n are written as zero.
//
// A BitsOp provides: G, In, load(g), bits(g, In) -> low G bits, bit_at(i) -> one row's predicate.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

template <int S>
__device__ __forceinline__ uint32_t merge_bit_groups(uint32_t v) {
#pragma unroll
  for (int off = 1; off < S; off <<= 1) v |= __shfl_xor_sync(0xFFFFFFFFu, v, off);
  return v;
}

template <class Op, int UNROLL>
__global__ void __launch_bounds__(kBlock) bits_kernel(const Op op, uint32_t* __restrict__ out, const size_t n,
                                                      const BmAnd bm) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int G = Op::G;
  constexpr int S = 32 / G;  // lanes per output word
  const size_t n_gran = n / G;
  const size_t nwords = (n + 31) / 32;
  const size_t tile_gran = (size_t)kBlock * UNROLL;
  const size_t g0 = (size_t)blockIdx.x * tile_gran + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int shift = (lane % S) * G;
  if (((size_t)blockIdx.x + 1) * tile_gran <= n_gran) {
    typename Op::In in[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) in[j] = op.load(g0 + (size_t)j * kBlock);
    uint32_t b[UNROLL];
    if constexpr (IsJointOp<Op>::value) {
      op.template bits_joint<UNROLL>(g0, in, b);
    } else {
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) b[j] = op.bits(g0 + (size_t)j * kBlock, in[j]);
    }
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const uint32_t w = merge_bit_groups<S>(b[j] << shift);
      if (lane % S == 0) out[(g0 + (size_t)j * kBlock) / S] = w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const size_t g = g0 + (size_t)j * kBlock;
      uint32_t b = 0;
      if (g < n_gran) {
        b = op.bits(g, op.load(g));
      } else if (g == n_gran) {  // the < G leftover rows
        for (size_t i = g * G; i < n; ++i) b |= (uint32_t)op.bit_at(i) << (i - g * G);
      }
      const uint32_t w = merge_bit_groups<S>(b << shift);
      if (lane % S == 0 && g / S < nwords) out[g / S] = w;
    }
  }
  constexpr int tile_words = kBlock * UNROLL * G / 32;
  bm_and_tile(bm, (size_t)blockIdx.x * tile_words, tile_words, nwords);
}

// unaligned fallback: one row per lane, warp ballot -> one word per warp
template <class Op>
__global__ void __launch_bounds__(kBlock) bits_kernel_unaligned(const Op op, uint32_t* __restrict__ out,
                                                                const size_t n, const BmAnd bm) {
  pdl_wait();
  pdl_launch_dependents();
  const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
  const bool p = i < n ? op.bit_at(i) : false;
  const uint32_t w = __ballot_sync(0xFFFFFFFFu, p);
  if ((threadIdx.x & 31) == 0 && i < n) out[i >> 5] = w;
  bm_and_tile(bm, (size_t)blockIdx.x * (kBlock / 32), kBlock / 32, (n + 31) / 32);
}

template <class Op, int UNROLL = 4>
static int launch_bits(agpu_device* dev, const Op& op, uint32_t* out, size_t n, const BmAnd& bm, bool aligned) {
  if (n == 0) return 0;
  if (aligned) {
    const size_t grid = ceil_div(n, (size_t)kBlock * UNROLL * Op::G);
    if (grid > 0x7FFFFFFFull) return AGPU_EINVAL;
    if constexpr (UNROLL > 1 && !IsJointOp<Op>::value) {  // small columns: see launch_ew
      if (grid < (size_t)4 * dev->sm_count) return launch_bits<Op, 1>(dev, op, out, n, bm, aligned);
    }
    AGPU_LAUNCH_PDL(dev, (bits_kernel<Op, UNROLL>), (unsigned)grid, kBlock, 0, op, out, n, bm);
  } else {
    const size_t grid = ceil_div(n, (size_t)kBlock);
    if (grid > 0x7FFFFFFFull) return AGPU_EINVAL;
    BmAnd scalar_bm = bm;
    scalar_bm.vec = 0;
    AGPU_LAUNCH_PDL(dev, (bits_kernel_unaligned<Op>), (unsigned)grid, kBlock, 0, op, out, n, scalar_bm);
  }
  return agpu_finish_launch();
}
