// routines.cu — merge (mask select), take (gather), put (scatter) and filter (compaction):
// crates/routines.  merge/take/put follow the reference shaders; filter is new surface
//  built from warp-level prefix sums with shared-memory staging.
#include <stdlib.h>

#include "bits.cuh"
#include "elementwise.cuh"

namespace {

// ============================================================================================
// merge: routines/compute_shaders/{32bit,16bit,8bit,bool}/merge.wgsl; validity merge.rs:17-86
// ============================================================================================
// vout = ((va & m) | (vb & ~m)) & vmask, missing bitmap = all ones (Q7) — one pass over the
// same tile instead of the reference's up to four dispatches.
struct BmMerge {
  const uint32_t *va, *vb, *mask, *vmask;
  uint32_t* out;
  int vec;
  size_t last_word;     // word holding row n-1: its padding bits are cleared
  uint32_t last_mask;
  __device__ __forceinline__ uint32_t word(size_t w) const {
    const uint32_t m = mask[w];
    const uint32_t a = va ? va[w] : 0xFFFFFFFFu;
    const uint32_t b = vb ? vb[w] : 0xFFFFFFFFu;
    const uint32_t vm = vmask ? vmask[w] : 0xFFFFFFFFu;
    const uint32_t r = ((a & m) | (b & ~m)) & vm;
    return w == last_word ? (r & last_mask) : r;
  }
  __device__ __forceinline__ void tile(size_t w0, int tile_words, size_t nwords) const {
    if (!out) return;
    for (int q = threadIdx.x; q < tile_words; q += blockDim.x) {
      const size_t w = w0 + q;
      if (w < nwords) out[w] = word(w);
    }
  }
};

template <typename U>  // U = unsigned type of the element width (select is type-agnostic)
struct MergeOp {
  static constexpr int G = 16 / sizeof(U);
  const U* a;
  const U* b;
  const uint32_t* mask;
  U* out;
  struct In { Vec<U, G> a, b; uint32_t m; };
  __device__ __forceinline__ In load(size_t g) const {
    const size_t bit = g * G;  // G divides 32: a granule's mask bits never straddle a word
    return In{ld_vec<U, G>(a, g), ld_vec<U, G>(b, g), __ldg(mask + (bit >> 5)) >> (bit & 31)};
  }
  __device__ __forceinline__ void run(size_t g, const In& in) const {
    Vec<U, G> o;
#pragma unroll
    for (int k = 0; k < G; ++k) o.e[k] = (in.m >> k) & 1u ? in.a.e[k] : in.b.e[k];
    st_vec<U, G>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const {
    out[i] = (mask[i >> 5] >> (i & 31)) & 1u ? a[i] : b[i];
  }
};

// bool merge: data and validity are both bitmaps over the same words
struct BoolMergeOp {
  const uint32_t *a, *b, *mask;
  uint32_t* out;
  BmMerge v;
  size_t last_word;
  uint32_t last_mask;
};

__global__ void __launch_bounds__(kBlock) bool_merge_kernel(const BoolMergeOp op, const size_t nwords) {
  const size_t w = (size_t)blockIdx.x * kBlock + threadIdx.x;
  if (w >= nwords) return;
  const uint32_t m = op.mask[w];
  uint32_t r = (op.a[w] & m) | (op.b[w] & ~m);
  uint32_t vr = op.v.out ? op.v.word(w) : 0u;
  if (w == op.last_word) { r &= op.last_mask; vr &= op.last_mask; }
  op.out[w] = r;
  if (op.v.out) op.v.out[w] = vr;
}

template <typename U>
int run_merge(agpu_device* dev, const void* a, const void* b, const uint32_t* mask, void* out, size_t n,
              const BmMerge& bm) {
  MergeOp<U> op{(const U*)a, (const U*)b, mask, (U*)out};
  return launch_ew<MergeOp<U>, 4, BmMerge>(dev, op, n, bm, aligned16(a) && aligned16(b) && aligned16(out));
}

// ============================================================================================
// take: routines/compute_shaders/32bit/take.wgsl:13-17, bool/take.wgsl:13-33
// ============================================================================================
// granule = 4 output rows: one 16-byte chunk of indices per lane, 4 gathers, one chunk store.
// Gathers go through the read-only path without L1 allocation and ask L2 for the smallest fill
// it offers (64 bytes instead of the default 128): a random 4-byte gather otherwise drags a whole
// 128-byte line out of HBM.
template <typename U>
__device__ __forceinline__ U ld_gather(const U* p) {
  uint32_t r;
  if constexpr (sizeof(U) == 4) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.b32 %0, [%1];" : "=r"(r) : "l"(p));
  else if constexpr (sizeof(U) == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u16 %0, [%1];" : "=r"(r) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.L2::64B.u8 %0, [%1];" : "=r"(r) : "l"(p));
  return (U)r;
}

template <typename U>
struct TakeOp {
  static constexpr int G = 4;
  const U* src;
  size_t src_len;
  const uint32_t* idx;
  U* out;
  struct In { Vec<uint32_t, 4> i; };
  __device__ __forceinline__ U fetch(uint32_t i) const { return i < src_len ? ld_gather<U>(src + i) : (U)0; }
  __device__ __forceinline__ In load(size_t g) const { return In{ld_vec<uint32_t, 4>(idx, g)}; }
  __device__ __forceinline__ void run(size_t g, const In& in) const {
    Vec<U, 4> o;
#pragma unroll
    for (int k = 0; k < 4; ++k) o.e[k] = fetch(in.i.e[k]);
    st_vec<U, 4>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t j) const { out[j] = fetch(idx[j]); }
};

// bit gather (bool data, or validity alone): bit j of out = bit idx[j] of src
struct TakeBitsOp {
  static constexpr int G = 4;
  const uint32_t* src;
  size_t src_len;
  const uint32_t* idx;
  struct In { Vec<uint32_t, 4> i; };
  __device__ __forceinline__ uint32_t fetch(uint32_t i) const {
    return i < src_len ? (__ldg(src + (i >> 5)) >> (i & 31)) & 1u : 0u;
  }
  __device__ __forceinline__ In load(size_t g) const { return In{ld_vec<uint32_t, 4>(idx, g)}; }
  __device__ __forceinline__ uint32_t bits(size_t, const In& in) const {
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) m |= fetch(in.i.e[k]) << k;
    return m;
  }
  __device__ __forceinline__ bool bit_at(size_t j) const { return fetch(idx[j]); }
};

// values and their validity bits in one pass over the indices
template <typename U>
struct TakeWithValidityOp {
  static constexpr int G = 4;
  TakeOp<U> val;
  TakeBitsOp bit;
  using In = typename TakeOp<U>::In;
  __device__ __forceinline__ In load(size_t g) const { return val.load(g); }
  __device__ __forceinline__ uint32_t bits(size_t g, const In& in) const {
    val.run(g, in);
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) m |= bit.fetch(in.i.e[k]) << k;
    return m;
  }
  __device__ __forceinline__ bool bit_at(size_t j) const {
    val.tail(j);
    return bit.fetch(bit.idx[j]);
  }
};

template <typename U>
int run_take(agpu_device* dev, const void* src, size_t src_len, const uint32_t* idx, void* out, size_t m,
             const uint32_t* vsrc, uint32_t* vout) {
  const bool al = aligned16(idx) && aligned16(out);
  TakeOp<U> val{(const U*)src, src_len, idx, (U*)out};
  BmAnd none{};
  if (vsrc && vout) {
    TakeWithValidityOp<U> op{val, TakeBitsOp{vsrc, src_len, idx}};
    return launch_bits(dev, op, vout, m, none, al);
  }
  return launch_ew(dev, val, m, none, al);
}

// ============================================================================================
// take with GLOBAL row indices over row-range shards
// ============================================================================================
// Each shard pointer is local memory or another GPU's memory opened through CUDA IPC; the gather
// loads go straight over NVLink/NVSwitch.  Same granule scheme as TakeOp: 4 indices per lane as
// one 16-byte chunk, 4 gathers, one chunk store; validity bits with the compare shuffle network.
struct ShardTable {
  const void* values[AGPU_MAX_SHARDS];
  const uint32_t* validity[AGPU_MAX_SHARDS];
  unsigned long long begin[AGPU_MAX_SHARDS + 1];
  int n;
  int has_validity;
};

template <typename U>
struct TakeShardedOp {
  static constexpr int G = 4;
  ShardTable t;
  const uint32_t* idx;
  U* out;
  struct In { Vec<uint32_t, 4> i; };
  __device__ __forceinline__ int shard_of(uint32_t r) const {
    int s = 0;
#pragma unroll
    for (int k = 1; k < AGPU_MAX_SHARDS; ++k) s += (k < t.n && r >= t.begin[k]) ? 1 : 0;
    return s;
  }
  __device__ __forceinline__ U fetch(uint32_t r, uint32_t& vbit) const {
    vbit = 0;
    if (r >= t.begin[t.n]) return (U)0;
    const int s = shard_of(r);
    const uint64_t local = r - t.begin[s];
    if (t.has_validity) {
      const uint32_t* v = t.validity[s];
      vbit = v ? (v[local >> 5] >> (local & 31)) & 1u : 1u;
    }
    return static_cast<const U*>(t.values[s])[local];
  }
  __device__ __forceinline__ In load(size_t g) const { return In{ld_vec<uint32_t, 4>(idx, g)}; }
  // BitsOp interface (values are stored as a side effect, validity bits returned)
  __device__ __forceinline__ uint32_t bits(size_t g, const In& in) const {
    Vec<U, 4> o;
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t vb;
      o.e[k] = fetch(in.i.e[k], vb);
      m |= vb << k;
    }
    st_vec<U, 4>(out, g, o);
    return m;
  }
  __device__ __forceinline__ bool bit_at(size_t j) const {
    uint32_t vb;
    out[j] = fetch(idx[j], vb);
    return vb;
  }
  // ew Op interface (no validity)
  __device__ __forceinline__ void run(size_t g, const In& in) const { (void)bits(g, in); }
  __device__ __forceinline__ void tail(size_t j) const { (void)bit_at(j); }
};

template <typename U>
int run_take_sharded(agpu_device* dev, const ShardTable& t, const uint32_t* idx, void* out, size_t m, uint32_t* vout) {
  TakeShardedOp<U> op{t, idx, (U*)out};
  BmAnd none{};
  const bool al = aligned16(idx) && aligned16(out);
  if (t.has_validity && vout) return launch_bits(dev, op, vout, m, none, al);
  return launch_ew(dev, op, m, none, al);
}

// ============================================================================================
// put: routines/compute_shaders/32bit/put.wgsl:17-23, bool/put.wgsl:17-34
// ============================================================================================
// granule = 4 (source index, destination index) pairs: two 16-byte index chunks per lane, all
// index chunks of the tile in flight first, then 4 gathers and 4 scattered stores per granule
template <typename U>
struct PutOp {
  static constexpr int G = 4;
  const U* src;
  size_t src_len;
  const uint32_t* si;
  U* dst;
  size_t dst_len;
  const uint32_t* di;
  struct In { Vec<uint32_t, 4> s, d; };
  __device__ __forceinline__ In load(size_t g) const { return In{ld_vec<uint32_t, 4>(si, g), ld_vec<uint32_t, 4>(di, g)}; }
  // wgpu robust buffer access (which the reference's put.wgsl relies on): an out-of-range source
  // index reads zero, an out-of-range destination index writes nothing
  __device__ __forceinline__ U fetch(uint32_t s) const { return s < src_len ? src[s] : (U)0; }
  __device__ __forceinline__ void run(size_t, const In& in) const {
    U v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = fetch(in.s.e[k]);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (in.d.e[k] < dst_len) dst[in.d.e[k]] = v[k];
  }
  __device__ __forceinline__ void tail(size_t j) const {
    const uint32_t d = di[j];
    if (d < dst_len) dst[d] = fetch(si[j]);
  }
};

__global__ void __launch_bounds__(kBlock) put_bits_kernel(const uint32_t* __restrict__ src, const size_t src_len,
                                                          const uint32_t* __restrict__ si, uint32_t* dst, const size_t dst_len,
                                                          const uint32_t* __restrict__ di, const size_t m) {
  const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= m) return;  // (the reference lacks this bound: Q9)
  const uint32_t s = si[i], d = di[i];
  if (d >= dst_len) return;
  const uint32_t bit = s < src_len ? (src[s >> 5] >> (s & 31)) & 1u : 0u;
  if (bit) atomicOr(dst + (d >> 5), 1u << (d & 31));
  else atomicAnd(dst + (d >> 5), ~(1u << (d & 31)));
}

// ============================================================================================
// filter (compaction)
// ============================================================================================
// Tile = 4096 rows = 128 mask words; group = 64 tiles.
// Pass 1 (count): one CTA per group, one warp per 8 tiles; each lane popcounts a 16-byte chunk of
//   (mask & vmask) per tile (all 8 loads in flight first), a shuffle reduce gives the tile count;
//   the CTA adds its 64 tile counts into one group total.  No atomics.  Traffic: N/8 bytes.
// Pass 2 (scan): one CTA scans the group totals (N / 262144 values) -> 64-bit group offsets + the
//   grand total.  Both run inside agpu_filter_count so the caller can size the output.
// Pass 3 (scatter): one CTA per tile.  Warp 0 prefix-sums the popcounts of the tile's 128
//   selection words and adds the counts of the preceding tiles of its group to the group offset.
//   Every thread loads its rows as coalesced 16-byte granules, a selected row's slot is
//   word_prefix + popc(selection bits below it); values are staged in shared memory shifted by
//   (output offset mod G) so that the CTA can stream them out as aligned 16-byte vectors.
constexpr int kFilterTileRows = 4096;
constexpr int kFilterTileWords = kFilterTileRows / 32;
constexpr int kFilterGroupTiles = 64;

struct FilterScratch {
  uint64_t* group_offsets;  // [groups]
  uint32_t* counts;         // [tiles]
};

inline size_t filter_tiles(size_t n) { return ceil_div(n, (size_t)kFilterTileRows); }
inline size_t filter_groups(size_t n) { return ceil_div(filter_tiles(n), (size_t)kFilterGroupTiles); }
inline FilterScratch filter_scratch(void* p, size_t n) {
  FilterScratch s;
  s.group_offsets = (uint64_t*)p;
  s.counts = (uint32_t*)((char*)p + ((filter_groups(n) * 8 + 15) / 16) * 16);
  return s;
}

__device__ __forceinline__ uint32_t sel_word(const uint32_t* mask, const uint32_t* vmask, size_t w, size_t nwords,
                                             size_t n) {
  if (w >= nwords) return 0u;
  uint32_t s = mask[w];
  if (vmask) s &= vmask[w];
  if (w == nwords - 1 && (n & 31)) s &= (1u << (n & 31)) - 1u;
  return s;
}

// exclusive scan of the group totals (in place) by ONE CTA of BLOCK threads, 8 values per thread per round
template <int BLOCK>
__device__ __forceinline__ void scan_group_totals(uint64_t* groups, const size_t n_groups, unsigned long long* total) {
  __shared__ uint64_t scan_warp_tot[BLOCK / 32];
  __shared__ uint64_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (size_t base = 0; base < n_groups; base += (size_t)BLOCK * 8) {
    const size_t g0 = base + (size_t)threadIdx.x * 8;
    uint64_t c[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      c[k] = g0 + k < n_groups ? __ldcg(groups + g0 + k) : 0ull;   // written by other CTAs of this launch: bypass L1
      sum += c[k];
    }
    uint64_t incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint64_t v = __shfl_up_sync(0xFFFFFFFFu, incl, off);
      if (lane >= off) incl += v;
    }
    if (lane == 31) scan_warp_tot[warp] = incl;
    __syncthreads();
    uint64_t before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < BLOCK / 32; ++k) {
      const uint64_t t = scan_warp_tot[k];
      if (k < warp) before += t;
      all += t;
    }
    uint64_t excl = carry_s + (incl - sum) + before;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (g0 + k < n_groups) groups[g0 + k] = excl;
      excl += c[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_s += all;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

// The scan of the group totals rides in the count kernel: every CTA takes a ticket after publishing
// its total, and the CTA that draws the last one scans all of them (threadfence reduction pattern).
// One launch and one ~5 us single-CTA kernel less per filter; `ticket` lives in the device handle,
// is zero between launches (the last CTA resets it) and launches of one handle are stream-ordered.
__global__ void __launch_bounds__(kBlock) filter_count_kernel(const uint32_t* __restrict__ mask,
                                                              const uint32_t* __restrict__ vmask, const size_t n,
                                                              uint32_t* __restrict__ counts,
                                                              uint64_t* group_totals, const int vec,
                                                              unsigned int* ticket, unsigned long long* total,
                                                              const ExchangePost post) {
  __shared__ uint32_t warp_tot[kBlock / 32];
  __shared__ bool is_last;
  const size_t nwords = (n + 31) / 32;
  const size_t tiles = (n + kFilterTileRows - 1) / kFilterTileRows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int TPW = kFilterGroupTiles / (kBlock / 32);  // tiles per warp = 8
  const size_t t0 = (size_t)blockIdx.x * kFilterGroupTiles + (size_t)warp * TPW;
  uint32_t c[TPW];
  if (vec && (t0 + TPW) * kFilterTileRows <= n) {  // all 8 tiles full: vector loads, issued up front
    uint4 m[TPW];
#pragma unroll
    for (int k = 0; k < TPW; ++k) m[k] = __ldcs(reinterpret_cast<const uint4*>(mask + (t0 + k) * kFilterTileWords) + lane);
    if (vmask) {
#pragma unroll
      for (int k = 0; k < TPW; ++k) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(vmask + (t0 + k) * kFilterTileWords) + lane);
        m[k].x &= v.x; m[k].y &= v.y; m[k].z &= v.z; m[k].w &= v.w;
      }
    }
#pragma unroll
    for (int k = 0; k < TPW; ++k) c[k] = __popc(m[k].x) + __popc(m[k].y) + __popc(m[k].z) + __popc(m[k].w);
  } else {
#pragma unroll
    for (int k = 0; k < TPW; ++k) {
      const size_t w0 = (t0 + k) * kFilterTileWords + (size_t)lane * 4;
      c[k] = 0;
      for (int j = 0; j < 4; ++j) c[k] += __popc(sel_word(mask, vmask, w0 + j, nwords, n));
    }
  }
  uint32_t wsum = 0;
#pragma unroll
  for (int k = 0; k < TPW; ++k) {
    uint32_t x = c[k];
#pragma unroll
    for (int off = 16; off; off >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, off);
    if (lane == 0 && t0 + k < tiles) counts[t0 + k] = x;
    wsum += x;
  }
  if (lane == 0) warp_tot[warp] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t tot = 0;
    for (int w = 0; w < kBlock / 32; ++w) tot += warp_tot[w];
    group_totals[blockIdx.x] = tot;
    __threadfence();                                        // the total is visible before the ticket
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();                                          // every other CTA's total is visible now
  scan_group_totals<kBlock>(group_totals, gridDim.x, total);
  if (threadIdx.x == 0) *ticket = 0u;
  if (post.world) {  // sharded caller: this shard's total goes straight to every peer (agpu_filter_count_post)
    __syncthreads();
    if (threadIdx.x < 32) exchange_post_lane(post, __ldcg(total), threadIdx.x);
  }
}

// store `val` at shared-memory byte address `sa` and advance `sa` by one element iff bit != 0
template <typename U>
__device__ __forceinline__ void stage_if(uint32_t& sa, U val, uint32_t bit) {
  if constexpr (sizeof(U) == 4) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t@p add.u32 %0, %0, 4;\n\t}"
                 : "+r"(sa) : "r"((uint32_t)val), "r"(bit) : "memory");
  } else if constexpr (sizeof(U) == 2) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u16 [%0], %1;\n\t@p add.u32 %0, %0, 2;\n\t}"
                 : "+r"(sa) : "h"((uint16_t)val), "r"(bit) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u8 [%0], %1;\n\t@p add.u32 %0, %0, 1;\n\t}"
                 : "+r"(sa) : "r"((uint32_t)val), "r"(bit) : "memory");
  }
}

// sub-word rows: the lane stays inside its packed 32-bit word until the predicated store (taking
// the lanes out up front costs 64 live registers for a 1-byte tile and spills)
template <int BYTES, int SHIFT>
__device__ __forceinline__ void stage_lane_if(uint32_t& sa, uint32_t word, uint32_t bit) {
  if constexpr (BYTES == 1) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tsetp.ne.u32 p, %2, 0;\n\tshr.u32 t, %1, %3;\n\t@p st.shared.u8 [%0], t;\n\t@p add.u32 %0, %0, 1;\n\t}"
                 : "+r"(sa) : "r"(word), "r"(bit), "n"(SHIFT) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tsetp.ne.u32 p, %2, 0;\n\tshr.u32 t, %1, %3;\n\t@p st.shared.u16 [%0], t;\n\t@p add.u32 %0, %0, 2;\n\t}"
                 : "+r"(sa) : "r"(word), "r"(bit), "n"(SHIFT) : "memory");
  }
}
template <typename U, int K>
__device__ __forceinline__ void stage_lanes(uint32_t& sa, const uint32_t (&w)[4], uint32_t bits) {
  constexpr int L = 4 / sizeof(U);  // lanes per word
  if constexpr (K < 16 / (int)sizeof(U)) {
    stage_lane_if<sizeof(U), (K % L) * 8 * (int)sizeof(U)>(sa, w[K / L], bits & (1u << K));
    stage_lanes<U, K + 1>(sa, w, bits);
  }
}

// One CTA compacts a "super tile" of M = 4/sizeof(U) count-tiles, i.e. always 16 KiB of rows
// (4096 x 4-byte, 8192 x 2-byte or 16384 x 1-byte rows): the per-tile barriers and prefix sums are
// amortised over the same number of bytes for every element width.
template <typename U, bool HAS_V, int BLOCK>
__global__ void __launch_bounds__(BLOCK, (sizeof(U) == 1 ? 1536 : 2048) / BLOCK) filter_scatter_kernel(const U* __restrict__ src,
                                                                const uint32_t* __restrict__ vsrc,
                                                                const uint32_t* __restrict__ mask,
                                                                const uint32_t* __restrict__ vmask, const size_t n,
                                                                const uint32_t* __restrict__ counts,
                                                                const uint64_t* __restrict__ group_offsets,
                                                                U* __restrict__ out, uint32_t* vout, const uint64_t cap) {
  constexpr int G = 16 / sizeof(U);                   // rows per 16-byte granule
  constexpr int M = 4 / sizeof(U);                    // count-tiles per super tile
  constexpr int ROWS = kFilterTileRows * M;           // rows per super tile (16 KiB of rows)
  constexpr int WORDS = ROWS / 32;                    // selection words per super tile
  constexpr int WPL = WORDS / 32;                     // selection words per lane of warp 0
  constexpr int GPT = ROWS / G / BLOCK;               // granules per thread (= 4 for BLOCK 256)
  __shared__ __align__(16) U stage[ROWS + G];
  __shared__ uint32_t sel[WORDS];
  __shared__ uint32_t pre[WORDS];
  __shared__ uint32_t vstage[HAS_V ? WORDS + 1 : 1];
  __shared__ uint8_t vbyte[HAS_V ? ROWS + G : 1];     // validity of each compacted row, one byte each
  __shared__ uint64_t off_s;
  __shared__ uint32_t count_s;

  const size_t tile = blockIdx.x;                     // super tile index
  const size_t nwords = (n + 31) / 32;
  const size_t w0 = tile * WORDS;
  const size_t row0 = tile * ROWS;
  const bool full = row0 + ROWS <= n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // everything that comes from global memory is requested up front: the rows (independent of
  // the selection bits), the selection words, and the counts that place this tile in the output
  Vec<U, G> v[GPT];
  if (full) {
#pragma unroll
    for (int j = 0; j < GPT; ++j) v[j] = ld_vec<U, G>(src + row0, (size_t)j * BLOCK + threadIdx.x);
  }
  uint32_t before = 0;
  uint64_t goff = 0;
  if (warp == 1) {  // warp 1: output offset = group offset + counts of the earlier tiles of the group
    const size_t t0 = tile * M;  // first count-tile of this super tile (M divides the group size)
    const size_t gstart = t0 / kFilterGroupTiles * kFilterGroupTiles;
    if (gstart + lane < t0) before += counts[gstart + lane];
    if (gstart + 32 + lane < t0) before += counts[gstart + 32 + lane];
    if (lane == 0) goff = group_offsets[t0 / kFilterGroupTiles];
  }
  if (full) {  // no ragged word in this tile: plain loads
    for (int w = threadIdx.x; w < WORDS; w += BLOCK) {
      uint32_t x = mask[w0 + w];
      if (vmask) x &= vmask[w0 + w];
      sel[w] = x;
    }
  } else {
    for (int w = threadIdx.x; w < WORDS; w += BLOCK) sel[w] = sel_word(mask, vmask, w0 + w, nwords, n);
  }
  if (warp == 1) {
#pragma unroll
    for (int off = 16; off; off >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, off);
    if (lane == 0) off_s = goff + before;
  }
  __syncthreads();
  if (warp == 0) {  // warp 0: exclusive prefix of the popcounts, WPL consecutive words per lane
    uint32_t c[WPL], s = 0;
#pragma unroll
    for (int k = 0; k < WPL; ++k) { c[k] = __popc(sel[lane * WPL + k]); s += c[k]; }
    uint32_t incl = s;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, off);
      if (lane >= off) incl += x;
    }
    uint32_t e = incl - s;
#pragma unroll
    for (int k = 0; k < WPL; ++k) { pre[lane * WPL + k] = e; e += c[k]; }
    if (lane == 31) count_s = incl;
  }
  __syncthreads();
  const uint64_t off = off_s;
  // rows past the capacity of the output buffers are dropped (a caller that sized the output from
  // a selectivity estimate learns the real total from the count pass and retries)
  const uint32_t count = off >= cap ? 0u : (uint32_t)min((uint64_t)count_s, cap - off);
  if (count == 0) return;  // uniform for the CTA
  const bool vec_out = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
  const uint32_t lead = vec_out ? (uint32_t)(off % G) : 0u;  // shift so that 16-byte vectors line up

  const uint32_t stage_sa = (uint32_t)__cvta_generic_to_shared(stage);
  const uint32_t vbyte_sa = (uint32_t)__cvta_generic_to_shared(vbyte);
  if (full) {
    // branch-free: every row's slot is computed, the store is predicated on its selection bit
    // (no divergent branches / reconvergence barriers in the hot loop)
#pragma unroll
    for (int j = 0; j < GPT; ++j) {
      const int r = (j * BLOCK + threadIdx.x) * G;  // first row of the granule within the tile
      const uint32_t sw = sel[r >> 5];
      const uint32_t bits = sw >> (r & 31);
      uint32_t pos = lead + pre[r >> 5] + __popc(sw & ((1u << (r & 31)) - 1u));
      uint32_t vw = 0;
      if (HAS_V) vw = vsrc[(row0 + r) >> 5] >> (r & 31);
      // per row: one predicated store and one predicated address bump (written in PTX: the
      // compiler's form of `pos += take` is a 3-instruction select-and-add per row)
      uint32_t sa = stage_sa + pos * (uint32_t)sizeof(U);
      uint32_t va = vbyte_sa + pos;
      if constexpr (sizeof(U) < 4) {
        uint32_t w[4];
        memcpy(w, &v[j], 16);
        // (two independent address chains per granule instead of one serial chain of predicated
        // bumps: measured, no difference — 162.6 vs 164.2 us for 1-byte rows)
        stage_lanes<U, 0>(sa, w, bits);
        if (HAS_V) {
#pragma unroll
          for (int k = 0; k < G; ++k) stage_if<uint8_t>(va, (uint8_t)((vw >> k) & 1u), bits & (1u << k));
        }
      } else {
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const uint32_t bit = bits & (1u << k);
          stage_if<U>(sa, v[j].e[k], bit);
          if (HAS_V) stage_if<uint8_t>(va, (uint8_t)((vw >> k) & 1u), bit);
        }
      }
    }
  } else {
#pragma unroll 1
    for (int j = 0; j < GPT; ++j) {
      const int r = (j * BLOCK + threadIdx.x) * G;
      const uint32_t sw = sel[r >> 5];
      const uint32_t bits = (sw >> (r & 31)) & ((G == 32) ? 0xFFFFFFFFu : ((1u << G) - 1u));
      if (bits == 0) continue;
      uint32_t pos = lead + pre[r >> 5] + __popc(sw & ((1u << (r & 31)) - 1u));
      uint32_t vw = 0;
      if (HAS_V) vw = vsrc[(row0 + r) >> 5] >> (r & 31);
      for (int k = 0; k < G; ++k) {
        if ((bits >> k) & 1u) {
          stage[pos] = src[row0 + r + k];
          if (HAS_V) vbyte[pos] = (uint8_t)((vw >> k) & 1u);
          ++pos;
        }
      }
    }
  }
  __syncthreads();

  if (vec_out) {
    // staged elements [lead, lead+count) map to out[off .. off+count); vector q covers staged
    // elements [q*G, q*G+G) = global elements (off - lead) + q*G .., 16-byte aligned
    U* gbase = out + (off - lead);
    const uint32_t end = lead + count;
    const uint32_t nvec = (end + G - 1) / G;
    for (uint32_t q = threadIdx.x; q < nvec; q += BLOCK) {
      const uint32_t e0 = q * G;
      if (e0 >= lead && e0 + G <= end) {
        Vec<U, G> t = *reinterpret_cast<const Vec<U, G>*>(stage + e0);
        st_vec<U, G>(gbase, q, t);
      } else {
        for (uint32_t k = 0; k < (uint32_t)G; ++k)
          if (e0 + k >= lead && e0 + k < end) gbase[e0 + k] = stage[e0 + k];
      }
    }
  } else {
    for (uint32_t i = threadIdx.x; i < count; i += BLOCK) out[off + i] = stage[i];
  }
  if (HAS_V) {
    // the validity bytes of 32 consecutive compacted rows become one word with a warp ballot
    const uint32_t lw_n = (count + 31) / 32;
    for (uint32_t w = warp; w < lw_n; w += BLOCK / 32) {
      const uint32_t i = w * 32 + lane;
      const uint32_t word = __ballot_sync(0xFFFFFFFFu, i < count && vbyte[lead + i] != 0);
      if (lane == 0) vstage[w] = word;
    }
    __syncthreads();
    const uint32_t s = (uint32_t)(off & 31);
    for (uint32_t w = threadIdx.x; w < lw_n; w += BLOCK) {
      const uint32_t val = vstage[w];
      const uint64_t gw = (off >> 5) + w;
      if (val << s) atomicOr(vout + gw, val << s);
      if (s && (val >> (32 - s))) atomicOr(vout + gw + 1, val >> (32 - s));
    }
  }
}

// --------------------------------------------------------------------------------------------
// validity of the compacted rows, as its own bitmap-only pass
// --------------------------------------------------------------------------------------------
// Compacting the validity bits inside the value scatter costs a byte store per kept row plus a
// ballot pass and a third barrier (f32 + validity ran at 0.67 of the roofline against 1.0 without).
// The bits do not need the values: one thread takes one 32-row word of (selection, validity),
// extracts the validity bits of the selected rows with a parallel-suffix "compress" (Hacker's
// Delight 7-4: 5 rounds of shift/xor, no table, no loop over rows), and ORs the result at the
// word's output bit position — the tile offsets are the ones the count pass already produced.
// 0.25 B/row of traffic; the kernel is ALU-bound and a fraction of the value pass.
__device__ __forceinline__ uint32_t compress_bits(uint32_t x, uint32_t m) {
  x &= m;
  uint32_t mk = ~m << 1;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    uint32_t mp = mk ^ (mk << 1);
    mp ^= mp << 2;
    mp ^= mp << 4;
    mp ^= mp << 8;
    mp ^= mp << 16;
    const uint32_t mv = mp & m;
    m = (m ^ mv) | (mv >> (1 << i));
    const uint32_t t = x & mv;
    x = (x ^ t) | (t >> (1 << i));
    mk &= ~mp;
  }
  return x;
}

// One thread owns kBitsWordsPerThread consecutive 32-row words (256 rows), so that it emits whole
// output words by itself: compacted bits are appended to a 64-bit accumulator and every time 32 are
// complete the low word goes straight to global memory with a plain store.  Only the first word (its
// low bits belong to the previous thread) and the final partial word are OR-ed in atomically: two
// reductions per 256 rows.  (A first version with one word per thread and shared-memory atomicOr
// for every word was LSU-bound — ATOMS costs 2 cycles per lane on this part — and took 90 us for
// 256 Mi rows; profiles/r02_filter_validity.md.)
constexpr int kBitsWordsPerThread = 8;
constexpr int kBitsBlock = 256;
constexpr int kBitsCtaWords = kBitsBlock * kBitsWordsPerThread;          // 2048 words = 65536 rows
constexpr int kBitsTilesPerCta = kBitsCtaWords / kFilterTileWords;        // = 16 count-tiles (divides the group size)
static_assert(kFilterGroupTiles % kBitsTilesPerCta == 0, "a CTA's tiles must lie in one group");

__device__ __forceinline__ void emit_bits(uint32_t* vout, uint64_t word_index, uint32_t word, bool atomic, uint64_t cap) {
  const uint64_t bit0 = word_index * 32;
  if (bit0 >= cap) return;                                            // past the capacity of the output
  if (cap - bit0 < 32) word &= (1u << (uint32_t)(cap - bit0)) - 1u;   // the capacity ends inside this word
  if (atomic) {
    if (word) atomicOr(vout + word_index, word);
  } else {
    vout[word_index] = word;
  }
}

template <bool CLIP>
__device__ __forceinline__ void compact_words(uint32_t* vout, uint64_t wp, const uint32_t lead,
                                              const uint32_t (&sel)[kBitsWordsPerThread], const uint32_t (&val)[kBitsWordsPerThread],
                                              const uint64_t cap) {
  uint32_t* out = vout + wp;
  uint64_t acc = 0;
  uint32_t nacc = lead;
  bool first = true;
#pragma unroll
  for (int k = 0; k < kBitsWordsPerThread; ++k) {
    const uint32_t cv = compress_bits(val[k], sel[k]);
    acc |= (uint64_t)cv << nacc;
    nacc += __popc(sel[k]);
    if (nacc >= 32) {
      const bool shared_word = first && lead != 0;   // its low bits belong to the previous thread
      if constexpr (CLIP) {
        emit_bits(vout, wp, (uint32_t)acc, shared_word, cap);
        ++wp;
      } else {
        if (shared_word) atomicOr(out, (uint32_t)acc);
        else *out = (uint32_t)acc;
        ++out;
      }
      acc >>= 32;
      nacc -= 32;
      first = false;
    }
  }
  if (nacc > (first ? lead : 0u)) {  // partial last word: the next thread fills the rest
    if constexpr (CLIP) emit_bits(vout, wp, (uint32_t)acc, true, cap);
    else atomicOr(out, (uint32_t)acc);
  }
}

__global__ void __launch_bounds__(kBitsBlock) filter_bits_kernel(const uint32_t* __restrict__ vsrc,
                                                                 const uint32_t* __restrict__ mask,
                                                                 const uint32_t* __restrict__ vmask, const size_t n,
                                                                 const uint32_t* __restrict__ counts,
                                                                 const uint64_t* __restrict__ group_offsets,
                                                                 uint32_t* vout, const uint64_t cap, const int vec) {
  __shared__ uint32_t warp_tot[kBitsBlock / 32];
  __shared__ uint64_t off_s;
  const size_t nwords = (n + 31) / 32;
  const size_t tiles = (n + kFilterTileRows - 1) / kFilterTileRows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t w0 = ((size_t)blockIdx.x * kBitsBlock + threadIdx.x) * kBitsWordsPerThread;
  uint32_t sel[kBitsWordsPerThread], val[kBitsWordsPerThread];
  if (vec && w0 + kBitsWordsPerThread <= nwords && (w0 + kBitsWordsPerThread) * 32 <= n) {
    // full words only: two 16-byte loads per bitmap
#pragma unroll
    for (int q = 0; q < kBitsWordsPerThread / 4; ++q) {
      uint4 m = __ldg(reinterpret_cast<const uint4*>(mask + w0) + q);
      if (vmask) {
        const uint4 x = __ldg(reinterpret_cast<const uint4*>(vmask + w0) + q);
        m.x &= x.x; m.y &= x.y; m.z &= x.z; m.w &= x.w;
      }
      const uint4 y = __ldg(reinterpret_cast<const uint4*>(vsrc + w0) + q);
      sel[4 * q] = m.x; sel[4 * q + 1] = m.y; sel[4 * q + 2] = m.z; sel[4 * q + 3] = m.w;
      val[4 * q] = y.x; val[4 * q + 1] = y.y; val[4 * q + 2] = y.z; val[4 * q + 3] = y.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kBitsWordsPerThread; ++k) {
      sel[k] = sel_word(mask, vmask, w0 + k, nwords, n);
      val[k] = w0 + k < nwords ? vsrc[w0 + k] : 0u;
    }
  }
  if (warp == 1) {  // output offset of this CTA's first row: group offset + counts of the earlier tiles of the group
    const size_t t0 = (size_t)blockIdx.x * kBitsTilesPerCta;
    const size_t gstart = t0 / kFilterGroupTiles * kFilterGroupTiles;
    uint32_t before = 0;
    if (gstart + lane < t0 && gstart + lane < tiles) before += counts[gstart + lane];
    if (gstart + 32 + lane < t0 && gstart + 32 + lane < tiles) before += counts[gstart + 32 + lane];
#pragma unroll
    for (int off = 16; off; off >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, off);
    if (lane == 0) off_s = group_offsets[t0 / kFilterGroupTiles] + before;
  }
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < kBitsWordsPerThread; ++k) c += __popc(sel[k]);
  uint32_t incl = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off) incl += x;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (int k = 0; k < kBitsBlock / 32; ++k)
    if (k < warp) base += warp_tot[k];
  if (c == 0) return;
  const uint64_t pos = off_s + base + (incl - c);   // output bit position of this thread's first kept row
  if (pos >= cap) return;
  const uint32_t lead = (uint32_t)(pos & 31);
  if (pos + c <= cap) compact_words<false>(vout, pos >> 5, lead, sel, val, cap);  // the usual case: no capacity checks per word
  else compact_words<true>(vout, pos >> 5, lead, sel, val, cap);
}

// --------------------------------------------------------------------------------------------
// TMA-staged variant of the scatter pass (opt-in with AGPU_FILTER_TMA=1; measured slower, see
// run_filter).
// Persistent CTAs walk tiles round-robin.  The 16 KiB of rows of the NEXT tile are fetched by
// one bulk asynchronous copy (cp.async.bulk global -> shared, completion on an mbarrier) while
// the CTA prefix-sums, compacts and streams out the CURRENT tile, so HBM reads never pause for
// the per-tile barriers.  Shared memory: 2 raw row buffers + 1 compaction stage.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <typename U>
struct FilterTmaSmem {
  static constexpr int G = 16 / sizeof(U);
  alignas(128) U raw[2][kFilterTileRows];
  alignas(16) U stage[kFilterTileRows + G];
  uint32_t sel[kFilterTileWords];
  uint32_t pre[kFilterTileWords];
  uint32_t vstage[kFilterTileWords + 1];
  alignas(8) uint64_t bar[2];
  uint64_t off;
  uint32_t count;
};

template <typename U, bool HAS_V>
__global__ void __launch_bounds__(kBlock) filter_scatter_tma_kernel(const U* __restrict__ src,
                                                                    const uint32_t* __restrict__ vsrc,
                                                                    const uint32_t* __restrict__ mask,
                                                                    const uint32_t* __restrict__ vmask, const size_t n,
                                                                    const uint32_t* __restrict__ counts,
                                                                    const uint64_t* __restrict__ group_offsets,
                                                                    U* __restrict__ out, uint32_t* vout, const uint64_t cap) {
  constexpr int G = 16 / sizeof(U);
  constexpr int GPT = kFilterTileRows / G / kBlock;
  constexpr uint32_t kTileBytes = kFilterTileRows * sizeof(U);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  FilterTmaSmem<U>& sm = *reinterpret_cast<FilterTmaSmem<U>*>(smem_raw);

  const size_t nwords = (n + 31) / 32;
  const size_t tiles = (n + kFilterTileRows - 1) / kFilterTileRows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool vec_out = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;

  if (threadIdx.x == 0) {
    mbar_init(&sm.bar[0], 1);
    mbar_init(&sm.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  size_t tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < tiles && (tile + 1) * kFilterTileRows <= n) {
    mbar_expect_tx(&sm.bar[0], kTileBytes);
    tma_load_1d(sm.raw[0], src + tile * kFilterTileRows, kTileBytes, &sm.bar[0]);
  }
  uint32_t parity0 = 0, parity1 = 0;
  // per-tile metadata (selection word, counts of the earlier tiles of the group, group offset) is
  // fetched one tile ahead into registers, like the rows, so no global-memory latency is exposed
  // between the barriers of a tile
  auto meta_sel = [&](size_t t) -> uint32_t {
    return (threadIdx.x < kFilterTileWords && t < tiles) ? sel_word(mask, vmask, t * kFilterTileWords + threadIdx.x, nwords, n) : 0u;
  };
  auto meta_before = [&](size_t t) -> uint32_t {
    uint32_t b = 0;
    if (warp == 1 && t < tiles) {
      const size_t gstart = t / kFilterGroupTiles * kFilterGroupTiles;
      if (gstart + lane < t) b += counts[gstart + lane];
      if (gstart + 32 + lane < t) b += counts[gstart + 32 + lane];
    }
    return b;
  };
  auto meta_goff = [&](size_t t) -> uint64_t {
    return (warp == 1 && lane == 0 && t < tiles) ? group_offsets[t / kFilterGroupTiles] : 0ull;
  };
  uint32_t cur_sel = meta_sel(tile), cur_before = meta_before(tile);
  uint64_t cur_goff = meta_goff(tile);
  for (int it = 0; tile < tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const size_t row0 = tile * kFilterTileRows;
    const bool full = row0 + kFilterTileRows <= n;
    // prefetch the rows of this CTA's next tile into the other buffer (its last readers passed
    // the barrier that ends the previous iteration)
    const size_t next = tile + gridDim.x;
    if (threadIdx.x == 0 && next < tiles && (next + 1) * kFilterTileRows <= n) {
      mbar_expect_tx(&sm.bar[buf ^ 1], kTileBytes);
      tma_load_1d(sm.raw[buf ^ 1], src + next * kFilterTileRows, kTileBytes, &sm.bar[buf ^ 1]);
    }
    const uint32_t nxt_sel = meta_sel(next), nxt_before = meta_before(next);
    const uint64_t nxt_goff = meta_goff(next);
    if (threadIdx.x < kFilterTileWords) {
      sm.sel[threadIdx.x] = cur_sel;
      if (HAS_V) sm.vstage[threadIdx.x] = 0u;
    }
    if (HAS_V && threadIdx.x == 0) sm.vstage[kFilterTileWords] = 0u;
    if (warp == 1) {
      uint32_t before = cur_before;
#pragma unroll
      for (int off = 16; off; off >>= 1) before += __shfl_xor_sync(0xFFFFFFFFu, before, off);
      if (lane == 0) sm.off = cur_goff + before;
    }
    cur_sel = nxt_sel;
    cur_before = nxt_before;
    cur_goff = nxt_goff;
    __syncthreads();
    if (warp == 0) {
      uint32_t c[4], s = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) { c[k] = __popc(sm.sel[lane * 4 + k]); s += c[k]; }
      uint32_t incl = s;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += x;
      }
      uint32_t e = incl - s;
#pragma unroll
      for (int k = 0; k < 4; ++k) { sm.pre[lane * 4 + k] = e; e += c[k]; }
      if (lane == 31) sm.count = incl;
    }
    __syncthreads();
    const uint64_t off = sm.off;
    const uint32_t count_all = sm.count;
    const uint32_t count = off >= cap ? 0u : (uint32_t)min((uint64_t)count_all, cap - off);
    const uint32_t lead = vec_out ? (uint32_t)(off % G) : 0u;
    if (full) {  // every thread observes the completion of this buffer's bulk copy
      mbar_wait(&sm.bar[buf], buf ? parity1 : parity0);
      if (buf) parity1 ^= 1; else parity0 ^= 1;
    }
    if (count_all) {
      const U* rows = sm.raw[buf];
#pragma unroll
      for (int j = 0; j < GPT; ++j) {
        const int r = (j * kBlock + threadIdx.x) * G;
        const uint32_t sw = sm.sel[r >> 5];
        const uint32_t bits = (sw >> (r & 31)) & ((1u << G) - 1u);
        if (bits == 0) continue;
        uint32_t pos = lead + sm.pre[r >> 5] + __popc(sw & ((1u << (r & 31)) - 1u));
        uint32_t vw = 0;
        if (HAS_V) vw = vsrc[(row0 + r) >> 5] >> (r & 31);
        Vec<U, G> v;
        if (full) v = *reinterpret_cast<const Vec<U, G>*>(rows + r);
#pragma unroll
        for (int k = 0; k < G; ++k) {
          if ((bits >> k) & 1u) {
            sm.stage[pos] = full ? v.e[k] : src[row0 + r + k];
            if (HAS_V && ((vw >> k) & 1u) && pos - lead < count) atomicOr(&sm.vstage[(pos - lead) >> 5], 1u << ((pos - lead) & 31));
            ++pos;
          }
        }
      }
    }
    __syncthreads();
    if (count) {
      if (vec_out) {
        U* gbase = out + (off - lead);
        const uint32_t end = lead + count;
        const uint32_t nvec = (end + G - 1) / G;
        for (uint32_t q = threadIdx.x; q < nvec; q += kBlock) {
          const uint32_t e0 = q * G;
          if (e0 >= lead && e0 + G <= end) {
            Vec<U, G> t = *reinterpret_cast<const Vec<U, G>*>(sm.stage + e0);
            st_vec<U, G>(gbase, q, t);
          } else {
            for (uint32_t k = 0; k < (uint32_t)G; ++k)
              if (e0 + k >= lead && e0 + k < end) gbase[e0 + k] = sm.stage[e0 + k];
          }
        }
      } else {
        for (uint32_t i = threadIdx.x; i < count; i += kBlock) out[off + i] = sm.stage[i];
      }
      if (HAS_V) {
        const uint32_t lw_n = (count + 31) / 32;
        const uint32_t s = (uint32_t)(off & 31);
        if (threadIdx.x < lw_n) {
          const uint32_t val = sm.vstage[threadIdx.x];
          const uint64_t gw = (off >> 5) + threadIdx.x;
          if (val << s) atomicOr(vout + gw, val << s);
          if (s && (val >> (32 - s))) atomicOr(vout + gw + 1, val >> (32 - s));
        }
      }
    }
    __syncthreads();
  }
}

template <typename U, bool HAS_V>
int launch_filter_tma(agpu_device* dev, const U* src, const uint32_t* vsrc, const uint32_t* mask, const uint32_t* vmask,
                      size_t n, const FilterScratch& sc, U* out, uint32_t* vout, uint64_t cap) {
  const size_t tiles = filter_tiles(n);
  const int smem = (int)sizeof(FilterTmaSmem<U>) + 128;
  static bool configured = false;  // per instantiation
  if (!configured) {
    AGPU_CUDA(cudaFuncSetAttribute(filter_scatter_tma_kernel<U, HAS_V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int per_sm = 0;
  AGPU_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, filter_scatter_tma_kernel<U, HAS_V>, kBlock, smem));
  if (per_sm < 1) per_sm = 1;
  size_t grid = (size_t)dev->sm_count * per_sm;
  if (grid > tiles) grid = tiles;
  AGPU_LAUNCH(dev, (filter_scatter_tma_kernel<U, HAS_V>), (unsigned)grid, kBlock, smem, src, vsrc, mask, vmask, n, sc.counts,
              sc.group_offsets, out, vout, cap);
  return agpu_finish_launch();
}

template <typename U>
int run_filter(agpu_device* dev, const void* src, const uint32_t* vsrc, const uint32_t* mask, const uint32_t* vmask,
               size_t n, const FilterScratch& sc, void* out, uint32_t* vout, size_t cap) {
  const size_t tiles = filter_tiles(n);
  if (tiles > 0x7FFFFFFFull) return AGPU_EINVAL;
  if (!aligned16(src)) return AGPU_EINVAL;  // tile bases must be 16-byte aligned
  constexpr int BLOCK = 256;  // measured: 128-thread CTAs (more resident tiles) are 2-6 % slower
  const size_t super_tiles = ceil_div(n, (size_t)kFilterTileRows * (4 / sizeof(U)));
  // The TMA-staged persistent variant is kept for A/B profiling only: with 2 x 16 KiB row buffers
  // + a 16 KiB stage only 4 CTAs fit per SM and it measured 4.9 ms vs 3.4 ms for this kernel's
  // 8 independent CTAs per SM on 4 G rows at 10 % selectivity (profiles/r01_filter_variants.md).
  static const bool use_tma = getenv("AGPU_FILTER_TMA") != nullptr;
  if (use_tma) {
    if (vsrc && vout) {
      AGPU_CUDA(cudaMemsetAsync(vout, 0, ((cap + 31) / 32) * 4, dev->stream));
      return launch_filter_tma<U, true>(dev, (const U*)src, vsrc, mask, vmask, n, sc, (U*)out, vout, cap);
    }
    return launch_filter_tma<U, false>(dev, (const U*)src, vsrc, mask, vmask, n, sc, (U*)out, vout, cap);
  }
  // 512- and 1024-thread CTAs measured 10-50 % slower (profiles/r01_filter_variants.md)
  static const bool fused_validity = getenv("AGPU_FILTER_FUSED_VALIDITY") != nullptr;  // A/B: the round-1 single kernel
  if (vsrc && vout) {
    // only the words the compacted rows can reach are cleared (the kernels OR bit groups into them)
    AGPU_CUDA(cudaMemsetAsync(vout, 0, ((cap + 31) / 32) * 4, dev->stream));
    if (fused_validity) {
      AGPU_LAUNCH(dev, (filter_scatter_kernel<U, true, BLOCK>), (unsigned)super_tiles, BLOCK, 0, (const U*)src, vsrc, mask,
                  vmask, n, sc.counts, sc.group_offsets, (U*)out, vout, (uint64_t)cap);
      return agpu_finish_launch();
    }
    const size_t bit_ctas = ceil_div(ceil_div(n, (size_t)32), (size_t)kBitsCtaWords);
    const int vec = aligned16(vsrc) && aligned16(mask) && (!vmask || aligned16(vmask));
    AGPU_LAUNCH(dev, filter_bits_kernel, (unsigned)bit_ctas, kBitsBlock, 0, vsrc, mask, vmask, n, sc.counts,
                sc.group_offsets, vout, (uint64_t)cap, vec);
  }
  // (1- and 2-byte rows: a word-level scatter — 64 contiguous bytes per thread, table-driven prmt
  // compaction of two packed words at a time, whole-word stores — was built and measured SLOWER than
  // this per-row kernel: commit 1c67e8c, profiles/r02_filter_validity.md.)
  AGPU_LAUNCH(dev, (filter_scatter_kernel<U, false, BLOCK>), (unsigned)super_tiles, BLOCK, 0, (const U*)src, vsrc, mask,
              vmask, n, sc.counts, sc.group_offsets, (U*)out, vout, (uint64_t)cap);
  return agpu_finish_launch();
}

}  // namespace

extern "C" int agpu_merge(agpu_device* dev, int dtype, const void* a, const void* b, const uint32_t* mask,
                          void* out, size_t n, const uint32_t* va, const uint32_t* vb, const uint32_t* vmask,
                          uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n && (!a || !b || !mask || !out)) return AGPU_EINVAL;
  if (vout && !va && !vb && !vmask) return AGPU_EINVAL;
  if (n == 0) return 0;
  BmMerge bm{va, vb, mask, vmask, vout, 0, (n - 1) >> 5, (n & 31) ? ((1u << (n & 31)) - 1u) : 0xFFFFFFFFu};
  if (dtype == AGPU_BOOL) {
    const size_t nwords = (n + 31) / 32;
    BoolMergeOp op{(const uint32_t*)a, (const uint32_t*)b, mask, (uint32_t*)out, bm, nwords - 1,
                   (n & 31) ? ((1u << (n & 31)) - 1u) : 0xFFFFFFFFu};
    AGPU_LAUNCH(dev, bool_merge_kernel, (unsigned)ceil_div(nwords, (size_t)kBlock), kBlock, 0, op, nwords);
    return agpu_finish_launch();
  }
  switch (agpu_dtype_size(dtype)) {
    case 4: return run_merge<uint32_t>(dev, a, b, mask, out, n, bm);
    case 2: return run_merge<uint16_t>(dev, a, b, mask, out, n, bm);
    case 1: return run_merge<uint8_t>(dev, a, b, mask, out, n, bm);
    default: return AGPU_EUNSUPPORTED;
  }
}

extern "C" int agpu_take(agpu_device* dev, int dtype, const void* src, size_t src_len, const uint32_t* idx,
                         void* out, size_t m, const uint32_t* vsrc, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (m && (!src || !idx || !out)) return AGPU_EINVAL;
  if (vout && !vsrc) return AGPU_EINVAL;
  if (m == 0) return 0;
  BmAnd none{};
  if (dtype == AGPU_BOOL) {
    int rc = launch_bits(dev, TakeBitsOp{(const uint32_t*)src, src_len, idx}, (uint32_t*)out, m, none, aligned16(idx));
    if (rc || !(vsrc && vout)) return rc;
    return launch_bits(dev, TakeBitsOp{vsrc, src_len, idx}, vout, m, none, aligned16(idx));
  }
  switch (agpu_dtype_size(dtype)) {
    case 4: return run_take<uint32_t>(dev, src, src_len, idx, out, m, vsrc, vout);
    case 2: return run_take<uint16_t>(dev, src, src_len, idx, out, m, vsrc, vout);
    case 1: return run_take<uint8_t>(dev, src, src_len, idx, out, m, vsrc, vout);
    default: return AGPU_EUNSUPPORTED;
  }
}

extern "C" int agpu_put(agpu_device* dev, int dtype, const void* src, size_t src_len, const uint32_t* src_idx, void* dst,
                        size_t dst_len, const uint32_t* dst_idx, size_t m) {
  if (!dev) return AGPU_ENODEVICE;
  if (m && (!src || !src_idx || !dst || !dst_idx)) return AGPU_EINVAL;
  if (m == 0) return 0;
  const unsigned grid = (unsigned)ceil_div(m, (size_t)kBlock);
  if (dtype == AGPU_BOOL) {
    AGPU_LAUNCH(dev, put_bits_kernel, grid, kBlock, 0, (const uint32_t*)src, src_len, src_idx, (uint32_t*)dst, dst_len,
                dst_idx, m);
    return agpu_finish_launch();
  }
  BmAnd none{};
  const bool al = aligned16(src_idx) && aligned16(dst_idx);
  switch (agpu_dtype_size(dtype)) {
    case 4: return launch_ew(dev, PutOp<uint32_t>{(const uint32_t*)src, src_len, src_idx, (uint32_t*)dst, dst_len, dst_idx}, m, none, al);
    case 2: return launch_ew(dev, PutOp<uint16_t>{(const uint16_t*)src, src_len, src_idx, (uint16_t*)dst, dst_len, dst_idx}, m, none, al);
    case 1: return launch_ew(dev, PutOp<uint8_t>{(const uint8_t*)src, src_len, src_idx, (uint8_t*)dst, dst_len, dst_idx}, m, none, al);
    default: return AGPU_EUNSUPPORTED;
  }
}

extern "C" size_t agpu_filter_scratch_bytes(size_t n) {
  return ((filter_groups(n) * 8 + 15) / 16) * 16 + ((filter_tiles(n) * 4 + 15) / 16) * 16 + 16;
}

static int filter_count_impl(agpu_device* dev, const uint32_t* mask, const uint32_t* vmask, size_t n, void* scratch,
                             uint64_t* total_dev, const ExchangePost& post) {
  if (!dev) return AGPU_ENODEVICE;
  if (!scratch || !total_dev || (n && !mask)) return AGPU_EINVAL;
  if (!dev->ticket) return AGPU_ENODEVICE;
  // n == 0 still runs the kernel when there is something to post (one empty group: total = 0)
  if (n == 0 && !post.world) {
    AGPU_CUDA(cudaMemsetAsync(total_dev, 0, 8, dev->stream));
    return 0;
  }
  const FilterScratch sc = filter_scratch(scratch, n);
  const size_t groups = n ? filter_groups(n) : 1;
  if (groups > 0x7FFFFFFFull) return AGPU_EINVAL;
  const int vec = aligned16(mask) && (!vmask || aligned16(vmask));
  AGPU_LAUNCH(dev, filter_count_kernel, (unsigned)groups, kBlock, 0, mask, vmask, n, sc.counts, sc.group_offsets, vec, dev->ticket,
              (unsigned long long*)total_dev, post);
  return agpu_finish_launch();
}

extern "C" int agpu_filter_count(agpu_device* dev, const uint32_t* mask, const uint32_t* vmask, size_t n,
                                 void* scratch, uint64_t* total_dev) {
  return filter_count_impl(dev, mask, vmask, n, scratch, total_dev, ExchangePost{});
}

extern "C" int agpu_filter_count_post(agpu_device* dev, const uint32_t* mask, const uint32_t* vmask, size_t n,
                                      void* scratch, uint64_t* total_dev, void* const* peer_slots, int rank, int world,
                                      uint32_t seq) {
  ExchangePost post{};
  const int rc = agpu_make_exchange_post(peer_slots, rank, world, seq, &post);
  if (rc) return rc;
  return filter_count_impl(dev, mask, vmask, n, scratch, total_dev, post);
}

extern "C" int agpu_filter_scatter(agpu_device* dev, int dtype, const void* src, const uint32_t* vsrc,
                                   const uint32_t* mask, const uint32_t* vmask, size_t n, void* scratch,
                                   void* out, uint32_t* vout, size_t out_capacity) {
  if (!dev) return AGPU_ENODEVICE;
  if (n && (!src || !mask || !scratch)) return AGPU_EINVAL;
  if (n == 0 || out_capacity == 0) return 0;
  if (!out) return AGPU_EINVAL;
  const FilterScratch sc = filter_scratch(scratch, n);
  switch (agpu_dtype_size(dtype)) {
    case 4: return run_filter<uint32_t>(dev, src, vsrc, mask, vmask, n, sc, out, vout, out_capacity);
    case 2: return run_filter<uint16_t>(dev, src, vsrc, mask, vmask, n, sc, out, vout, out_capacity);
    case 1: return run_filter<uint8_t>(dev, src, vsrc, mask, vmask, n, sc, out, vout, out_capacity);
    default: return AGPU_EUNSUPPORTED;
  }
}

extern "C" int agpu_take_sharded(agpu_device* dev, int dtype, int n_shards, const void* const* shard_values,
                                 const uint32_t* const* shard_validity, const uint64_t* shard_begin,
                                 const uint32_t* idx, void* out, size_t m, uint32_t* vout) {
  if (!dev) return AGPU_ENODEVICE;
  if (n_shards < 1 || n_shards > AGPU_MAX_SHARDS || !shard_values || !shard_begin) return AGPU_EINVAL;
  if (m && (!idx || !out)) return AGPU_EINVAL;
  if (m == 0) return 0;
  ShardTable t{};
  t.n = n_shards;
  t.has_validity = 0;
  for (int s = 0; s < n_shards; ++s) {
    t.values[s] = shard_values[s];
    t.validity[s] = shard_validity ? shard_validity[s] : nullptr;
    if (t.validity[s]) t.has_validity = 1;
    t.begin[s] = shard_begin[s];
    if (shard_begin[s + 1] > shard_begin[s] && !shard_values[s]) return AGPU_EINVAL;
  }
  t.begin[n_shards] = shard_begin[n_shards];
  if (vout && !t.has_validity) return AGPU_EINVAL;
  switch (agpu_dtype_size(dtype)) {
    case 4: return run_take_sharded<uint32_t>(dev, t, idx, out, m, vout);
    case 2: return run_take_sharded<uint16_t>(dev, t, idx, out, m, vout);
    case 1: return run_take_sharded<uint8_t>(dev, t, idx, out, m, vout);
    default: return AGPU_EUNSUPPORTED;
  }
}
