// reduce.cu — the "next" rows of SURVEY.md §8f: broadcast, sum, any, all.
#include "elementwise.cuh"

namespace {

// ---- broadcast: array/compute_shaders/{f32,i32,u32}/broadcast.wgsl:9-13 ----
template <typename U>
struct BroadcastOp {
  static constexpr int G = 16 / sizeof(U);
  U* out;
  U value;
  struct In {};
  __device__ __forceinline__ In load(size_t) const { return In{}; }
  __device__ __forceinline__ void run(size_t g, const In&) const {
    Vec<U, G> o;
#pragma unroll
    for (int k = 0; k < G; ++k) o.e[k] = value;
    st_vec<U, G>(out, g, o);
  }
  __device__ __forceinline__ void tail(size_t i) const { out[i] = value; }
};

// ---- sum: arithmetic/compute_shaders/{f32,i32,u32}/aggregate.wgsl:13-42 ----
// The reference reduces every 256-element workgroup with the pairwise tree
//   s = 1,2,4..128: x[2*s*k] += x[2*s*k + s]
// and repeats the pass on the partials (aggregate_kernels.rs:24-52).  f32 addition is not
// associative, so the same tree is rebuilt here: one warp owns one 256-element group; lane l
// holds elements 4l..4l+3 of each 128-element half (two coalesced 16-byte loads), sums them as
// (e0+e1)+(e2+e3) (levels s=1,2), a 5-step xor-butterfly adds lanes l and l^1, l^2 .. l^16
// (levels s=4..64; a+b == b+a bit-for-bit, so every lane ends with the group's half sum) and
// the two halves are added last (s=128).  Lanes past the end contribute 0 like the shader's
// bounds check.  A warp owns 4 consecutive groups and issues their 8 loads before the first add.
template <typename T>
__device__ __forceinline__ T butterfly(T e0, T e1, T e2, T e3) {
  T s = (e0 + e1) + (e2 + e3);  // levels s = 1, 2
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) s = s + __shfl_xor_sync(0xFFFFFFFFu, s, off);  // levels 4 .. 64
  return s;
}

// bounds-checked half group (ragged end / unaligned input)
template <typename T>
__device__ __forceinline__ T half_sum_checked(const T* __restrict__ in, size_t base, size_t len, int lane) {
  const size_t i = base + (size_t)lane * 4;
  T e0 = 0, e1 = 0, e2 = 0, e3 = 0;
  if (i < len) e0 = in[i];
  if (i + 1 < len) e1 = in[i + 1];
  if (i + 2 < len) e2 = in[i + 2];
  if (i + 3 < len) e3 = in[i + 3];
  return butterfly(e0, e1, e2, e3);
}

template <typename T, int GROUPS_PER_WARP>
__global__ void __launch_bounds__(kBlock) sum_pass_kernel(const T* __restrict__ in, const size_t len,
                                                          T* __restrict__ out, const size_t groups, const int vec) {
  const int lane = threadIdx.x & 31;
  const size_t warp = (size_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  const size_t g0 = warp * GROUPS_PER_WARP;
  if (g0 >= groups) return;
  T r[GROUPS_PER_WARP];
  if (vec && (g0 + GROUPS_PER_WARP) * 256 <= len) {
    // all 2*GROUPS_PER_WARP coalesced 16-byte loads of the warp are in flight before the first add
    Vec<T, 4> v[GROUPS_PER_WARP][2];
#pragma unroll
    for (int k = 0; k < GROUPS_PER_WARP; ++k) {
      v[k][0] = ld_vec<T, 4>(in + (g0 + k) * 256, lane);
      v[k][1] = ld_vec<T, 4>(in + (g0 + k) * 256 + 128, lane);
    }
#pragma unroll
    for (int k = 0; k < GROUPS_PER_WARP; ++k) {
      const T lo = butterfly(v[k][0].e[0], v[k][0].e[1], v[k][0].e[2], v[k][0].e[3]);
      const T hi = butterfly(v[k][1].e[0], v[k][1].e[1], v[k][1].e[2], v[k][1].e[3]);
      r[k] = lo + hi;  // level s = 128
    }
  } else {
#pragma unroll
    for (int k = 0; k < GROUPS_PER_WARP; ++k) {
      const size_t g = g0 + k;
      r[k] = 0;
      if (g < groups) r[k] = half_sum_checked<T>(in, g * 256, len, lane) + half_sum_checked<T>(in, g * 256 + 128, len, lane);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < GROUPS_PER_WARP; ++k)
      if (g0 + k < groups) out[g0 + k] = r[k];
  }
}

template <typename T>
int run_sum(agpu_device* dev, const T* a, size_t n, T* out_dev) {
  if (n == 0) {
    AGPU_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(T), dev->stream));
    return 0;
  }
  constexpr int GPW = 4;
  size_t len = n;
  const T* in = a;
  // ping-pong scratch for the partials of each pass
  const size_t cap = ceil_div(n, (size_t)256);
  T* buf[2] = {nullptr, nullptr};
  if (cap > 1) {  // scratch for the partials comes from the handle's block cache
    int rc0 = agpu_alloc(dev, cap * sizeof(T), (void**)&buf[0]);
    if (rc0) return rc0;
    rc0 = agpu_alloc(dev, ceil_div(cap, (size_t)256) * sizeof(T), (void**)&buf[1]);
    if (rc0) { agpu_free(dev, buf[0]); return rc0; }
  }
  int which = 0, rc = 0;
  while (true) {
    const size_t groups = ceil_div(len, (size_t)256);
    T* dst = groups == 1 ? out_dev : buf[which];
    const size_t warps = ceil_div(groups, (size_t)GPW);
    const size_t grid = ceil_div(warps, (size_t)(kBlock / 32));
    AGPU_LAUNCH(dev, (sum_pass_kernel<T, GPW>), (unsigned)grid, kBlock, 0, in, len, dst, groups, aligned16(in) ? 1 : 0);
    rc = agpu_finish_launch();
    if (rc || groups == 1) break;
    in = dst;
    len = groups;
    which ^= 1;
  }
  if (buf[0]) agpu_free(dev, buf[0]);
  if (buf[1]) agpu_free(dev, buf[1]);
  return rc;
}

// ---- any / all: logical/compute_shaders/u32/any.wgsl, countbitones.wgsl + Sum ----
// mode 0 (any): a word with a set bit; mode 1 (all): a word with a clear bit among the first n_bits.
// One launch: warps that found something raise `found_flag` (word 1 of the handle's ticket block,
// zero between launches), every CTA then takes a ticket and the last one writes the result and
// puts both words back to zero — no memset before and no finishing kernel after (these ops are a
// 5 us pass over a bitmap followed by a read-back: launches and the host round trip are the cost).
__global__ void __launch_bounds__(kBlock) bits_find_kernel(const uint32_t* __restrict__ bits, const size_t nwords,
                                                           const size_t n_bits, const int mode, const int vec,
                                                           unsigned int* __restrict__ ticket, uint32_t* __restrict__ result) {
  const uint32_t tail_mask = (n_bits & 31) ? ((1u << (n_bits & 31)) - 1u) : 0xFFFFFFFFu;
  const uint32_t flip = mode == 0 ? 0u : 0xFFFFFFFFu;  // all: look for a clear bit
  unsigned int* found_flag = ticket + 1;
  uint32_t found = 0;
  // body: whole 16-byte chunks except the one holding the last word, 4 chunks in flight per thread
  const size_t nvec = vec ? (nwords - 1) / 4 : 0;
  const uint4* __restrict__ v = reinterpret_cast<const uint4*>(bits);
  const size_t stride = (size_t)gridDim.x * kBlock;
  size_t q = (size_t)blockIdx.x * kBlock + threadIdx.x;
  for (; q + 3 * stride < nvec; q += 4 * stride) {
    uint4 x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = __ldcs(v + q + k * stride);
#pragma unroll
    for (int k = 0; k < 4; ++k) found |= (x[k].x ^ flip) | (x[k].y ^ flip) | (x[k].z ^ flip) | (x[k].w ^ flip);
  }
  for (; q < nvec; q += stride) {
    const uint4 x = __ldcs(v + q);
    found |= (x.x ^ flip) | (x.y ^ flip) | (x.z ^ flip) | (x.w ^ flip);
  }
  // leftover words (at most 4 with vector access, everything otherwise), the last one masked
  for (size_t w = nvec * 4 + (size_t)blockIdx.x * kBlock + threadIdx.x; w < nwords; w += stride) {
    uint32_t x = bits[w] ^ flip;
    if (w == nwords - 1) x &= tail_mask;
    found |= x;
  }
  if (__syncthreads_or(found != 0) && threadIdx.x == 0) atomicOr(found_flag, 1u);
  if (threadIdx.x == 0) {
    __threadfence();                                        // the flag is visible before the ticket
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {           // last CTA: every flag raised before its ticket is visible
      __threadfence();
      const unsigned int f = atomicExch(found_flag, 0u);
      *result = mode == 0 ? (f ? 1u : 0u) : (f ? 0u : 1u);
      *ticket = 0u;
    }
  }
}

}  // namespace

extern "C" int agpu_broadcast(agpu_device* dev, int dtype, const void* scalar_host, void* out, size_t n) {
  if (!dev) return AGPU_ENODEVICE;
  if (!scalar_host || (n && !out)) return AGPU_EINVAL;
  BmAnd none{};
  switch (agpu_dtype_size(dtype)) {
    case 4: { BroadcastOp<uint32_t> op{(uint32_t*)out, *(const uint32_t*)scalar_host}; return launch_ew(dev, op, n, none, aligned16(out)); }
    case 2: { BroadcastOp<uint16_t> op{(uint16_t*)out, *(const uint16_t*)scalar_host}; return launch_ew(dev, op, n, none, aligned16(out)); }
    case 1: { BroadcastOp<uint8_t> op{(uint8_t*)out, *(const uint8_t*)scalar_host}; return launch_ew(dev, op, n, none, aligned16(out)); }
    default: return AGPU_EUNSUPPORTED;
  }
}

extern "C" int agpu_sum(agpu_device* dev, int dtype, const void* a, size_t n, void* out_dev) {
  if (!dev) return AGPU_ENODEVICE;
  if (!out_dev || (n && !a)) return AGPU_EINVAL;
  switch (dtype) {
    case AGPU_F32: return run_sum<float>(dev, (const float*)a, n, (float*)out_dev);
    case AGPU_I32: case AGPU_U32: return run_sum<uint32_t>(dev, (const uint32_t*)a, n, (uint32_t*)out_dev);
    default: return AGPU_EUNSUPPORTED;
  }
}

static int find_bits(agpu_device* dev, const uint32_t* bits, size_t n_bits, int mode, uint32_t* result_dev) {
  if (!dev) return AGPU_ENODEVICE;
  if (!result_dev || (n_bits && !bits)) return AGPU_EINVAL;
  if (!dev->ticket) return AGPU_ENODEVICE;
  const size_t nwords = (n_bits + 31) / 32;
  size_t grid = ceil_div(nwords, (size_t)kBlock * 16);
  const size_t cap = (size_t)dev->sm_count * 32;
  if (grid > cap) grid = cap;
  if (grid == 0) grid = 1;  // n_bits == 0: one CTA that finds nothing -> any = 0, all = 1
  AGPU_LAUNCH(dev, bits_find_kernel, (unsigned)grid, kBlock, 0, bits, nwords, n_bits, mode,
              nwords && aligned16(bits) ? 1 : 0, dev->ticket, result_dev);
  return agpu_finish_launch();
}

extern "C" int agpu_any(agpu_device* dev, const uint32_t* bits, size_t n_bits, uint32_t* result_dev) {
  return find_bits(dev, bits, n_bits, 0, result_dev);
}
extern "C" int agpu_all(agpu_device* dev, const uint32_t* bits, size_t n_bits, uint32_t* result_dev) {
  return find_bits(dev, bits, n_bits, 1, result_dev);
}
