// common.cuh — shared pieces of libagpu.so (sm_100a only; no CPU fallback anywhere).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/agpu.h"

struct agpu_graph;

struct PendingBlock {  // a freed block that other handles' streams have not got past yet (device.cu)
  void* ptr;
  std::vector<cudaEvent_t> events;
};

struct agpu_device {
  int ordinal;
  cudaStream_t stream;
  cudaMemPool_t pool;
  unsigned long long launches;  // kernels launched through this handle
  int sm_count;
  // Stream-ordered caching allocator in front of cudaMallocAsync (device.cu): every op allocates
  // a fresh output (like the reference), so freed blocks are kept by size and handed out again
  // without going back to the driver pool — whose remapping when block sizes alternate costs
  // milliseconds (measured: profiles/r01_size_sweep.md).  Reuse is safe because a block only
  // re-enters its OWNER's cache after every other handle that used it (agpu_buffer_record_use)
  // has got past that work: agpu_free records one event per such handle and parks the block in
  // `pending` until they have all completed.
  // All allocator state of all handles is guarded by one global mutex (device.cu: g_mem_mu).
  std::multimap<size_t, void*> free_blocks;  // size -> cached block
  size_t cached_bytes = 0;
  std::vector<PendingBlock> pending;         // freed, but still in use on another handle's stream
  cudaEvent_t order_event = nullptr;         // spare event of the handle
  // stream capture (agpu_graph_begin .. agpu_graph_end): blocks freed while capturing stay with the
  // graph (its kernels write them at every replay), they do not go back to the cache
  bool capturing = false;
  unsigned long long capture_launches0 = 0;
  std::vector<void*> capture_freed;
  int pdl = 1;                               // programmatic dependent launch for the streaming kernels
  unsigned int* ticket = nullptr;            // device word, zero between launches: "last CTA done" of filter_count
};

struct agpu_event {
  cudaEvent_t ev;
};

// a captured ArrowComputePipeline: one cudaGraphLaunch replays every kernel recorded between
// agpu_graph_begin and agpu_graph_end (compute_pipeline.rs:259-273: one submit per pipeline)
struct agpu_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  agpu_device* dev = nullptr;
  unsigned long long kernels = 0;   // kernel launches captured
  std::vector<void*> blocks;        // temporaries freed during the capture: owned until destroy
};

#define AGPU_CUDA(expr)                          \
  do {                                           \
    cudaError_t _e = (expr);                     \
    if (_e != cudaSuccess) return (int)_e;       \
  } while (0)

#define AGPU_REQUIRE(cond)             \
  do {                                 \
    if (!(cond)) return AGPU_EINVAL;   \
  } while (0)

// Work is always issued with the handle's device current: a process may hold handles of several
// ordinals, and a worker thread starts on device 0 (cudaGetDevice is a thread-local read).
static inline void agpu_make_current(const agpu_device* dev) {
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != dev->ordinal) cudaSetDevice(dev->ordinal);
}

// Every kernel launch goes through one of these macros so that agpu_launch_count() is exact.
#define AGPU_LAUNCH(dev, kernel, grid, block, smem, ...)                       \
  do {                                                                         \
    agpu_make_current(dev);                                                    \
    kernel<<<(grid), (block), (smem), (dev)->stream>>>(__VA_ARGS__);           \
    (dev)->launches++;                                                         \
  } while (0)

// Programmatic dependent launch (sm_90+): the grid may start while the previous kernel of the
// stream drains its last wave; the kernel itself executes `griddepcontrol.wait` (pdl_wait())
// before its first global-memory access, which returns once every earlier grid has completed and
// its writes are visible.  Launch ramp and tail of back-to-back streaming kernels overlap.
#define AGPU_LAUNCH_PDL(dev, kernel, grid, block, smem, ...)                                  \
  do {                                                                                        \
    agpu_make_current(dev);                                                                   \
    cudaLaunchConfig_t _cfg = {};                                                             \
    _cfg.gridDim = dim3(grid);                                                                \
    _cfg.blockDim = dim3(block);                                                              \
    _cfg.dynamicSmemBytes = (smem);                                                           \
    _cfg.stream = (dev)->stream;                                                              \
    cudaLaunchAttribute _attr[1];                                                             \
    _attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                         \
    _attr[0].val.programmaticStreamSerializationAllowed = (dev)->pdl;                         \
    _cfg.attrs = _attr;                                                                       \
    _cfg.numAttrs = 1;                                                                        \
    cudaLaunchKernelEx(&_cfg, kernel, __VA_ARGS__);                                           \
    (dev)->launches++;                                                                        \
  } while (0)

#ifdef __CUDACC__
// first statement of every kernel launched with AGPU_LAUNCH_PDL (a no-op for a normal launch)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// lets the next kernel of the stream begin launching; its own pdl_wait() still orders the data
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

static inline int agpu_finish_launch() {
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kBlock = 256;  // threads per CTA for all streaming kernels

static inline size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Streaming global memory access.  Columns are read once and written once, so loads use the
// evict-first path (ld.global.cs) and stores st.global.cs: nothing here is worth keeping in
// L1/L2.  BYTES is the per-lane chunk: 16 (LDG.128), 8 (LDG.64) or 4 (LDG.32); a warp always
// covers 32 consecutive chunks, i.e. one fully coalesced 512/256/128-byte request.
// ---------------------------------------------------------------------------------------------
template <int BYTES> struct Raw;
template <> struct Raw<16> { using type = uint4; };
template <> struct Raw<8> { using type = uint2; };
template <> struct Raw<4> { using type = unsigned int; };
template <> struct Raw<2> { using type = unsigned short; };
template <> struct Raw<1> { using type = unsigned char; };

// G elements of T, loaded/stored as one chunk
template <typename T, int G>
struct alignas(sizeof(T) * G) Vec {
  T e[G];
};

template <typename T, int G>
__device__ __forceinline__ Vec<T, G> ld_vec(const T* base, size_t granule) {
  using R = typename Raw<sizeof(T) * G>::type;
  R r = __ldcs(reinterpret_cast<const R*>(base) + granule);
  Vec<T, G> v;
  memcpy(&v, &r, sizeof(v));
  return v;
}

template <typename T, int G>
__device__ __forceinline__ void st_vec(T* base, size_t granule, const Vec<T, G>& v) {
  using R = typename Raw<sizeof(T) * G>::type;
  R r;
  memcpy(&r, &v, sizeof(v));
  __stcs(reinterpret_cast<R*>(base) + granule, r);
}

// ---------------------------------------------------------------------------------------------
// Validity bitmaps handled inside the value kernel.  A tile of the value kernel covers
// `tile_words` consecutive 32-row words of every bitmap; the first tile_words/4 threads of the
// CTA move them as 16-byte vectors (128-byte lines per warp), AND-ing up to four inputs.
// nin == 0 means "no bitmap work".  NULL inputs were removed on the host (NULL = all valid).
// ---------------------------------------------------------------------------------------------
struct BmAnd {
  const uint32_t* in[4];
  uint32_t* out;
  int nin;
  int vec;  // all pointers 16-byte aligned
  __device__ __forceinline__ void tile(size_t w0, int tile_words, size_t nwords) const;
};

__device__ __forceinline__ void bm_and_tile(const BmAnd& bm, size_t w0, int tile_words, size_t nwords) {
  if (bm.nin == 0) return;
  for (int q = threadIdx.x; q * 4 < tile_words; q += blockDim.x) {
    size_t w = w0 + (size_t)q * 4;
    if (w >= nwords) break;
    if (bm.vec && w + 4 <= nwords) {
      uint4 r = __ldcs(reinterpret_cast<const uint4*>(bm.in[0] + w));
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (k < bm.nin) {
          uint4 s = __ldcs(reinterpret_cast<const uint4*>(bm.in[k] + w));
          r.x &= s.x; r.y &= s.y; r.z &= s.z; r.w &= s.w;
        }
      __stcs(reinterpret_cast<uint4*>(bm.out + w), r);
    } else {
      for (int j = 0; j < 4 && w + j < nwords; ++j) {
        uint32_t r = bm.in[0][w + j];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (k < bm.nin) r &= bm.in[k][w + j];
        bm.out[w + j] = r;
      }
    }
  }
}

__device__ __forceinline__ void BmAnd::tile(size_t w0, int tile_words, size_t nwords) const {
  bm_and_tile(*this, w0, tile_words, nwords);
}

// host helper: collect the non-NULL validity inputs
static inline BmAnd make_bm(const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d,
                            uint32_t* out) {
  BmAnd bm{};
  bm.out = out;
  bm.nin = 0;
  if (!out) return bm;
  const uint32_t* ins[4] = {a, b, c, d};
  bool al = aligned16(out);
  for (int k = 0; k < 4; ++k)
    if (ins[k]) {
      bm.in[bm.nin++] = ins[k];
      al = al && aligned16(ins[k]);
    }
  bm.vec = al ? 1 : 0;
  return bm;
}

// ---------------------------------------------------------------------------------------------
// dtype -> C type
// ---------------------------------------------------------------------------------------------
static inline size_t agpu_dtype_size(int dtype) {
  switch (dtype) {
    case AGPU_I8: case AGPU_U8: return 1;
    case AGPU_I16: case AGPU_U16: return 2;
    case AGPU_I32: case AGPU_U32: case AGPU_F32: case AGPU_DATE32: return 4;
    default: return 0;
  }
}

// ---------------------------------------------------------------------------------------------
// count exchange between the GPUs of one box (exchange.cu; the filter's count kernel can post too)
// ---------------------------------------------------------------------------------------------
constexpr int kExchangeTagBits = 24;
constexpr int kExchangeValueBits = 64 - kExchangeTagBits;  // 40: counts below 2^40 rows per shard
constexpr unsigned long long kExchangeValueMask = (1ull << kExchangeValueBits) - 1ull;

__host__ __device__ inline unsigned long long exchange_tag(uint32_t seq) {
  return (unsigned long long)(seq % 0xFFFFFFu) + 1ull;  // never 0: a zeroed slot is "nothing posted"
}

struct ExchangePost {  // world == 0: nothing to post
  unsigned long long* slots[AGPU_MAX_SHARDS];  // slot area of every rank as mapped into this process
  int rank, world;
  uint32_t seq;
};

#ifdef __CUDACC__
// lane r < world of ONE warp stores {tag, value} into slot[rank] of rank r's area: one 64-bit word
// carries tag AND value, so a relaxed system-scope store is enough (nothing else is published)
__device__ __forceinline__ void exchange_post_lane(const ExchangePost& p, unsigned long long v, int r) {
  if (r >= p.world) return;
  if (v > kExchangeValueMask) v = kExchangeValueMask;  // cannot happen for row counts; keeps the tag intact
  const unsigned long long word = (exchange_tag(p.seq) << kExchangeValueBits) | v;
  unsigned long long* dst = p.slots[r] + (size_t)(p.seq % AGPU_EXCHANGE_RING) * p.world + p.rank;
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
}
#endif

// internal launchers shared between translation units
int agpu_launch_bitmap_and(agpu_device* dev, const BmAnd& bm, size_t n_bits);
int agpu_make_exchange_post(void* const* peer_slots, int rank, int world, uint32_t seq, ExchangePost* out);
