// device.cu — device handle, stream-ordered buffers, copies, events (replaces the reference's
// GpuDevice: crates/array/src/gpu_utils/gpu_device.rs).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

extern "C" int agpu_abi_version(void) { return AGPU_ABI_VERSION; }

// ---------------------------------------------------------------------------------------------
// allocator state shared by every handle of the process
// ---------------------------------------------------------------------------------------------
struct Block {
  size_t size;
  agpu_device* owner;                // the handle whose cache the block returns to (nullptr: owner destroyed)
  bool cached;                       // freed: sitting in owner->free_blocks or in owner->pending
  bool from_malloc;                  // cudaMalloc (allocated during a stream capture) instead of the pool
  std::vector<agpu_device*> users;   // other handles that enqueued work on it (agpu_buffer_record_use)
};
static std::vector<cudaEvent_t> g_event_pool;    // recycled "position of a stream" events (guarded by g_mem_mu)
static std::mutex g_mem_mu;                      // guards g_blocks and every handle's free_blocks
static std::unordered_map<void*, Block> g_blocks;  // every block handed out or cached
static std::mutex g_registry_mu;
static std::vector<agpu_device*> g_registry;     // every live handle

static void release_cache(agpu_device* dev);     // caller holds g_mem_mu

extern "C" int agpu_device_count(int* out) {
  if (!out) return AGPU_EINVAL;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *out = 0;
    cudaGetLastError();
    return (int)e;
  }
  *out = n;
  return 0;
}

extern "C" int agpu_device_create(int ordinal, agpu_device** out) {
  if (!out) return AGPU_EINVAL;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return AGPU_ENODEVICE;
  }
  if (ordinal < 0 || ordinal >= n) return AGPU_ENODEVICE;
  AGPU_CUDA(cudaSetDevice(ordinal));
  agpu_device* d = new agpu_device();
  d->ordinal = ordinal;
  d->launches = 0;
  cudaError_t e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete d; return (int)e; }
  e = cudaEventCreateWithFlags(&d->order_event, cudaEventDisableTiming);
  if (e != cudaSuccess) { cudaStreamDestroy(d->stream); delete d; return (int)e; }
  e = cudaDeviceGetDefaultMemPool(&d->pool, ordinal);
  if (e != cudaSuccess) { cudaEventDestroy(d->order_event); cudaStreamDestroy(d->stream); delete d; return (int)e; }
  // keep freed blocks in the pool: every op allocates a fresh output (like the reference) and
  // the allocation must not cost a cudaMalloc each time
  unsigned long long threshold = ~0ull;
  cudaMemPoolSetAttribute(d->pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, ordinal);
  // one zeroed device word per handle: the "last CTA done" ticket of filter_count (routines.cu)
  if (cudaMalloc((void**)&d->ticket, 256) == cudaSuccess) cudaMemset(d->ticket, 0, 256);
  else d->ticket = nullptr;
  // experiment knobs: L2 -> DRAM fetch granularity hint in bytes (32/64/128), and PDL off
  if (const char* g = getenv("AGPU_L2_FETCH_GRANULARITY")) {
    const size_t v = (size_t)atoi(g);
    if (v) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, v);
  }
  if (const char* g = getenv("AGPU_PDL")) d->pdl = atoi(g) ? 1 : 0;
  {
    std::lock_guard<std::mutex> lock(g_registry_mu);
    g_registry.push_back(d);
  }
  *out = d;
  return 0;
}

extern "C" int agpu_device_destroy(agpu_device* dev) {
  if (!dev) return AGPU_EINVAL;
  {
    std::lock_guard<std::mutex> lock(g_registry_mu);
    for (size_t k = 0; k < g_registry.size(); ++k)
      if (g_registry[k] == dev) { g_registry.erase(g_registry.begin() + k); break; }
  }
  cudaSetDevice(dev->ordinal);
  {
    std::lock_guard<std::mutex> lock(g_mem_mu);
    release_cache(dev);
    for (auto& kv : g_blocks) {  // blocks still alive outlive their handle: freed by whoever drops them
      Block& b = kv.second;
      if (b.owner == dev) b.owner = nullptr;
      for (size_t k = 0; k < b.users.size();)
        if (b.users[k] == dev) b.users.erase(b.users.begin() + k); else ++k;
    }
  }
  cudaStreamSynchronize(dev->stream);
  if (dev->ticket) cudaFree(dev->ticket);
  cudaEventDestroy(dev->order_event);
  cudaStreamDestroy(dev->stream);
  delete dev;
  return 0;
}

extern "C" void* agpu_device_stream(agpu_device* dev) { return dev ? (void*)dev->stream : nullptr; }
extern "C" int agpu_device_ordinal(agpu_device* dev) { return dev ? dev->ordinal : -1; }
extern "C" uint64_t agpu_launch_count(agpu_device* dev) { return dev ? dev->launches : 0; }

extern "C" const char* agpu_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case AGPU_EUNSUPPORTED: return "operation not supported for this dtype";
    case AGPU_EINVAL: return "invalid argument";
    case AGPU_ENODEVICE: return "no CUDA device";
    case AGPU_EDOUBLEFREE: return "buffer freed twice or not allocated by agpu_alloc";
    case AGPU_ETIMEOUT: return "a peer GPU did not post its value in time";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown agpu error";
}

static size_t round_block(size_t bytes) {
  if (bytes == 0) bytes = 16;  // keep a distinct non-NULL pointer for empty columns
  const size_t gran = bytes < (1u << 20) ? 512 : (bytes < (64u << 20) ? (1u << 20) : (16u << 20));
  return (bytes + gran - 1) / gran * gran;
}

static void release_block(agpu_device* dev, void* ptr, const Block& b) {
  if (b.from_malloc) cudaFree(ptr);
  else cudaFreeAsync(ptr, dev->stream);
}

static void process_pending(agpu_device* dev, bool wait);

static void release_cache(agpu_device* dev) {  // caller holds g_mem_mu
  process_pending(dev, true);
  for (auto& kv : dev->free_blocks) {
    auto it = g_blocks.find(kv.second);
    if (it != g_blocks.end()) {
      release_block(dev, kv.second, it->second);
      g_blocks.erase(it);
    }
  }
  dev->free_blocks.clear();
  dev->cached_bytes = 0;
}

// cudaMalloc, cudaHostAlloc, cudaEventDestroy ... are "potentially unsafe" calls while this thread
// captures a stream: one of them invalidates the capture (error 901 on the next launch) unless the
// thread is in relaxed mode, so the thread's capture mode is switched around them.  They can come
// at any moment — a garbage-collected host object dropping an event in the middle of a recording.
struct RelaxedCaptureMode {
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  RelaxedCaptureMode() { cudaThreadExchangeStreamCaptureMode(&mode); }
  ~RelaxedCaptureMode() { cudaThreadExchangeStreamCaptureMode(&mode); }
};

static cudaError_t malloc_during_capture(void** out, size_t bytes) {
  RelaxedCaptureMode relaxed;
  return cudaMalloc(out, bytes);
}

static bool any_handle_capturing() {
  std::lock_guard<std::mutex> reg(g_registry_mu);
  for (agpu_device* d : g_registry)
    if (d->capturing) return true;
  return false;
}

extern "C" int agpu_alloc(agpu_device* dev, size_t bytes, void** out) {
  if (!dev) return AGPU_ENODEVICE;
  if (!out) return AGPU_EINVAL;
  *out = nullptr;
  const size_t want = round_block(bytes);
  std::lock_guard<std::mutex> lock(g_mem_mu);
  process_pending(dev, false);
  // best fit among cached blocks, but never waste more than 25 % (+1 MiB) of a block
  auto it = dev->free_blocks.lower_bound(want);
  if (it != dev->free_blocks.end() && it->first <= want + want / 4 + (1u << 20)) {
    *out = it->second;
    dev->cached_bytes -= it->first;
    dev->free_blocks.erase(it);
    g_blocks[*out].cached = false;
    return 0;
  }
  agpu_make_current(dev);
  cudaError_t e = dev->capturing ? malloc_during_capture(out, want) : cudaMallocAsync(out, want, dev->stream);
  if (e == cudaErrorMemoryAllocation && !dev->capturing) {  // give the caches back to the driver and retry once
    cudaGetLastError();
    release_cache(dev);
    {
      std::lock_guard<std::mutex> reg(g_registry_mu);
      for (agpu_device* other : g_registry)
        if (other != dev && other->ordinal == dev->ordinal && !other->capturing) release_cache(other);
    }
    cudaDeviceSynchronize();  // the stream-ordered frees of every handle have to complete first
    e = cudaMallocAsync(out, want, dev->stream);
  }
  if (e != cudaSuccess) return (int)e;
  g_blocks[*out] = Block{want, dev, false, dev->capturing, {}};
  return 0;
}

static cudaEvent_t take_event() {  // caller holds g_mem_mu
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  return e;
}

// Blocks that other handles were still using when they were freed sit in owner->pending, each with
// one event per such handle (its stream position at the time of the free).  They enter the cache —
// and can be handed out again — only once all those events have completed.  Nothing waits: neither
// the host nor any stream (same idea as record_stream in torch's caching allocator).
static void process_pending(agpu_device* dev, bool wait) {  // caller holds g_mem_mu
  if (dev->pending.empty() || dev->capturing) return;
  size_t keep = 0;
  for (size_t k = 0; k < dev->pending.size(); ++k) {
    PendingBlock& p = dev->pending[k];
    bool done = true;
    for (cudaEvent_t e : p.events) {
      if (wait) cudaEventSynchronize(e);
      else if (cudaEventQuery(e) != cudaSuccess) { done = false; break; }
    }
    if (!done) {
      cudaGetLastError();  // cudaErrorNotReady is sticky-free but shows up in cudaPeekAtLastError
      if (keep != k) dev->pending[keep] = std::move(p);
      ++keep;
      continue;
    }
    for (cudaEvent_t e : p.events) g_event_pool.push_back(e);
    auto it = g_blocks.find(p.ptr);
    if (it != g_blocks.end()) {
      dev->free_blocks.emplace(it->second.size, p.ptr);
      dev->cached_bytes += it->second.size;
    }
  }
  dev->pending.resize(keep);
}

extern "C" int agpu_free(agpu_device* dev, void* ptr) {
  if (!dev) return AGPU_ENODEVICE;
  if (!ptr) return 0;
  std::lock_guard<std::mutex> lock(g_mem_mu);
  auto it = g_blocks.find(ptr);
  if (it == g_blocks.end() || it->second.cached) return AGPU_EDOUBLEFREE;
  Block& b = it->second;
  if (!b.owner) {  // the allocating handle is gone: hand the block back to the driver once every user is done
    for (agpu_device* u : b.users) cudaStreamSynchronize(u->stream);
    agpu_make_current(dev);
    if (b.from_malloc) { cudaStreamSynchronize(dev->stream); cudaFree(ptr); }
    else cudaFreeAsync(ptr, dev->stream);
    g_blocks.erase(it);
    return 0;
  }
  agpu_device* owner = b.owner;
  if (owner->capturing) {  // a temporary of the graph being captured: it stays with the graph
    b.users.clear();
    owner->capture_freed.push_back(ptr);
    return 0;
  }
  // the block may only be handed out again (on the owner's stream) after every OTHER handle that
  // read or wrote it has got past the work it had enqueued: the freeing handle and the recorded users
  if (dev != owner) {
    bool known = false;
    for (agpu_device* u : b.users) known = known || u == dev;
    if (!known) b.users.push_back(dev);
  }
  b.cached = true;
  if (b.users.empty()) {
    owner->free_blocks.emplace(b.size, ptr);
    owner->cached_bytes += b.size;
  } else {
    PendingBlock p;
    p.ptr = ptr;
    for (agpu_device* u : b.users) {
      if (u->capturing) continue;
      agpu_make_current(u);
      cudaEvent_t e = take_event();
      if (e && cudaEventRecord(e, u->stream) == cudaSuccess) p.events.push_back(e);
    }
    b.users.clear();
    owner->pending.push_back(std::move(p));
  }
  process_pending(owner, false);
  return 0;
}

extern "C" int agpu_buffer_record_use(agpu_device* dev, const void* ptr) {
  if (!dev) return AGPU_ENODEVICE;
  if (!ptr) return 0;
  std::lock_guard<std::mutex> lock(g_mem_mu);
  auto it = g_blocks.find(const_cast<void*>(ptr));
  if (it == g_blocks.end()) return 0;  // not pool memory (IPC / foreign allocation): nothing to guard
  Block& b = it->second;
  if (b.cached) return AGPU_EDOUBLEFREE;  // use after free
  if (b.owner == dev) return 0;
  for (agpu_device* u : b.users)
    if (u == dev) return 0;
  b.users.push_back(dev);
  return 0;
}

/* give every cached block back to the driver pool (e.g. before another library needs the memory) */
extern "C" int agpu_trim(agpu_device* dev) {
  if (!dev) return AGPU_ENODEVICE;
  std::lock_guard<std::mutex> lock(g_mem_mu);
  agpu_make_current(dev);
  release_cache(dev);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// one submit per recorded pipeline: stream capture -> CUDA graph (compute_pipeline.rs:259-273)
// ---------------------------------------------------------------------------------------------
static std::mutex g_graveyard_mu;
static std::vector<agpu_graph*> g_graveyard;  // graphs dropped while some handle was recording

static int destroy_graph_now(agpu_graph* g);

extern "C" int agpu_graph_begin(agpu_device* dev) {
  if (!dev) return AGPU_ENODEVICE;
  if (dev->capturing) return AGPU_EINVAL;
  agpu_make_current(dev);
  AGPU_CUDA(cudaStreamBeginCapture(dev->stream, cudaStreamCaptureModeThreadLocal));
  std::lock_guard<std::mutex> lock(g_mem_mu);
  dev->capturing = true;
  dev->capture_launches0 = dev->launches;
  dev->capture_freed.clear();
  return 0;
}

static void return_blocks_to_cache(agpu_device* dev, std::vector<void*>& blocks) {  // caller holds g_mem_mu
  for (void* p : blocks) {
    auto it = g_blocks.find(p);
    if (it == g_blocks.end()) continue;
    Block& b = it->second;
    agpu_device* owner = b.owner ? b.owner : dev;
    b.owner = owner;
    b.cached = true;
    owner->free_blocks.emplace(b.size, p);
    owner->cached_bytes += b.size;
  }
  blocks.clear();
}

extern "C" int agpu_graph_end(agpu_device* dev, agpu_graph** out) {
  if (!dev) return AGPU_ENODEVICE;
  if (!dev->capturing || !out) return AGPU_EINVAL;
  *out = nullptr;
  agpu_make_current(dev);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(dev->stream, &graph);
  agpu_graph* g = new agpu_graph();
  {
    std::lock_guard<std::mutex> lock(g_mem_mu);
    dev->capturing = false;
    g->blocks.swap(dev->capture_freed);
    g->kernels = dev->launches - dev->capture_launches0;
    dev->launches = dev->capture_launches0;  // captured kernels have not run yet: counted per replay
    if (e != cudaSuccess || !graph) return_blocks_to_cache(dev, g->blocks);
  }
  if (e != cudaSuccess || !graph) {
    delete g;
    cudaGetLastError();
    return e != cudaSuccess ? (int)e : AGPU_EINVAL;
  }
  g->graph = graph;
  g->dev = dev;
  const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
  if (ie != cudaSuccess) {
    cudaGraphDestroy(graph);
    std::lock_guard<std::mutex> lock(g_mem_mu);
    return_blocks_to_cache(dev, g->blocks);
    delete g;
    return (int)ie;
  }
  *out = g;
  {
    std::vector<agpu_graph*> todo;
    {
      std::lock_guard<std::mutex> lock(g_graveyard_mu);
      if (!any_handle_capturing()) todo.swap(g_graveyard);
    }
    for (agpu_graph* x : todo) destroy_graph_now(x);
  }
  return 0;
}

extern "C" int agpu_graph_launch(agpu_device* dev, agpu_graph* g) {
  if (!dev) return AGPU_ENODEVICE;
  if (!g || !g->exec || g->dev != dev || dev->capturing) return AGPU_EINVAL;
  agpu_make_current(dev);
  AGPU_CUDA(cudaGraphLaunch(g->exec, dev->stream));
  dev->launches += g->kernels;
  return 0;
}

extern "C" uint64_t agpu_graph_kernel_count(agpu_graph* g) { return g ? g->kernels : 0; }

extern "C" int agpu_graph_destroy(agpu_graph* g) {
  if (!g) return 0;
  // A host object owning an older graph is often dropped in the middle of the NEXT recording
  // (the variable is rebound after the new pipeline began capturing).  Destroying a graph there
  // would invalidate that capture: park it until no handle is recording.
  std::vector<agpu_graph*> todo;
  {
    std::lock_guard<std::mutex> lock(g_graveyard_mu);
    g_graveyard.push_back(g);
    if (!any_handle_capturing()) todo.swap(g_graveyard);
  }
  for (agpu_graph* x : todo) destroy_graph_now(x);
  return 0;
}

static int destroy_graph_now(agpu_graph* g) {
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  {
    // replays were ordered on the handle's stream, and so is every later user of these blocks
    std::lock_guard<std::mutex> lock(g_mem_mu);
    bool alive = false;
    {
      std::lock_guard<std::mutex> reg(g_registry_mu);
      for (agpu_device* d : g_registry) alive = alive || d == g->dev;
    }
    if (alive) return_blocks_to_cache(g->dev, g->blocks);
  }
  delete g;
  return 0;
}

extern "C" int agpu_h2d(agpu_device* dev, void* dst, const void* src, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  if (dev->capturing) return AGPU_EINVAL;  // would read host memory / synchronise inside a stream capture
  agpu_make_current(dev);
  if (bytes == 0) return 0;
  AGPU_REQUIRE(dst && src);
  AGPU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, dev->stream));
  return 0;
}

extern "C" int agpu_d2h(agpu_device* dev, void* dst, const void* src, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  if (dev->capturing) return AGPU_EINVAL;  // would read host memory / synchronise inside a stream capture
  agpu_make_current(dev);
  if (bytes) {
    AGPU_REQUIRE(dst && src);
    AGPU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dev->stream));
  }
  AGPU_CUDA(cudaStreamSynchronize(dev->stream));
  return 0;
}

extern "C" int agpu_d2h_async(agpu_device* dev, void* dst, const void* src, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  agpu_make_current(dev);
  if (bytes == 0) return 0;
  AGPU_REQUIRE(dst && src);
  AGPU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dev->stream));
  return 0;
}

extern "C" int agpu_d2d(agpu_device* dev, void* dst, const void* src, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  agpu_make_current(dev);
  if (bytes == 0) return 0;
  AGPU_REQUIRE(dst && src);
  AGPU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, dev->stream));
  return 0;
}

extern "C" int agpu_memset(agpu_device* dev, void* dst, int byte_value, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  agpu_make_current(dev);
  if (bytes == 0) return 0;
  AGPU_REQUIRE(dst);
  AGPU_CUDA(cudaMemsetAsync(dst, byte_value, bytes, dev->stream));
  return 0;
}

extern "C" int agpu_sync(agpu_device* dev) {
  if (!dev) return AGPU_ENODEVICE;
  if (dev->capturing) return AGPU_EINVAL;  // would read host memory / synchronise inside a stream capture
  agpu_make_current(dev);
  AGPU_CUDA(cudaStreamSynchronize(dev->stream));
  return 0;
}

extern "C" int agpu_host_alloc(size_t bytes, void** out) {
  if (!out) return AGPU_EINVAL;
  RelaxedCaptureMode relaxed;
  AGPU_CUDA(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
  return 0;
}

extern "C" int agpu_host_free(void* ptr) {
  if (!ptr) return 0;
  RelaxedCaptureMode relaxed;
  AGPU_CUDA(cudaFreeHost(ptr));
  return 0;
}

extern "C" int agpu_event_create(agpu_event** out) {
  if (!out) return AGPU_EINVAL;
  agpu_event* e = new agpu_event();
  RelaxedCaptureMode relaxed;
  cudaError_t err = cudaEventCreate(&e->ev);
  if (err != cudaSuccess) { delete e; return (int)err; }
  *out = e;
  return 0;
}

extern "C" int agpu_event_destroy(agpu_event* ev) {
  if (!ev) return 0;
  {
    RelaxedCaptureMode relaxed;
    cudaEventDestroy(ev->ev);
  }
  delete ev;
  return 0;
}

extern "C" int agpu_event_record(agpu_device* dev, agpu_event* ev) {
  if (!dev) return AGPU_ENODEVICE;
  agpu_make_current(dev);
  AGPU_REQUIRE(ev);
  AGPU_CUDA(cudaEventRecord(ev->ev, dev->stream));
  return 0;
}

extern "C" int agpu_stream_wait_event(agpu_device* dev, agpu_event* ev) {
  if (!dev) return AGPU_ENODEVICE;
  agpu_make_current(dev);
  AGPU_REQUIRE(ev);
  AGPU_CUDA(cudaStreamWaitEvent(dev->stream, ev->ev, 0));
  return 0;
}

extern "C" int agpu_event_elapsed_ms(agpu_event* start, agpu_event* stop, float* ms) {
  AGPU_REQUIRE(start && stop && ms);
  AGPU_CUDA(cudaEventSynchronize(stop->ev));
  AGPU_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
  return 0;
}

// ---- CUDA IPC: shards visible to the other GPUs of the box (peer memory over NVLink) ----
extern "C" int agpu_ipc_alloc(agpu_device* dev, size_t bytes, void** out) {
  if (!dev) return AGPU_ENODEVICE;
  AGPU_REQUIRE(out);
  AGPU_CUDA(cudaSetDevice(dev->ordinal));
  // Sizes are rounded up to 32 MiB (2 MiB below that).  Measured on B200 (profiles/
  // r01_peer_take_probe.md): a 1 999 998 976-byte shard (not a multiple of 2 MiB) mapped into a
  // peer makes random peer gathers over more than 1 GiB of it 40x slower (0.17 vs 7 G rows/s),
  // while 1984 MiB and 2048 MiB shards do not — the odd tail is evidently mapped with small
  // pages and overflows the peer translation caches.
  const size_t gran = bytes >= (32u << 20) ? (32u << 20) : (2u << 20);
  const size_t rounded = bytes ? (bytes + gran - 1) / gran * gran : gran;
  AGPU_CUDA(cudaMalloc(out, rounded));
  return 0;
}

extern "C" int agpu_ipc_free(agpu_device* dev, void* ptr) {
  if (!dev) return AGPU_ENODEVICE;
  if (!ptr) return 0;
  AGPU_CUDA(cudaStreamSynchronize(dev->stream));
  AGPU_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" int agpu_ipc_export(agpu_device* dev, const void* ptr, unsigned char handle[AGPU_IPC_HANDLE_BYTES]) {
  if (!dev) return AGPU_ENODEVICE;
  AGPU_REQUIRE(ptr && handle);
  static_assert(sizeof(cudaIpcMemHandle_t) == AGPU_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  AGPU_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle, &h, sizeof(h));
  return 0;
}

extern "C" int agpu_ipc_open(agpu_device* dev, const unsigned char handle[AGPU_IPC_HANDLE_BYTES], void** out) {
  if (!dev) return AGPU_ENODEVICE;
  AGPU_REQUIRE(handle && out);
  AGPU_CUDA(cudaSetDevice(dev->ordinal));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  AGPU_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int agpu_ipc_close(agpu_device* dev, void* ptr) {
  if (!dev) return AGPU_ENODEVICE;
  if (!ptr) return 0;
  AGPU_CUDA(cudaStreamSynchronize(dev->stream));
  AGPU_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}
