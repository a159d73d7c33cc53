// device.cu — device handle, stream-ordered buffers, copies, events (replaces the reference's
// GpuDevice: crates/array/src/gpu_utils/gpu_device.rs).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

extern "C" int agpu_abi_version(void) { return AGPU_ABI_VERSION; }

static void release_cache(agpu_device* dev);

// every live handle of this process: on an out-of-memory a handle asks the other handles of the
// same GPU (e.g. an upload stream's handle) to give their cached blocks back as well
static std::mutex g_registry_mu;
static std::vector<agpu_device*> g_registry;

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
  e = cudaDeviceGetDefaultMemPool(&d->pool, ordinal);
  if (e != cudaSuccess) { cudaStreamDestroy(d->stream); delete d; return (int)e; }
  // keep freed blocks in the pool: every op allocates a fresh output (like the reference) and
  // the allocation must not cost a cudaMalloc each time
  unsigned long long threshold = ~0ull;
  cudaMemPoolSetAttribute(d->pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, ordinal);
  // experiment knob: L2 -> DRAM fetch granularity hint in bytes (32/64/128); gathers fetch less
  // with a small value, streaming kernels request whole lines either way
  if (const char* g = getenv("AGPU_L2_FETCH_GRANULARITY")) {
    const size_t v = (size_t)atoi(g);
    if (v) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, v);
  }
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
    std::lock_guard<std::mutex> lock(dev->mu);
    release_cache(dev);
  }
  cudaStreamSynchronize(dev->stream);
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

static void release_cache(agpu_device* dev) {  // caller holds dev->mu
  for (auto& kv : dev->free_blocks) {
    cudaFreeAsync(kv.second, dev->stream);
    dev->block_size.erase(kv.second);
  }
  dev->free_blocks.clear();
  dev->cached_bytes = 0;
}

extern "C" int agpu_alloc(agpu_device* dev, size_t bytes, void** out) {
  if (!dev) return AGPU_ENODEVICE;
  if (!out) return AGPU_EINVAL;
  *out = nullptr;
  const size_t want = round_block(bytes);
  std::lock_guard<std::mutex> lock(dev->mu);
  // best fit among cached blocks, but never waste more than 25 % (+1 MiB) of a block
  auto it = dev->free_blocks.lower_bound(want);
  if (it != dev->free_blocks.end() && it->first <= want + want / 4 + (1u << 20)) {
    *out = it->second;
    dev->cached_bytes -= it->first;
    dev->free_blocks.erase(it);
    return 0;
  }
  AGPU_CUDA(cudaSetDevice(dev->ordinal));
  cudaError_t e = cudaMallocAsync(out, want, dev->stream);
  if (e == cudaErrorMemoryAllocation) {  // give the caches back to the driver and retry once
    cudaGetLastError();
    release_cache(dev);
    {
      // other handles of the same GPU: try_lock, so two handles running out of memory at the same
      // moment cannot wait on each other
      std::lock_guard<std::mutex> reg(g_registry_mu);
      for (agpu_device* other : g_registry) {
        if (other == dev || other->ordinal != dev->ordinal) continue;
        if (other->mu.try_lock()) {
          release_cache(other);
          other->mu.unlock();
        }
      }
    }
    cudaDeviceSynchronize();  // the stream-ordered frees of every handle have to complete first
    e = cudaMallocAsync(out, want, dev->stream);
  }
  if (e != cudaSuccess) return (int)e;
  dev->block_size[*out] = want;
  return 0;
}

extern "C" int agpu_free(agpu_device* dev, void* ptr) {
  if (!dev) return AGPU_ENODEVICE;
  if (!ptr) return 0;
  std::lock_guard<std::mutex> lock(dev->mu);
  auto it = dev->block_size.find(ptr);
  if (it == dev->block_size.end()) {  // not one of ours (should not happen): plain stream-ordered free
    AGPU_CUDA(cudaFreeAsync(ptr, dev->stream));
    return 0;
  }
  dev->free_blocks.emplace(it->second, ptr);
  dev->cached_bytes += it->second;
  return 0;
}

/* give every cached block back to the driver pool (e.g. before another library needs the memory) */
extern "C" int agpu_trim(agpu_device* dev) {
  if (!dev) return AGPU_ENODEVICE;
  std::lock_guard<std::mutex> lock(dev->mu);
  release_cache(dev);
  return 0;
}

extern "C" int agpu_h2d(agpu_device* dev, void* dst, const void* src, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  if (bytes == 0) return 0;
  AGPU_REQUIRE(dst && src);
  AGPU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, dev->stream));
  return 0;
}

extern "C" int agpu_d2h(agpu_device* dev, void* dst, const void* src, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  if (bytes) {
    AGPU_REQUIRE(dst && src);
    AGPU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dev->stream));
  }
  AGPU_CUDA(cudaStreamSynchronize(dev->stream));
  return 0;
}

extern "C" int agpu_d2h_async(agpu_device* dev, void* dst, const void* src, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  if (bytes == 0) return 0;
  AGPU_REQUIRE(dst && src);
  AGPU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, dev->stream));
  return 0;
}

extern "C" int agpu_d2d(agpu_device* dev, void* dst, const void* src, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  if (bytes == 0) return 0;
  AGPU_REQUIRE(dst && src);
  AGPU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, dev->stream));
  return 0;
}

extern "C" int agpu_memset(agpu_device* dev, void* dst, int byte_value, size_t bytes) {
  if (!dev) return AGPU_ENODEVICE;
  if (bytes == 0) return 0;
  AGPU_REQUIRE(dst);
  AGPU_CUDA(cudaMemsetAsync(dst, byte_value, bytes, dev->stream));
  return 0;
}

extern "C" int agpu_sync(agpu_device* dev) {
  if (!dev) return AGPU_ENODEVICE;
  AGPU_CUDA(cudaStreamSynchronize(dev->stream));
  return 0;
}

extern "C" int agpu_host_alloc(size_t bytes, void** out) {
  if (!out) return AGPU_EINVAL;
  AGPU_CUDA(cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault));
  return 0;
}

extern "C" int agpu_host_free(void* ptr) {
  if (!ptr) return 0;
  AGPU_CUDA(cudaFreeHost(ptr));
  return 0;
}

extern "C" int agpu_event_create(agpu_event** out) {
  if (!out) return AGPU_EINVAL;
  agpu_event* e = new agpu_event();
  cudaError_t err = cudaEventCreate(&e->ev);
  if (err != cudaSuccess) { delete e; return (int)err; }
  *out = e;
  return 0;
}

extern "C" int agpu_event_destroy(agpu_event* ev) {
  if (!ev) return 0;
  cudaEventDestroy(ev->ev);
  delete ev;
  return 0;
}

extern "C" int agpu_event_record(agpu_device* dev, agpu_event* ev) {
  if (!dev) return AGPU_ENODEVICE;
  AGPU_REQUIRE(ev);
  AGPU_CUDA(cudaEventRecord(ev->ev, dev->stream));
  return 0;
}

extern "C" int agpu_stream_wait_event(agpu_device* dev, agpu_event* ev) {
  if (!dev) return AGPU_ENODEVICE;
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
