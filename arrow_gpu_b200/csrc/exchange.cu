// exchange.cu — per-shard count / offset exchange between the GPUs of one box WITHOUT the host
// (BASELINE.json north_star (4): "NCCL over NVLink used only to exchange per-shard counts and
// offsets for compaction outputs").  The payload is one u64 per rank, so the cost of a collective
// library call (launch + proxy + protocol, ~20-80 us) dwarfs the data: here every rank owns a small
// slot area in cudaMalloc'ed memory that all peers have mapped through CUDA IPC, and
//   post : one warp, lane r stores {tag, value} as ONE 64-bit word into slot[my_rank] of peer r
//          (st.relaxed.sys straight over NVLink/NVSwitch — a single word carries tag and value,
//          so no separate flag, no fence and no release/acquire pairing is needed)
//   wait : one warp, lane r polls its own slot[r] until the tag of this exchange shows up
//          (ld.relaxed.sys), then the warp prefix-sums the values into global offsets.
// Both are ordinary kernels on the handle's stream: a sharded filter enqueues
// count -> post -> scatter -> wait with no host synchronisation in between.
//
// Slot reuse: exchange number `seq` uses ring row seq % AGPU_EXCHANGE_RING.  A rank can post
// exchange k+2 only after its own wait(k+1) finished, i.e. after every peer posted k+1, i.e. after
// every peer finished wait(k) — so two rows would do; four leave slack.  Every rank must call
// post and wait once per exchange, in the same order.
#include "common.cuh"

namespace {

__global__ void exchange_post_kernel(const unsigned long long* __restrict__ value, const ExchangePost post) {
  exchange_post_lane(post, *value, threadIdx.x);
}

__global__ void exchange_wait_kernel(const unsigned long long* __restrict__ my_slots, const int world, const uint32_t seq,
                                     unsigned long long* __restrict__ out, const unsigned long long timeout_ns) {
  const int r = threadIdx.x;  // one warp
  const unsigned long long want = exchange_tag(seq);
  unsigned long long v = 0;
  int ok = 1;
  if (r < world) {
    const unsigned long long* src = my_slots + (size_t)(seq % AGPU_EXCHANGE_RING) * world + r;
    unsigned long long t0, now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      unsigned long long word;
      asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(word) : "l"(src) : "memory");
      if ((word >> kExchangeValueBits) == want) { v = word & kExchangeValueMask; break; }
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > timeout_ns) { ok = 0; break; }  // a peer died or skipped the exchange: report, never hang
      __nanosleep(64);
    }
  }
  const unsigned all_ok = __all_sync(0xFFFFFFFFu, ok);
  unsigned long long incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned long long x = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (r >= off) incl += x;
  }
  if (r < world) {
    out[r] = incl - v;                    // global offset of rank r's first output row
    out[world + 2 + r] = v;               // rank r's own count
  }
  if (r == world - 1) out[world] = incl;  // total
  if (r == 0) out[world + 1] = all_ok ? 0ull : 1ull;
}

}  // namespace

extern "C" size_t agpu_exchange_bytes(int world) {
  if (world < 1 || world > AGPU_MAX_SHARDS) return 0;
  return ((size_t)AGPU_EXCHANGE_RING * world * 8 + 255) / 256 * 256;
}

int agpu_make_exchange_post(void* const* peer_slots, int rank, int world, uint32_t seq, ExchangePost* out) {
  if (!peer_slots || world < 1 || world > AGPU_MAX_SHARDS || rank < 0 || rank >= world) return AGPU_EINVAL;
  ExchangePost p{};
  for (int r = 0; r < world; ++r) {
    if (!peer_slots[r]) return AGPU_EINVAL;
    p.slots[r] = (unsigned long long*)peer_slots[r];
  }
  p.rank = rank;
  p.world = world;
  p.seq = seq;
  *out = p;
  return 0;
}

extern "C" int agpu_exchange_post(agpu_device* dev, const uint64_t* value_dev, void* const* peer_slots, int rank,
                                  int world, uint32_t seq) {
  if (!dev) return AGPU_ENODEVICE;
  if (!value_dev) return AGPU_EINVAL;
  ExchangePost p{};
  const int rc = agpu_make_exchange_post(peer_slots, rank, world, seq, &p);
  if (rc) return rc;
  AGPU_LAUNCH(dev, exchange_post_kernel, 1, 32, 0, (const unsigned long long*)value_dev, p);
  return agpu_finish_launch();
}

extern "C" int agpu_exchange_wait(agpu_device* dev, const void* my_slots, int world, uint32_t seq, uint64_t* out_dev,
                                  uint32_t timeout_ms) {
  if (!dev) return AGPU_ENODEVICE;
  if (!my_slots || !out_dev || world < 1 || world > AGPU_MAX_SHARDS) return AGPU_EINVAL;
  const unsigned long long timeout_ns = (unsigned long long)(timeout_ms ? timeout_ms : 10000u) * 1000000ull;
  AGPU_LAUNCH(dev, exchange_wait_kernel, 1, 32, 0, (const unsigned long long*)my_slots, world, seq,
              (unsigned long long*)out_dev, timeout_ns);
  return agpu_finish_launch();
}
