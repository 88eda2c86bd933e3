/*
 * tagset.cu — slot reclamation of the device-resident flow_tag "seen" sets (tagset.cuh).
 * The reference's lru.Cache evicts its least-recently-used entry at capacity (server/libs/lru/lru.go:44-58);
 * an open-addressing set cannot delete in place, so it is rebuilt into a spare table between batches.
 */
#include <cuda_runtime.h>

#include "tagset.cuh"

namespace dfgpu {

namespace {

__device__ __forceinline__ bool ts_live(const FtEntry& e, uint32_t now, uint32_t timeout) {
  if ((e.fp_lo | e.fp_hi) == 0 || e.state == 0) return false;
  return !((uint32_t)(e.time + timeout) < now);  // lru rule: old + timeout < time => emitted again = absent
}
__device__ __forceinline__ uint32_t ts_age_bucket(uint32_t time, uint32_t now, uint32_t timeout) {
  const uint32_t age = now > time ? now - time : 0u;
  const uint64_t b = (uint64_t)age * 64 / ((uint64_t)timeout + 1);
  return b > 63 ? 63u : (uint32_t)b;
}

__global__ void k_ts_hist(const FtEntry* t, uint32_t mask, uint32_t timeout, TagSetState* st) {
  const uint32_t now = st->now;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= mask; i += gridDim.x * blockDim.x) {
    const FtEntry e = t[i];
    if ((e.fp_lo | e.fp_hi) == 0) continue;
    if (ts_live(e, now, timeout)) atomicAdd(&st->hist[ts_age_bucket(e.time, now, timeout)], 1u);
    else atomicAdd(&st->expired, 1u);
  }
}
// keep the youngest buckets while they fit the target
__global__ void k_ts_pick(TagSetState* st, uint32_t target) {
  uint32_t cum = 0, cutoff = 64;  // bucket index: entries with bucket >= cutoff are dropped
  for (uint32_t b = 0; b < 64; b++) {
    if (cum + st->hist[b] > target && b > 0) {
      cutoff = b;
      break;
    }
    cum += st->hist[b];
  }
  uint32_t dropped = 0;
  for (uint32_t b = cutoff; b < 64; b++) dropped += st->hist[b];
  st->cutoff_age = cutoff;
  st->evicted += dropped;
  st->count = 0;
  st->rebuilds++;
  for (uint32_t b = 0; b < 64; b++) st->hist[b] = 0;
}
__global__ void k_ts_rebuild(const FtEntry* old, FtEntry* neu, uint32_t mask, uint32_t timeout, TagSetState* st) {
  const uint32_t now = st->now, cutoff = st->cutoff_age;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= mask; i += gridDim.x * blockDim.x) {
    const FtEntry e = old[i];
    if (!ts_live(e, now, timeout)) continue;
    if (ts_age_bucket(e.time, now, timeout) >= cutoff) continue;
    FtEntry* n = ts_upsert(neu, mask, U128{e.fp_lo, e.fp_hi}, st);
    if (n) {
      n->time = e.time;
      n->state = 1;
      n->first_pos = 0;
    }
  }
}

}  // namespace

int tagset_create(TagSet& t, uint64_t min_slots, uint32_t timeout) {
  uint32_t n = 1u << 12;
  while (n < min_slots && n < (1u << 28)) n <<= 1;
  if (cudaMalloc(&t.cur, (size_t)n * sizeof(FtEntry)) != cudaSuccess) return -1;
  if (cudaMalloc(&t.alt, (size_t)n * sizeof(FtEntry)) != cudaSuccess) return -1;
  if (cudaMalloc(&t.d_state, sizeof(TagSetState)) != cudaSuccess) return -1;
  if (cudaMemset(t.cur, 0, (size_t)n * sizeof(FtEntry)) != cudaSuccess) return -1;
  if (cudaMemset(t.d_state, 0, sizeof(TagSetState)) != cudaSuccess) return -1;
  t.mask = n - 1;
  t.timeout = timeout;
  t.count_seen = 0;
  return 0;
}
void tagset_destroy(TagSet& t) {
  cudaFree(t.cur);
  cudaFree(t.alt);
  cudaFree(t.d_state);
  t.cur = t.alt = nullptr;
  t.d_state = nullptr;
}
int tagset_maybe_rebuild(TagSet& t, cudaStream_t s, uint64_t* launches) {
  if ((uint64_t)t.count_seen * 2 <= t.slots()) return 0;
  cudaMemsetAsync(t.alt, 0, (size_t)t.slots() * sizeof(FtEntry), s);
  k_ts_hist<<<1024, 256, 0, s>>>(t.cur, t.mask, t.timeout, t.d_state);
  k_ts_pick<<<1, 1, 0, s>>>(t.d_state, t.slots() / 4);
  k_ts_rebuild<<<1024, 256, 0, s>>>(t.cur, t.alt, t.mask, t.timeout, t.d_state);
  FtEntry* x = t.cur;
  t.cur = t.alt;
  t.alt = x;
  t.count_seen = t.slots() / 4;  // upper bound until the next fetch reports the real count
  if (launches) *launches += 3;
  return 1;
}

}  // namespace dfgpu
