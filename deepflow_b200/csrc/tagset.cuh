/*
 * tagset.cuh — the device-resident "seen" set that replaces the per-decoder lru.Cache
 * of the flow_tag writers (server/libs/lru/lru.go:44-76, flow_tag_writer.go:65-77,
 * app_service_tag_writer.go:52).  Used by the L7 flow_tag stage (l7.cu) and by the
 * flow_metrics app_service tags (metrics.cu).
 *
 * A slot is 32 B (one HBM sector): a 128-bit fingerprint of the lru key published with
 * one atom.cas.b128, the time stored by the last Add(), and an epoch-tagged "first
 * candidate" word so that of all records of one batch that find the key new or expired,
 * the FIRST in input order emits the row — the order the sequential Go loop produces.
 */
#pragma once
#include <cstdint>

#include "tables.cuh"

namespace dfgpu {

struct FtEntry {
  unsigned long long fp_lo, fp_hi;  // 128-bit fingerprint of the lru key; 0/0 = empty
  unsigned long long first_pos;     // (epoch << 48) | (2^48-1 - position): atomicMax keeps the FIRST
                                    // candidate of the NEWEST batch; 0 = none
  uint32_t time;                    // value stored by lru Add(): time of the last emission
  uint32_t state;                   // 0 = inserted but never emitted, 1 = emitted
};
static_assert(sizeof(FtEntry) == 32, "FtEntry");

struct U128 {
  unsigned long long lo, hi;
};
__device__ __forceinline__ U128 cas128(FtEntry* e, U128 cmp, U128 val) {
  U128 old;
  asm volatile(
      "{\n\t.reg .b128 c, v, o;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 v, {%4, %5};\n\t"
      "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\t"
      "mov.b128 {%0, %1}, o;\n\t}"
      : "=l"(old.lo), "=l"(old.hi)
      : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(e)
      : "memory");
  return old;
}

// Strings are hashed once (two independent 64-bit streams, 8 bytes per step); a key
// combines its string hashes with its integer members.
struct StrHash {
  uint64_t a, b;
};
// Little-endian 8-byte words of the string, the last one zero padded.  The bytes come out of aligned
// 8-byte blocks (one load per 8 bytes instead of 8 byte loads); a block is only touched when it holds
// a byte of the string.
__device__ __forceinline__ StrHash str_hash(const uint8_t* p, uint32_t n) {
  uint64_t a = 0x9E3779B97F4A7C15ull ^ n, b = 0xC2B2AE3D27D4EB4Full + (uint64_t)n * 0x165667B19E3779F9ull;
  uint64_t w = 0;
  if (n) {
    const unsigned long long* sb = reinterpret_cast<const unsigned long long*>((uintptr_t)p & ~(uintptr_t)7);
    const uint32_t sh = ((uint32_t)(uintptr_t)p & 7u) * 8u;
    unsigned long long cur = *sb;
    uint32_t i = 0;
    for (; i + 8 <= n; i += 8) {
      const bool need = sh != 0 || i + 8 < n;
      const unsigned long long nxt = need ? sb[1] : 0ull;
      w = sh ? (cur >> sh) | (nxt << (64 - sh)) : cur;
      a = (a ^ w) * 0x100000001B3ull;
      a ^= a >> 32;
      b = (b + w) * 0xFF51AFD7ED558CCDull;
      b ^= b >> 29;
      cur = nxt;
      sb++;
    }
    const uint32_t r = n - i;  // 0..7 tail bytes
    w = 0;
    if (r) {
      const bool need = sh + r * 8 > 64;  // the tail runs into the next block
      const unsigned long long nxt = need ? sb[1] : 0ull;
      w = sh ? (cur >> sh) | (nxt << (64 - sh)) : cur;
      w &= ~0ull >> (64 - r * 8);
    }
  }
  a = (a ^ w) * 0x100000001B3ull;
  a ^= a >> 32;
  b = (b + w) * 0xFF51AFD7ED558CCDull;
  b ^= b >> 29;
  return StrHash{a, b};
}
__device__ __forceinline__ U128 fp_finish(uint64_t lo, uint64_t hi) {
  U128 r{mix64(lo), mix64(hi)};
  if (r.lo == 0) r.lo = 1;  // both halves non-zero: 0 marks an empty / half-published slot
  if (r.hi == 0) r.hi = 1;
  return r;
}

// Device-resident bookkeeping of one set (survives across batches).
struct TagSetState {
  uint32_t count;       // entries inserted since the last rebuild + entries the rebuild kept
  uint32_t now;         // newest lru time seen (atomicMax)
  uint32_t unrecorded;  // emissions whose key could not be stored: the set was saturated (duplicates later)
  uint32_t evicted;     // live entries dropped oldest-first by rebuilds (beyond what expiry frees)
  uint32_t expired;     // entries dropped by rebuilds because their lru time had expired (exact: expired == absent)
  uint32_t rebuilds;
  uint32_t cutoff_age;  // scratch of a rebuild: entries older than this are dropped
  uint32_t kept;        // scratch of a rebuild
  uint32_t hist[64];    // scratch of a rebuild: live entries by age bucket
};
constexpr uint32_t TS_MAX_PROBES = 128;

// find-or-insert; returns the slot, or nullptr when no slot is free within TS_MAX_PROBES (saturated set:
// the caller emits the row anyway — a duplicate flow_tag row is idempotent in ClickHouse,
// flow_tag.go:153-160 — and the next rebuild makes room)
__device__ __forceinline__ FtEntry* ts_upsert(FtEntry* tbl, uint32_t mask, U128 fp, TagSetState* st) {
  uint32_t s = (uint32_t)(fp.lo ^ (fp.hi >> 17)) & mask;
  for (uint32_t probes = 0; probes < TS_MAX_PROBES; probes++) {
    FtEntry* e = tbl + s;
    unsigned long long lo, hi;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(e) : "memory");
    if (lo == fp.lo && hi == fp.hi) return e;
    if ((lo | hi) == 0) {
      U128 old = cas128(e, U128{0, 0}, fp);
      if ((old.lo | old.hi) == 0) {
        atomicAdd(&st->count, 1u);
        return e;
      }
      if (old.lo == fp.lo && old.hi == fp.hi) return e;
    } else if (lo == 0 || hi == 0) {
      continue;  // torn read of a concurrent 128-bit publish: re-read this slot
    }
    s = (s + 1) & mask;
  }
  return nullptr;
}
// scratch sets (the reduce's grouping): find-or-insert, flag the caller when the set is too small
__device__ __forceinline__ FtEntry* ts_upsert_flag(FtEntry* tbl, uint32_t mask, U128 fp, uint32_t* full_flag) {
  uint32_t s = (uint32_t)(fp.lo ^ (fp.hi >> 17)) & mask;
  for (uint32_t probes = 0; probes < 512; probes++) {
    FtEntry* e = tbl + s;
    unsigned long long lo, hi;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(e) : "memory");
    if (lo == fp.lo && hi == fp.hi) return e;
    if ((lo | hi) == 0) {
      U128 old = cas128(e, U128{0, 0}, fp);
      if ((old.lo | old.hi) == 0 || (old.lo == fp.lo && old.hi == fp.hi)) return e;
    } else if (lo == 0 || hi == 0) {
      continue;  // torn read of a concurrent 128-bit publish: re-read this slot
    }
    s = (s + 1) & mask;
  }
  *full_flag = 1;
  return nullptr;
}
__device__ __forceinline__ FtEntry* ts_find(FtEntry* tbl, uint32_t mask, U128 fp) {
  uint32_t s = (uint32_t)(fp.lo ^ (fp.hi >> 17)) & mask;
  for (uint32_t probes = 0; probes < TS_MAX_PROBES; probes++) {
    FtEntry* e = tbl + s;
    if (e->fp_lo == fp.lo && e->fp_hi == fp.hi) return e;
    if ((e->fp_lo | e->fp_hi) == 0) return nullptr;
    s = (s + 1) & mask;
  }
  return nullptr;
}

// The first slot of a probe sequence, whole (one 32-byte sector, two loads issued together), so that
// several probes can be in flight before any is consumed.
struct TsPeek {
  unsigned long long lo, hi, first_pos, ts;  // ts = time | state << 32
};
__device__ __forceinline__ void ts_peek(const FtEntry* tbl, uint32_t mask, U128 fp, TsPeek& k) {
  const FtEntry* e = tbl + ((uint32_t)(fp.lo ^ (fp.hi >> 17)) & mask);
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(k.lo), "=l"(k.hi) : "l"(e) : "memory");
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2+16];" : "=l"(k.first_pos), "=l"(k.ts) : "l"(e) : "memory");
}
// the key sits in its first slot, has been emitted, and its lru time has not expired: nothing to do
__device__ __forceinline__ bool ts_peek_fresh(const TsPeek& k, U128 fp, uint32_t now, uint32_t timeout) {
  if (k.lo != fp.lo || k.hi != fp.hi) return false;
  const uint32_t time = (uint32_t)k.ts, state = (uint32_t)(k.ts >> 32);
  if (state == 0) return false;
  return !((uint32_t)(time + timeout) < now);
}

// lru semantics (l7_flow_log.go:693-701, app_service_tag_writer.go:100-108): emit when
// the key is new, or when old + CacheFlushTimeout < time (uint32 arithmetic)
__device__ __forceinline__ bool ts_candidate(const FtEntry* e, uint32_t now, uint32_t timeout) {
  if (e->state == 0) return true;
  return (uint32_t)(e->time + timeout) < now;
}

// epoch-tagged position for atomicMax(first_pos): larger = newer batch, then earlier position
__device__ __forceinline__ unsigned long long ts_pos_key(uint32_t epoch, unsigned long long pos) {
  return (unsigned long long)(epoch & 0xffffu) << 48 | (0xffffffffffffull - pos);
}

// ---- host side: a set with a spare table for rebuilds (tagset.cu) ---------------------------------------
struct TagSet {
  FtEntry* cur = nullptr;
  FtEntry* alt = nullptr;
  TagSetState* d_state = nullptr;
  uint32_t mask = 0;
  uint32_t timeout = 0;
  uint32_t count_seen = 0;  // TagSetState::count as of the last fetch
  uint32_t slots() const { return mask + 1; }
};
int tagset_create(TagSet& t, uint64_t min_slots, uint32_t timeout);
void tagset_destroy(TagSet& t);
// Reclaims slots when the set is more than half full: drops expired entries (exact: an expired key is
// emitted again exactly like an absent one) and, if the set is still more than a quarter full, the oldest
// entries by lru time (approximates the reference LRU's eviction, libs/lru/lru.go:52-57; evicted keys
// are emitted again when they come back).  Asynchronous on `s`; swaps cur / alt.  Returns 1 when it ran.
int tagset_maybe_rebuild(TagSet& t, cudaStream_t s, uint64_t* launches);

}  // namespace dfgpu
