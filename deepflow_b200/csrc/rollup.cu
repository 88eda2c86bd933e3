/*
 * rollup.cu — the flow_metrics rollup accumulator (BASELINE config 3: 1s -> 1m).
 *
 * A persistent device hash table per flow_metrics table, keyed by (time bucket, every tag column of the
 * table), holding the RAW meters of each group.  Every dfgpu_rollup_add folds the rows of one batch into
 * it with the ConcurrentMerge rules of the reference:
 *   FlowMeter.ConcurrentMerge  flow-metrics/flow_meter.go:91-99 -> Traffic basic_meter.go:94-113 (sums,
 *   DirectionScore = max), Latency :277-315 (maxes, u64 sums, u32 counts), Performance :417-424,
 *   Anomaly :556-577 (sums; OooTx / OooRx ASSIGNED from the merged row; the derived *Fail fields :540-546
 *   are not merged: they stay those of the group's first row), FlowLoad :707-709; AppMeter / UsageMeter
 *   analogues (app_meter.go, usage_meter.go)
 * and dfgpu_rollup_flush emits the groups, in first-appearance order, as a column block of the table.
 *
 * A batch is folded without serialising on hot groups (traffic is Zipf distributed):
 *   key     thread per row: 128-bit fingerprint of the tag columns -> find-or-insert in the persistent
 *           table (atom.cas.b128); first / last row sequence numbers per group (atomicMax)
 *   rank    new groups are numbered in first-appearance order (ballot + scan over "row is its group's first")
 *   sort    counting sort of the batch's row indices by group slot (warp-aggregated atomics)
 *   fold    warps take fixed-size chunks of the sorted order (balanced whatever the skew), read each row's
 *           raw meters as one contiguous segment (lane = field) and touch the accumulator once per
 *           (chunk, group): atomics per field <= rows / CHUNK + groups, not rows
 *   edge    rows that are their group's first / last write the FST / LST fields
 * Algorithmic bytes = N x (key + meters) + K x (key + meters); this design adds N x 12 B of index traffic.
 *
 * Identity of a group is its 128-bit fingerprint (no full-key compare; collision probability for 10^6
 * groups ~ 10^-27).
 */
#include <algorithm>
#include <cstring>

#include "metrics_host.h"

namespace dfgpu {

namespace {

constexpr int RU_MAX_TAG = 72;
constexpr uint32_t RU_CHUNK = 128;  // sorted positions per warp in the fold

struct RuGroup {  // 32 B, one HBM sector
  unsigned long long fp_lo, fp_hi;  // 0/0 = empty
  unsigned long long first_inv;     // ~(sequence number of the group's first row); atomicMax keeps the first
  unsigned long long last_seq;      // sequence number + 1 of its last row
};

struct RuSrc {  // one batch of rows: columnar tags + row-major raw meters
  uint32_t n, ntag, nf, bucket;
  uint64_t seq_base;  // rows are numbered seq_base + row: later batches have larger numbers
  const uint8_t* col[RU_MAX_TAG];  // fixed kinds: values (elem stride); strings: bytes
  const uint8_t* aux[RU_MAX_TAG];  // strings: dfgpu_pos per row
  uint8_t kind[RU_MAX_TAG], elem[RU_MAX_TAG], which[RU_MAX_TAG];
  // plain fixed-width columns (no bucket / rank translation) grouped by element width, for the key kernel:
  // plain[w] .. plain[w + 1] index into plain_col (w = 0..3: 1, 2, 4, 8 bytes); the others are in special
  uint8_t plain_col[RU_MAX_TAG], special_col[RU_MAX_TAG];
  uint8_t plain[5], nspecial;
  const uint64_t* raw;      // [n][nf]
  const uint8_t* obs_inv;   // [256] LowCardinality rank -> canonical tap-side code (nullptr: the column holds codes)
};

struct RuDev {  // the persistent accumulator of one table
  RuGroup* groups;
  uint8_t* keys;       // [slots][key_stride] tag values of the creating row (time = bucket start; strings: heap off, len)
  uint64_t* acc;       // [slots][nf]
  uint32_t* order_id;  // [slots] rank of the group in first-appearance order
  uint32_t* slen;      // [3][slots] string lengths by order id (app tables)
  uint32_t* cnt;       // [slots] scratch of the counting sort (all zero between adds)
  uint32_t* off;       // [slots]
  uint32_t* ctr;       // [0] groups, [1] table full, [2] heap full, [3] new groups of the running add
  uint8_t* heap;       // string bytes of the groups' first rows
  unsigned long long* heap_used;
  uint64_t heap_cap;
  uint32_t mask, key_stride, nf, ntag;
  uint16_t key_off[RU_MAX_TAG];
  uint8_t ops[M_MAX_FIELDS];
};

__device__ __forceinline__ U128 ru_cas128(RuGroup* e, U128 cmp, U128 val) {
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

__device__ __forceinline__ uint64_t ru_load(const uint8_t* src, uint32_t elem) {
  switch (elem) {
    case 1: return *src;
    case 2: return *reinterpret_cast<const uint16_t*>(src);
    case 4: return *reinterpret_cast<const uint32_t*>(src);
    default: return *reinterpret_cast<const uint64_t*>(src);
  }
}

__device__ __forceinline__ void ru_round(uint64_t& lo, uint64_t& hi, uint64_t a, uint64_t b, uint32_t c) {
  const uint64_t x = (lo ^ a) + (uint64_t)c * 0x9E3779B97F4A7C15ull;
  lo = ((x << 27) | (x >> 37)) * 0xff51afd7ed558ccdull;
  const uint64_t y = (hi + b) ^ ((uint64_t)c << 32);
  hi = ((y << 31) | (y >> 33)) * 0xc4ceb9fe1a85ec53ull;
}
// plain columns k0..k1 of one element width: four loads in flight per thread, then their rounds
template <class T>
__device__ __forceinline__ void ru_hash_cols(const RuSrc& s, uint32_t k0, uint32_t k1, uint64_t row, uint64_t& lo, uint64_t& hi) {
  uint32_t k = k0;
  for (; k + 4 <= k1; k += 4) {
    const uint32_t c0 = s.plain_col[k], c1 = s.plain_col[k + 1], c2 = s.plain_col[k + 2], c3 = s.plain_col[k + 3];
    const uint64_t a0 = reinterpret_cast<const T*>(s.col[c0])[row], a1 = reinterpret_cast<const T*>(s.col[c1])[row];
    const uint64_t a2 = reinterpret_cast<const T*>(s.col[c2])[row], a3 = reinterpret_cast<const T*>(s.col[c3])[row];
    ru_round(lo, hi, a0, a0 * 0x9E3779B97F4A7C15ull, c0);
    ru_round(lo, hi, a1, a1 * 0x9E3779B97F4A7C15ull, c1);
    ru_round(lo, hi, a2, a2 * 0x9E3779B97F4A7C15ull, c2);
    ru_round(lo, hi, a3, a3 * 0x9E3779B97F4A7C15ull, c3);
  }
  for (; k < k1; k++) {
    const uint32_t c = s.plain_col[k];
    const uint64_t a = reinterpret_cast<const T*>(s.col[c])[row];
    ru_round(lo, hi, a, a * 0x9E3779B97F4A7C15ull, c);
  }
}

// ---- key: fingerprint, find-or-insert, first / last sequence numbers -----------------------------------
__global__ void __launch_bounds__(256) k_ru_key(const RuDev d, const __grid_constant__ RuSrc s, uint32_t* gslot) {
  for (uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < s.n; row += (uint64_t)gridDim.x * blockDim.x) {
    // 128-bit fingerprint of (time bucket, every tag column): two accumulators, one multiply-rotate round each
    // per column (the column index enters the round), avalanche at the end (fp_finish).  Plain fixed-width
    // columns are walked by element width, without a switch per value.  (r2: a general loop with two full
    // mix64 per column was 7 000 instructions per row — 52 of the 75 ms of a 100 M-row add.)
    uint64_t lo = 0x243F6A8885A308D3ull, hi = 0x13198A2E03707344ull;
#define RU_ROUND(a_, b_, c_) ru_round(lo, hi, (a_), (b_), (c_))
    ru_hash_cols<uint8_t>(s, s.plain[0], s.plain[1], row, lo, hi);
    ru_hash_cols<uint16_t>(s, s.plain[1], s.plain[2], row, lo, hi);
    ru_hash_cols<uint32_t>(s, s.plain[2], s.plain[3], row, lo, hi);
    ru_hash_cols<uint64_t>(s, s.plain[3], s.plain[4], row, lo, hi);
    for (uint32_t k = 0; k < s.nspecial; k++) {  // time, observation_point, IPv6, strings
      const uint32_t c = s.special_col[k];
      uint64_t a = 0, b = 0;
      const uint32_t kind = s.kind[c], elem = s.elem[c];
      if (kind == MK_LCS || kind == MK_STR) {
        const dfgpu_pos ps = reinterpret_cast<const dfgpu_pos*>(s.aux[c])[row];
        const StrHash h = str_hash(s.col[c] + ps.start, (uint32_t)(ps.end - ps.start));
        a = h.a;
        b = h.b;
      } else if (kind == MK_IP6) {
        const ulonglong2 v = reinterpret_cast<const ulonglong2*>(s.col[c])[row];
        a = v.x;
        b = v.y;
      } else {
        a = ru_load(s.col[c] + row * elem, elem);
        if (kind == MK_OBS && s.obs_inv) a = s.obs_inv[a & 0xff];
        if (c == 0) a = (uint32_t)a / s.bucket * s.bucket;  // the time column: bucket start
        b = a * 0x9E3779B97F4A7C15ull;
      }
      RU_ROUND(a, b, c);
    }
#undef RU_ROUND
    const U128 fp = fp_finish(lo, hi);
    uint32_t slot = (uint32_t)(fp.lo ^ (fp.hi >> 17)) & d.mask;
    RuGroup* g = nullptr;
    bool created = false;
    for (uint32_t probes = 0; probes < 1024; probes++) {
      RuGroup* e = d.groups + slot;
      unsigned long long elo, ehi;
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(elo), "=l"(ehi) : "l"(e) : "memory");
      if (elo == fp.lo && ehi == fp.hi) {
        g = e;
        break;
      }
      if ((elo | ehi) == 0) {
        const U128 old = ru_cas128(e, U128{0, 0}, fp);
        if ((old.lo | old.hi) == 0) {
          g = e;
          created = true;
          break;
        }
        if (old.lo == fp.lo && old.hi == fp.hi) {
          g = e;
          break;
        }
      } else if (elo == 0 || ehi == 0) {
        continue;  // torn read of a concurrent publish: re-read this slot
      }
      slot = (slot + 1) & d.mask;
    }
    if (!g) {
      d.ctr[1] = 1;
      gslot[row] = 0xffffffffu;
      continue;
    }
    if (created) {
      // the creating row stores the group's tag values (all rows of a group carry the same ones)
      if (atomicAdd(&d.ctr[0], 1u) * 2 > d.mask) d.ctr[1] = 1;  // load factor above 1/2: caller must flush
      uint8_t* k = d.keys + (size_t)slot * d.key_stride;
      for (uint32_t c = 0; c < s.ntag; c++) {
        const uint32_t kind = s.kind[c], elem = s.elem[c];
        uint8_t* dst = k + d.key_off[c];
        if (kind == MK_LCS || kind == MK_STR) {
          const dfgpu_pos ps = reinterpret_cast<const dfgpu_pos*>(s.aux[c])[row];
          const uint32_t len = (uint32_t)(ps.end - ps.start);
          const unsigned long long ho = len ? atomicAdd(d.heap_used, (unsigned long long)len) : 0ull;
          uint32_t rec[2] = {(uint32_t)ho, len};
          if (ho + len > d.heap_cap) {
            d.ctr[2] = 1;
            rec[1] = 0;
          } else {
            const uint8_t* src = s.col[c] + ps.start;
            for (uint32_t i = 0; i < len; i++) d.heap[ho + i] = src[i];
          }
          memcpy(dst, rec, 8);
        } else if (kind == MK_IP6) {
          const ulonglong2 v = reinterpret_cast<const ulonglong2*>(s.col[c])[row];
          memcpy(dst, &v, 16);
        } else {
          uint64_t a = ru_load(s.col[c] + row * elem, elem);
          if (kind == MK_OBS && s.obs_inv) a = s.obs_inv[a & 0xff];
          if (c == 0) a = (uint32_t)a / s.bucket * s.bucket;
          for (uint32_t i = 0; i < elem; i++) dst[i] = (uint8_t)(a >> (8 * i));
        }
      }
    }
    // first row of the group: only a row EARLIER than the one recorded has anything to say, and after the first
    // few there is none — a plain read decides (a stale read only costs a redundant atomic).  The last row is
    // recorded by the fold, once per run of the group's rows.  (r2: one atomicMax pair per row was 55 of the 75 ms
    // of a 100 M-row add — the hot groups of a Zipf key distribution serialise on their two words.)
    const unsigned long long seq = s.seq_base + row;
    unsigned long long seen;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(seen) : "l"(&g->first_inv) : "memory");
    if (~seq > seen) atomicMax(&g->first_inv, ~seq);
    gslot[row] = slot;
  }
}

// ---- rank: new groups in first-appearance order --------------------------------------------------------
// grid covers ceil(n / 32) warps exactly once (no grid stride: the scan works on per-warp counts)
__global__ void k_ru_lead(const RuDev d, const uint32_t* gslot, uint32_t n, uint64_t seq_base, uint32_t* lead_cnt) {
  const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool lead = false;
  if (row < n) {
    const uint32_t sl = gslot[row];
    if (sl != 0xffffffffu) lead = d.groups[sl].first_inv == ~(seq_base + row);
  }
  const uint32_t m = __ballot_sync(0xffffffffu, lead);
  if ((threadIdx.x & 31) == 0) lead_cnt[row >> 5] = __popc(m);
}
__global__ void k_ru_rank(const RuDev d, const __grid_constant__ RuSrc s, const uint32_t* gslot, const uint32_t* lead_base) {
  const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  bool lead = false;
  uint32_t sl = 0xffffffffu;
  if (row < s.n) {
    sl = gslot[row];
    if (sl != 0xffffffffu) lead = d.groups[sl].first_inv == ~(s.seq_base + row);
  }
  const uint32_t m = __ballot_sync(0xffffffffu, lead);
  if (!lead) return;
  const uint32_t rank = d.ctr[3] + lead_base[row >> 5] + __popc(m & ((1u << lane) - 1));
  d.order_id[sl] = rank;
  for (uint32_t c = 0; c < s.ntag; c++)
    if (s.kind[c] == MK_LCS || s.kind[c] == MK_STR) {
      uint32_t rec[2];
      memcpy(rec, d.keys + (size_t)sl * d.key_stride + d.key_off[c], 8);
      d.slen[(size_t)s.which[c] * (d.mask + 1) + rank] = rec[1];
    }
}
__global__ void k_ru_rank_done(const RuDev d, const uint32_t* total_new) { d.ctr[3] += *total_new; }

// ---- counting sort of the row indices by group slot ------------------------------------------------------
__global__ void k_ru_hist(const RuDev d, const uint32_t* gslot, uint32_t n) {
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = base + threadIdx.x;
    const uint32_t sl = row < n ? gslot[row] : 0xffffffffu;
    const uint32_t m = __match_any_sync(0xffffffffu, sl);  // one atomic per distinct slot of the warp
    if (sl != 0xffffffffu && (int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&d.cnt[sl], __popc(m));
  }
}
__global__ void k_ru_scatter(const RuDev d, const uint32_t* gslot, uint32_t n, uint32_t* order) {
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < n; base += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t row = base + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t sl = row < n ? gslot[row] : 0xffffffffu;
    const uint32_t m = __match_any_sync(0xffffffffu, sl);
    const int leader = __ffs(m) - 1;
    uint32_t basepos = 0;
    // cnt counts down to zero again: no reset between adds
    if (sl != 0xffffffffu && (int)lane == leader) basepos = atomicSub(&d.cnt[sl], __popc(m)) - __popc(m);
    basepos = __shfl_sync(0xffffffffu, basepos, leader);
    if (sl != 0xffffffffu) order[d.off[sl] + basepos + __popc(m & ((1u << lane) - 1))] = (uint32_t)row;
  }
}

// ---- fold: balanced chunks of the sorted order ---------------------------------------------------------------
__device__ __forceinline__ uint64_t ru_op(uint32_t op, uint64_t a, uint64_t b) {
  return op <= MOP_S32 ? a + b : (op <= MOP_X8 ? (a < b ? b : a) : a);
}
__device__ __forceinline__ void ru_flush(const RuDev& d, uint32_t slot, uint32_t f, uint32_t op, uint64_t v) {
  if (v == 0 || op > MOP_X8) return;
  unsigned long long* p = reinterpret_cast<unsigned long long*>(d.acc + (size_t)slot * d.nf + f);
  if (op <= MOP_S32) atomicAdd(p, (unsigned long long)v);  // S32: wraps at 32 bits when the column is written
  else atomicMax(p, (unsigned long long)v);
}
__global__ void __launch_bounds__(256) k_ru_fold(const RuDev d, const __grid_constant__ RuSrc s, const uint32_t* gslot,
                                                const uint32_t* order, const uint32_t* n_sorted_dev) {
  const uint32_t n_sorted = *n_sorted_dev;  // rows that found a group (all of them unless the table is full)
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), total_warps = gridDim.x * (blockDim.x >> 5);
  const uint32_t nf = d.nf;
  const uint32_t f0 = lane, f1 = lane + 32, f2 = lane + 64;
  const uint32_t op0 = f0 < nf ? d.ops[f0] : MOP_FST, op1 = f1 < nf ? d.ops[f1] : MOP_FST, op2 = f2 < nf ? d.ops[f2] : MOP_FST;
  const uint32_t nchunks = (n_sorted + RU_CHUNK - 1) / RU_CHUNK;
  for (uint32_t ch = gwarp; ch < nchunks; ch += total_warps) {
    const uint32_t lo = ch * RU_CHUNK, hi = min(n_sorted, lo + RU_CHUNK);
    uint32_t cur = 0xffffffffu;
    uint64_t a0 = 0, a1 = 0, a2 = 0;
    unsigned long long run_max = 0;  // largest row index of the current run: the group's last row so far
    for (uint32_t base = lo; base < hi; base += 32) {
      const uint32_t i = base + lane;
      const uint32_t my_row = i < hi ? order[i] : 0u;
      const uint32_t my_slot = i < hi ? gslot[my_row] : 0xffffffffu;
      const uint32_t cnt = min(32u, hi - base);
      for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t row = __shfl_sync(0xffffffffu, my_row, j), slot = __shfl_sync(0xffffffffu, my_slot, j);
        const uint64_t* R = s.raw + (size_t)row * nf;
        const uint64_t v0 = f0 < nf ? __ldg(R + f0) : 0ull, v1 = f1 < nf ? __ldg(R + f1) : 0ull, v2 = f2 < nf ? __ldg(R + f2) : 0ull;
        if (slot != cur) {
          if (cur != 0xffffffffu) {
            ru_flush(d, cur, f0, op0, a0);
            ru_flush(d, cur, f1, op1, a1);
            ru_flush(d, cur, f2, op2, a2);
            if (lane == 0) atomicMax(&d.groups[cur].last_seq, s.seq_base + run_max + 1);
          }
          cur = slot;
          a0 = a1 = a2 = 0;
          run_max = 0;
        }
        run_max = row > run_max ? row : run_max;
        a0 = ru_op(op0, a0, v0);
        a1 = ru_op(op1, a1, v1);
        a2 = ru_op(op2, a2, v2);
      }
    }
    if (cur != 0xffffffffu) {
      ru_flush(d, cur, f0, op0, a0);
      ru_flush(d, cur, f1, op1, a1);
      ru_flush(d, cur, f2, op2, a2);
      if (lane == 0) atomicMax(&d.groups[cur].last_seq, s.seq_base + run_max + 1);
    }
  }
}
// rows that are their group's first / last: the fields that are not merged (first) or assigned (last)
__global__ void k_ru_edge(const RuDev d, const __grid_constant__ RuSrc s, const uint32_t* gslot) {
  for (uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < s.n; row += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t sl = gslot[row];
    if (sl == 0xffffffffu) continue;
    const RuGroup g = d.groups[sl];
    const unsigned long long seq = s.seq_base + row;
    const bool first = g.first_inv == ~seq, last = g.last_seq == seq + 1;
    if (!first && !last) continue;
    for (uint32_t f = 0; f < d.nf; f++) {
      const uint32_t op = d.ops[f];
      if ((op == MOP_FST && first) || (op == MOP_LST && last)) d.acc[(size_t)sl * d.nf + f] = s.raw[(size_t)row * d.nf + f];
    }
  }
}

// ---- multi-block exclusive scan of u32 (n up to 2^32) -------------------------------------------------------------
constexpr uint32_t SCAN_ITEMS = 2048;  // per CTA of 256 threads
__global__ void __launch_bounds__(256) k_scan_sums(const uint32_t* in, uint32_t n, uint32_t* bsum) {
  __shared__ uint32_t wsum[8];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_ITEMS;
  uint32_t v = 0;
  for (uint32_t k = 0; k < SCAN_ITEMS / 256; k++) {
    const uint64_t i = base + k * 256 + threadIdx.x;
    if (i < n) v += in[i];
  }
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; w++) t += wsum[w];
    bsum[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(1024) k_scan_top(uint32_t* bsum, uint32_t nb, uint32_t* total) {  // single CTA
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nb; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nb ? bsum[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t dd = 1; dd < 1024; dd <<= 1) {
      const uint32_t t = threadIdx.x >= dd ? sh[threadIdx.x - dd] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < nb) bsum[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0 && total) *total = carry;
}
__global__ void __launch_bounds__(256) k_scan_apply(const uint32_t* in, uint32_t* out, uint32_t n, const uint32_t* bsum) {
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t run;
  if (threadIdx.x == 0) run = bsum[blockIdx.x];
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_ITEMS;
  for (uint32_t k = 0; k < SCAN_ITEMS / 256; k++) {
    const uint64_t i = base + k * 256 + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0;
    const uint32_t e = warp_excl_scan(v);
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = e + v;
    __syncthreads();
    uint32_t wb = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) wb += wsum[w];
    if (i < n) out[i] = run + wb + e;
    __syncthreads();
    if (threadIdx.x == 255) run += wb + e + v;
    __syncthreads();
  }
}

// ---- flush: groups -> column block ----------------------------------------------------------------------------------
struct RuOut {
  uint8_t* out;
  const uint64_t* off;  // [2 * ncols]: values (or string bytes), aux (string positions)
  const uint32_t* soff;  // [3][slots] exclusive scan of slen
  uint32_t ncols, table, meter;
  const MDev* dv;
  uint32_t dbase;
};
__global__ void __launch_bounds__(256) k_ru_emit(const RuDev d, const RuOut o) {
  for (uint64_t sl = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; sl <= d.mask; sl += (uint64_t)gridDim.x * blockDim.x) {
    const RuGroup g = d.groups[sl];
    if ((g.fp_lo | g.fp_hi) == 0) continue;
    const uint32_t r = d.order_id[sl];
    const uint8_t* k = d.keys + (size_t)sl * d.key_stride;
    for (uint32_t c = 0; c < d.ntag; c++) {
      const MColDesc dc = o.dv->descs[o.dbase + c];
      const uint8_t* src = k + d.key_off[c];
      if (dc.kind == MK_LCS || dc.kind == MK_STR) {
        uint32_t rec[2];
        memcpy(rec, src, 8);
        const uint64_t so = o.soff[(size_t)dc.which * (d.mask + 1) + r];
        reinterpret_cast<dfgpu_pos*>(o.out + o.off[o.ncols + c])[r] = dfgpu_pos{(int64_t)so, (int64_t)(so + rec[1])};
        uint8_t* dst = o.out + o.off[c] + so;
        for (uint32_t i = 0; i < rec[1]; i++) dst[i] = d.heap[rec[0] + i];
      } else {
        uint8_t* dst = o.out + o.off[c] + (size_t)r * dc.elem;
        for (uint32_t i = 0; i < dc.elem; i++) dst[i] = src[i];
      }
    }
    const uint64_t* acc = d.acc + (size_t)sl * d.nf;
    uint32_t i = 0, c = d.ntag;
#define X(name, op) m.name = (decltype(m.name))acc[i++];
#define M(name, ch, ctype, expr)                                        \
  reinterpret_cast<ctype*>(o.out + o.off[c])[r] = (ctype)(expr);        \
  c++;
    if (o.meter == MT_METER_FLOW) {
      MFlowMeter m;
      M_FLOW_FIELDS(X)
      MT_FLOW_METER_COLUMNS(M)
    } else if (o.meter == MT_METER_APP) {
      MAppMeter m;
      M_APP_FIELDS(X)
      MT_APP_METER_COLUMNS(M)
    } else {
      MUsageMeter m;
      M_USAGE_FIELDS(X)
      MT_USAGE_METER_COLUMNS(M)
    }
#undef X
#undef M
  }
}

// ---- host state -------------------------------------------------------------------------------------------------------
struct RuTable {
  RuDev dev{};
  uint8_t* blob = nullptr;  // one allocation behind every pointer of dev
  uint64_t batches = 0;     // sequence number base of the next add: batches << 32... (rows per add < 2^32)
  bool dirty = false;
};
struct RuHost {
  RuTable t[MT_TABLE_COUNT];
  // scratch of an add, grown to the largest batch seen
  uint32_t* d_gslot = nullptr;
  uint32_t* d_order = nullptr;
  uint32_t* d_lead = nullptr;  // lead_cnt | lead_base | block sums
  uint64_t scratch_rows = 0;
  uint32_t* d_small = nullptr;  // [0] total new groups
  uint8_t* d_out = nullptr;     // flush output
  size_t out_cap = 0;
  uint64_t* d_ooff = nullptr;
  uint8_t* d_obs_inv = nullptr;
  uint32_t* h_ctr = nullptr;  // pinned
  uint8_t* h_inv = nullptr;   // pinned
};

}  // namespace

static RuHost* ru_host(dfgpu_ctx* ctx) {
  if (!ctx->ru) ctx->ru = new RuHost();
  return reinterpret_cast<RuHost*>(ctx->ru);
}

static int ru_ensure_table(dfgpu_ctx* ctx, MHost* mh, RuHost* h, int table) {
  RuTable& T = h->t[table];
  if (T.blob) return 0;
  const MDev& md = mh->dev;
  RuDev& d = T.dev;
  memset(&d, 0, sizeof d);
  d.ntag = md.ntag[table];
  d.nf = md.raw_nf[table];
  if (d.ntag > RU_MAX_TAG) return fail(ctx, DFGPU_EINVAL, "table %d has too many tag columns", table);
  uint32_t ko = 0;
  for (uint32_t c = 0; c < d.ntag; c++) {
    const MColDesc& dc = md.descs[md.desc_base[table] + c];
    d.key_off[c] = (uint16_t)ko;
    ko += (dc.kind == MK_LCS || dc.kind == MK_STR) ? 8 : dc.elem;
  }
  d.key_stride = (ko + 7) & ~7u;
  {
    uint32_t i = 0;
#define X(name, op) d.ops[i++] = op;
    if (md.table_meter[table] == MT_METER_FLOW) { M_FLOW_FIELDS(X) }
    else if (md.table_meter[table] == MT_METER_APP) { M_APP_FIELDS(X) }
    else { M_USAGE_FIELDS(X) }
#undef X
  }
  uint64_t groups = ctx->cfg.rollup_max_groups ? ctx->cfg.rollup_max_groups : (1u << 20);
  uint32_t slots = 1u << 10;
  while (slots < groups * 2 && slots < (1u << 30)) slots <<= 1;
  d.mask = slots - 1;
  d.heap_cap = m_app_index(table) >= 0 ? (uint64_t)groups * 64 : 256;
  size_t need = 0;
  auto carve = [&](size_t bytes) {
    size_t o = need;
    need += align_up(bytes + 16, 256);
    return o;
  };
  const size_t o_g = carve((size_t)slots * sizeof(RuGroup)), o_k = carve((size_t)slots * d.key_stride),
               o_a = carve((size_t)slots * d.nf * 8), o_o = carve((size_t)slots * 4), o_s = carve((size_t)slots * 12),
               o_c = carve((size_t)slots * 4), o_f = carve((size_t)slots * 4), o_ctr = carve(64), o_hu = carve(16),
               o_h = carve((size_t)d.heap_cap);
  CU(cudaMalloc(&T.blob, need));
  CU(cudaMemsetAsync(T.blob, 0, need, ctx->stream));
  uint8_t* B = T.blob;
  d.groups = (RuGroup*)(B + o_g);
  d.keys = B + o_k;
  d.acc = (uint64_t*)(B + o_a);
  d.order_id = (uint32_t*)(B + o_o);
  d.slen = (uint32_t*)(B + o_s);
  d.cnt = (uint32_t*)(B + o_c);
  d.off = (uint32_t*)(B + o_f);
  d.ctr = (uint32_t*)(B + o_ctr);
  d.heap_used = (unsigned long long*)(B + o_hu);
  d.heap = B + o_h;
  return 0;
}

static int ru_ensure_scratch(dfgpu_ctx* ctx, RuHost* h, uint64_t n) {
  if (!h->d_small) {
    CU(cudaMalloc(&h->d_small, 64));
    CU(cudaMalloc(&h->d_obs_inv, 256));
    CU(cudaHostAlloc(&h->h_ctr, 64, cudaHostAllocDefault));
    CU(cudaHostAlloc(&h->h_inv, 256, cudaHostAllocDefault));
  }
  if (n <= h->scratch_rows) return 0;
  if (h->d_gslot) CU(cudaFree(h->d_gslot));
  h->d_gslot = nullptr;
  const uint64_t cap = n + n / 8 + 4096;
  const uint64_t nw = cap / 32 + 2;
  // gslot | order | lead_cnt | lead_base | block sums
  CU(cudaMalloc(&h->d_gslot, (cap * 2 + nw * 2 + nw / SCAN_ITEMS + 4096) * 4));
  h->d_order = h->d_gslot + cap;
  h->d_lead = h->d_order + cap;
  h->scratch_rows = cap;
  return 0;
}

static uint32_t* ru_lead_base(RuHost* h) { return h->d_lead + (h->scratch_rows / 32 + 2); }
static uint32_t* ru_bsum(RuHost* h) { return ru_lead_base(h) + (h->scratch_rows / 32 + 2); }

static void ru_scan(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* bsum, uint32_t* total, cudaStream_t s) {
  const uint32_t nb = (n + SCAN_ITEMS - 1) / SCAN_ITEMS;
  if (nb) k_scan_sums<<<nb, 256, 0, s>>>(in, n, bsum);
  k_scan_top<<<1, 1024, 0, s>>>(bsum, nb, total);
  if (nb) k_scan_apply<<<nb, 256, 0, s>>>(in, out, n, bsum);
}

// folds one batch of rows (device resident) into the table's accumulator; asynchronous on the ctx stream
static int ru_add(dfgpu_ctx* ctx, MHost* mh, RuHost* h, int table, RuSrc& src) {
  int rc = ru_ensure_table(ctx, mh, h, table);
  if (rc) return rc;
  if (src.n == 0) return 0;
  rc = ru_ensure_scratch(ctx, h, src.n);
  if (rc) return rc;
  RuTable& T = h->t[table];
  cudaStream_t s = ctx->stream;
  src.seq_base = T.batches << 32;
  T.batches++;
  T.dirty = true;
  const uint32_t n = src.n;
  const unsigned grid = (unsigned)std::min<uint64_t>(((uint64_t)n + 255) / 256, 148 * 16);
  const unsigned rgrid = (unsigned)(((uint64_t)n + 255) / 256);
  const uint32_t nw = (n + 31) / 32;
  uint32_t* lead_cnt = h->d_lead;
  uint32_t* lead_base = ru_lead_base(h);
  uint32_t* bsum = ru_bsum(h);
  k_ru_key<<<grid, 256, 0, s>>>(T.dev, src, h->d_gslot);
  k_ru_lead<<<rgrid, 256, 0, s>>>(T.dev, h->d_gslot, n, src.seq_base, lead_cnt);
  ru_scan(lead_cnt, lead_base, nw, bsum, h->d_small, s);
  k_ru_rank<<<rgrid, 256, 0, s>>>(T.dev, src, h->d_gslot, lead_base);
  k_ru_rank_done<<<1, 1, 0, s>>>(T.dev, h->d_small);
  k_ru_hist<<<grid, 256, 0, s>>>(T.dev, h->d_gslot, n);
  ru_scan(T.dev.cnt, T.dev.off, T.dev.mask + 1, bsum, h->d_small + 1, s);
  k_ru_scatter<<<grid, 256, 0, s>>>(T.dev, h->d_gslot, n, h->d_order);
  // rows whose key found no slot (table full) are not in the order: the fold covers what was counted
  k_ru_fold<<<148 * 8, 256, 0, s>>>(T.dev, src, h->d_gslot, h->d_order, h->d_small + 1);
  k_ru_edge<<<grid, 256, 0, s>>>(T.dev, src, h->d_gslot);
  ctx->launches += 14;
  CU(cudaGetLastError());
  return 0;
}

void rollup_free_state(dfgpu_ctx* ctx) {
  RuHost* h = reinterpret_cast<RuHost*>(ctx->ru);
  if (!h) return;
  for (auto& T : h->t) cudaFree(T.blob);
  cudaFree(h->d_gslot);
  cudaFree(h->d_small);
  cudaFree(h->d_out);
  cudaFree(h->d_ooff);
  cudaFree(h->d_obs_inv);
  cudaFreeHost(h->h_ctr);
  cudaFreeHost(h->h_inv);
  delete h;
  ctx->ru = nullptr;
}

// rows of `table` of the last metrics run (still resident in the arena, raw meters beside them)
// groups the tag columns for the key kernel (RuSrc::plain_col / special_col)
static void ru_classify(RuSrc& src) {
  uint32_t np = 0;
  src.nspecial = 0;
  for (int w = 0; w < 4; w++) {
    src.plain[w] = (uint8_t)np;
    const uint32_t width = 1u << w;
    for (uint32_t c = 0; c < src.ntag; c++) {
      const bool plain = src.kind[c] == MK_FIX && c != 0 && (src.elem[c] == 1 || src.elem[c] == 2 || src.elem[c] == 4 || src.elem[c] == 8);
      if (plain && src.elem[c] == width) src.plain_col[np++] = (uint8_t)c;
    }
  }
  src.plain[4] = (uint8_t)np;
  for (uint32_t c = 0; c < src.ntag; c++) {
    const bool plain = src.kind[c] == MK_FIX && c != 0 && (src.elem[c] == 1 || src.elem[c] == 2 || src.elem[c] == 4 || src.elem[c] == 8);
    if (!plain) src.special_col[src.nspecial++] = (uint8_t)c;
  }
}

int rollup_add_resident(dfgpu_ctx* ctx, int table, uint32_t bucket) {
  MHost* mh = reinterpret_cast<MHost*>(ctx->mt);
  if (!mh) return fail(ctx, DFGPU_ENOENT, "no metrics batch resident on the device");
  if (!(mh->dev.raw_tables & (1u << table)))
    return fail(ctx, DFGPU_EINVAL, "table %d is not enabled in dfgpu_cfg.rollup_tables", table + DFGPU_TABLE_METRICS_BASE);
  RuHost* h = ru_host(ctx);
  const MLayout& L = *mh->h_lay;  // read back by the fetch of that batch
  RuSrc src;
  memset(&src, 0, sizeof src);
  src.n = L.nrows[table];
  src.ntag = mh->dev.ntag[table];
  src.nf = mh->dev.raw_nf[table];
  src.bucket = bucket;
  const uint8_t* A = ctx->bp.arena;
  bool has_obs = false;
  for (uint32_t c = 0; c < src.ntag; c++) {
    const uint32_t dsc = mh->dev.desc_base[table] + c;
    const MColDesc& dc = mh->dev.descs[dsc];
    src.kind[c] = dc.kind;
    src.elem[c] = dc.elem;
    src.which[c] = dc.which;
    src.col[c] = A + L.col_off[dsc];
    src.aux[c] = A + L.aux_off[dsc];
    has_obs |= dc.kind == MK_OBS;
  }
  src.raw = reinterpret_cast<const uint64_t*>(A + L.raw_off[table]);
  ru_classify(src);
  int rc = ru_ensure_scratch(ctx, h, src.n);
  if (rc) return rc;
  if (has_obs) {
    // observation_point keys of a batch are LowCardinality ranks (first appearance inside that batch):
    // translate back to the canonical tap-side code, which is what identifies the group across batches
    uint8_t* inv = h->h_inv;  // pinned; the previous add's copy completed before the fetch this call follows
    memset(inv, 0, 256);
    const uint32_t* first = ctx->h_dyn->stats.lc_first[m_obs_space(table)];
    std::vector<std::pair<uint32_t, uint32_t>> present;
    for (uint32_t k = 0; k < 256; k++)
      if (first[k] != NO_BAD) present.push_back({first[k], k});
    std::sort(present.begin(), present.end());
    for (size_t r = 0; r < present.size(); r++) inv[r] = (uint8_t)present[r].second;
    CU(cudaMemcpyAsync(h->d_obs_inv, inv, 256, cudaMemcpyHostToDevice, ctx->stream));
    src.obs_inv = h->d_obs_inv;
  }
  return ru_add(ctx, mh, h, table, src);
}

int metrics_ensure_state(dfgpu_ctx* ctx);  // metrics.cu

// decoded rows handed over by the caller (device resident): BASELINE config 3
int rollup_add_rows(dfgpu_ctx* ctx, int table, uint32_t bucket, const dfgpu_rollup_rows* rows) {
  int rc = metrics_ensure_state(ctx);
  if (rc) return rc;
  MHost* mh = reinterpret_cast<MHost*>(ctx->mt);
  RuHost* h = ru_host(ctx);
  if (rows->nrows >= (1ull << 32)) return fail(ctx, DFGPU_EINVAL, "at most 2^32 - 1 rows per add");
  if ((uint32_t)rows->ntag != mh->dev.ntag[table])
    return fail(ctx, DFGPU_EINVAL, "table %d has %u tag columns, %d given", table, mh->dev.ntag[table], rows->ntag);
  RuSrc src;
  memset(&src, 0, sizeof src);
  src.n = (uint32_t)rows->nrows;
  src.ntag = mh->dev.ntag[table];
  src.nf = mh->dev.raw_nf[table];
  src.bucket = bucket;
  for (uint32_t c = 0; c < src.ntag; c++) {
    const MColDesc& dc = mh->dev.descs[mh->dev.desc_base[table] + c];
    if (dc.kind == MK_LCS || dc.kind == MK_STR)
      return fail(ctx, DFGPU_EINVAL, "dfgpu_rollup_add_rows: tables with string tag columns are fed through dfgpu_submit");
    src.kind[c] = dc.kind;
    src.elem[c] = dc.elem;
    src.which[c] = dc.which;
    src.col[c] = reinterpret_cast<const uint8_t*>(rows->tag_cols[c]);
  }
  src.raw = rows->raw;
  ru_classify(src);
  src.obs_inv = nullptr;  // the observation_point column carries canonical tap-side codes
  return ru_add(ctx, mh, h, table, src);
}

// schema of a table's rollup input: tag columns (bytes per value, kind) and raw meter fields
int rollup_describe(dfgpu_ctx* ctx, int table, int* ntag, int* nfields, uint8_t* elem, uint8_t* kind) {
  int rc = metrics_ensure_state(ctx);
  if (rc) return rc;
  MHost* mh = reinterpret_cast<MHost*>(ctx->mt);
  if (ntag) *ntag = (int)mh->dev.ntag[table];
  if (nfields) *nfields = mh->dev.raw_nf[table];
  for (uint32_t c = 0; c < mh->dev.ntag[table]; c++) {
    const MColDesc& dc = mh->dev.descs[mh->dev.desc_base[table] + c];
    if (elem) elem[c] = dc.elem;
    if (kind) kind[c] = dc.kind;
  }
  return 0;
}

// every group of the table -> REDUCED column block, first-appearance order; the accumulator starts over
int rollup_flush(dfgpu_ctx* ctx, int table, dfgpu_batch* b, uint8_t* (*acquire)(dfgpu_ctx*, size_t)) {
  int rc = metrics_ensure_state(ctx);
  if (rc) return rc;
  MHost* mh = reinterpret_cast<MHost*>(ctx->mt);
  RuHost* h = ru_host(ctx);
  rc = ru_ensure_table(ctx, mh, h, table);
  if (rc) return rc;
  rc = ru_ensure_scratch(ctx, h, 4096);
  if (rc) return rc;
  RuTable& T = h->t[table];
  RuDev& d = T.dev;
  cudaStream_t s = ctx->stream;
  const auto& cs = mh->cols[table];
  const uint32_t ncols = (uint32_t)cs.size();
  const int app = m_app_index(table);
  const uint32_t slots = d.mask + 1;
  uint32_t* bsum = ru_bsum(h);
  CU(cudaMemcpyAsync(h->h_ctr, d.ctr, 16, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  const uint64_t K = h->h_ctr[0];
  if (h->h_ctr[1]) return fail(ctx, DFGPU_ENOMEM, "rollup accumulator of table %d is full (dfgpu_cfg.rollup_max_groups)", table + DFGPU_TABLE_METRICS_BASE);
  if (h->h_ctr[2]) return fail(ctx, DFGPU_ENOMEM, "rollup string heap of table %d is full", table + DFGPU_TABLE_METRICS_BASE);
  uint64_t sbytes[3] = {0, 0, 0};
  // output layout (same column walk as k_m_layout, K rows)
  std::vector<uint64_t> ooff(2 * (size_t)ncols, 0);
  // string totals need the scans first
  uint32_t* d_soff = nullptr;
  if (app >= 0 && K) {
    CU(cudaMalloc(&d_soff, (size_t)3 * slots * 4 + 64));
    for (int w = 0; w < 3; w++) ru_scan(d.slen + (size_t)w * slots, d_soff + (size_t)w * slots, (uint32_t)K, bsum, h->d_small + 4 + w, s);
    CU(cudaMemcpyAsync(h->h_ctr + 4, h->d_small + 4, 12, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int w = 0; w < 3; w++) sbytes[w] = h->h_ctr[4 + w];
  }
  const uint32_t dbase = mh->dev.desc_base[table];
  uint64_t cur = 0;
  for (uint32_t c = 0; c < ncols; c++) {
    const MColDesc& dc = mh->dev.descs[dbase + c];
    if (dc.kind == MK_LCS || dc.kind == MK_STR) {
      ooff[ncols + c] = cur;
      cur += align_up(K * 16, 256);
      ooff[c] = cur;
      cur += align_up(sbytes[dc.which], 256);
    } else {
      ooff[c] = cur;
      cur += align_up(K * dc.elem, 256);
    }
  }
  const uint64_t total = std::max<uint64_t>(cur, 256);
  if (total > h->out_cap) {
    if (h->d_out) CU(cudaFree(h->d_out));
    h->d_out = nullptr;
    CU(cudaMalloc(&h->d_out, (size_t)total + total / 4));
    h->out_cap = (size_t)total + total / 4;
  }
  if (!h->d_ooff) CU(cudaMalloc(&h->d_ooff, 2 * 1024 * 8));
  b->tables.emplace_back();
  dfgpu_batch::Table& tb = b->tables.back();
  tb.id = DFGPU_TABLE_METRICS_REDUCED;
  tb.cols.resize(ncols);
  b->host = acquire(ctx, (size_t)total + 64);
  if (!b->host) {
    if (d_soff) cudaFree(d_soff);
    return fail(ctx, DFGPU_ENOMEM, "pinned host allocation of %llu bytes failed", (unsigned long long)total);
  }
  uint8_t* hb = b->host + 64;
  if (K) {
    CU(cudaMemcpyAsync(h->d_ooff, ooff.data(), ooff.size() * 8, cudaMemcpyHostToDevice, s));
    RuOut o;
    memset(&o, 0, sizeof o);
    o.out = h->d_out;
    o.off = h->d_ooff;
    o.soff = d_soff;
    o.ncols = ncols;
    o.table = (uint32_t)table;
    o.meter = mh->dev.table_meter[table];
    o.dv = mh->d_dev;
    o.dbase = dbase;
    // the column descriptors live in MDev on the device: make sure a copy is there even if no batch ran
    CU(cudaMemcpyAsync(mh->d_dev, &mh->dev, sizeof(MDev), cudaMemcpyHostToDevice, s));
    k_ru_emit<<<148 * 8, 256, 0, s>>>(d, o);
    ctx->launches++;
    CU(cudaMemcpyAsync(hb, h->d_out, (size_t)total, cudaMemcpyDeviceToHost, s));
  }
  // start over: groups, accumulators, counters, string heap
  {
    const size_t upto = (size_t)((uint8_t*)d.heap - T.blob);
    CU(cudaMemsetAsync(T.blob, 0, upto, s));
    T.batches = 0;
    T.dirty = false;
  }
  CU(cudaStreamSynchronize(s));
  if (d_soff) cudaFree(d_soff);
  tb.nrows = K;
  static const std::vector<std::string> tap(kTapSideNames, kTapSideNames + 18);
  b->owned.reserve(4);
  b->owned_pos.reserve(4);
  for (uint32_t c = 0; c < ncols; c++) {
    const MColDesc& dc = mh->dev.descs[dbase + c];
    dfgpu_col& col = tb.cols[c];
    memset(&col, 0, sizeof col);
    col.name = cs[c].name;
    col.ch_type = cs[c].ch_type;
    col.elem_size = cs[c].elem_size;
    col.nrows = K;
    col.data = hb + ooff[c];
    if (dc.kind == MK_LCS || dc.kind == MK_STR) {
      col.pos = (const dfgpu_pos*)(hb + ooff[ncols + c]);
      col.data_bytes = sbytes[dc.which];
    } else {
      col.data_bytes = K * dc.elem;
      if (dc.kind == MK_OBS) {
        // the groups carry canonical tap-side codes: the LowCardinality dictionary is their names in
        // first-row order, the keys become ranks (<= 17 entries; ch-go builds it the same way at encode)
        uint8_t* keys = hb + ooff[c];
        int rank_of[256];
        for (int i = 0; i < 256; i++) rank_of[i] = -1;
        b->owned.emplace_back();
        b->owned_pos.emplace_back();
        auto& buf = b->owned.back();
        auto& pos = b->owned_pos.back();
        for (uint64_t r = 0; r < K; r++) {
          const uint32_t code = keys[r];
          if (rank_of[code] < 0) {
            rank_of[code] = (int)pos.size();
            const std::string& str = code < tap.size() ? tap[code] : tap[0];
            pos.push_back(dfgpu_pos{(int64_t)buf.size(), (int64_t)(buf.size() + str.size())});
            buf.insert(buf.end(), str.begin(), str.end());
          }
          keys[r] = (uint8_t)rank_of[code];
        }
        col.dict_data = buf.data();
        col.dict_pos = pos.data();
        col.dict_n = pos.size();
      }
    }
  }
  b->counters.rows[DFGPU_TABLE_METRICS_REDUCED] = K;
  return 0;
}

}  // namespace dfgpu
