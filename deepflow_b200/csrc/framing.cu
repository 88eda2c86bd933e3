/*
 * framing.cu — kernels shared by every message type: record framing
 * (SimpleDecoder.ReadPB / IsEnd, server/libs/codec/simple_codec.go:229-248,404),
 * the per-slot exclusive scans that turn per-warp sizes into arena offsets, stream
 * placement + LowCardinality ranking, and the end-of-batch counter commit.
 */
#include "pipeline.cuh"

namespace dfgpu {

// flow-metrics/tag.go:211-233; index = canonical code of tapside_code(), 0 = unnamed slot ""
extern const char* const kTapSideNames[18] = {"", "rest", "c", "s", "local", "c-nd", "s-nd", "c-hv", "s-hv", "c-gw-hv",
                                              "s-gw-hv", "c-gw", "s-gw", "c-p", "s-p", "c-app", "s-app", "app"};

// ---- framing ----------------------------------------------------------------------
// One walker per frame follows the {u32 LE len, pb}* chain (inherently serial: each length says
// where the next one is).  It counts the records and drops each record's payload offset into a scratch
// slot derived from the frame's own byte range (a frame of `len` bytes holds at most len/4 records),
// so the per-record arrays can be filled afterwards by a fully parallel kernel.
//
// The chain is a pointer chase: ~600 dependent loads per 256 KiB frame.  Chasing through L2 (with an
// L2 prefetch ahead) cost 0.38 ms per 625 k records, i.e. ~0.6 us per hop.  Now the frame streams
// through shared memory instead: one CTA (one working thread) per frame keeps a ring of kStages x
// kWin bytes filled by bulk async copies (cp.async.bulk global -> shared, completion on an mbarrier
// per stage), and the chase reads its headers from the ring at shared-memory latency.  The input is
// read once at HBM rate; the serial part is ~600 x (LDS + a few ALU ops).
namespace {
constexpr uint32_t kWin = 4096, kStages = 6, kRing = kWin * kStages;
constexpr uint32_t kSmallFrame = 2048;  // below this a handful of direct loads beats setting the ring up

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const uint8_t* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
}  // namespace

__global__ void __launch_bounds__(32) k_frame_walk(const BatchParams p) {
  __shared__ __align__(128) uint8_t ring[kRing];
  __shared__ __align__(8) uint64_t bars[kStages];
  const uint32_t f = blockIdx.x;
  if (threadIdx.x != 0 || f >= p.nframes) return;
  const dfgpu_frame fr = p.frames[f];
  const uint8_t* buf = p.in + fr.offset;
  uint32_t* scratch = p.rec_scratch + (fr.offset >> 2);
  const uint32_t len = fr.len;
  uint32_t off = 0, n = 0, err = 0;

  // the 16-byte aligned interior of the frame is what the bulk copies may touch
  const uintptr_t a0 = (uintptr_t)buf, a1 = a0 + len;
  const uintptr_t lo16 = (a0 + 15) & ~(uintptr_t)15, hi16 = a1 & ~(uintptr_t)15;
  const bool staged = len >= kSmallFrame && hi16 > lo16;
  const uint32_t span = staged ? (uint32_t)(hi16 - lo16) : 0;
  const uint32_t nwin = (span + kWin - 1) / kWin;
  const uint32_t lo_off = (uint32_t)(lo16 - a0);  // frame offset of ring byte 0
  uint32_t issued = 0, waited = 0;
  const uint32_t ring_s = smem_addr(ring), bar_s = smem_addr(bars);
  auto issue = [&](uint32_t w) {  // window w -> stage w % kStages
    const uint32_t st = w % kStages, o = w * kWin;
    const uint32_t bytes = span - o < kWin ? span - o : kWin;
    bulk_load(ring_s + st * kWin, reinterpret_cast<const uint8_t*>(lo16) + o, bytes, bar_s + st * 8);
  };
  auto wait = [&](uint32_t w) { bar_wait(bar_s + (w % kStages) * 8, (w / kStages) & 1); };
  if (staged) {
    for (uint32_t s = 0; s < kStages; s++) bar_init(bar_s + s * 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (; issued < nwin && issued < kStages; issued++) issue(issued);
  }

  while (off < len) {  // !IsEnd()
    if (len - off < 4) { err = 1; break; }  // offset(+4) out of buf len
    uint32_t l;
    if (staged && off >= lo_off && off + 4 <= lo_off + span) {
      const uint32_t r = off - lo_off, w0 = r / kWin, w1 = (r + 3) / kWin;
      // a record longer than the ring: stream the skipped windows through (each reuse of a stage waits
      // for the window that held it before)
      while (issued <= w1) {
        while (issued >= kStages && waited <= issued - kStages) wait(waited++);
        issue(issued++);
      }
      while (waited <= w1) wait(waited++);
      l = (uint32_t)ring[r % kRing] | (uint32_t)ring[(r + 1) % kRing] << 8 | (uint32_t)ring[(r + 2) % kRing] << 16 |
          (uint32_t)ring[(r + 3) % kRing] << 24;
      // windows before w0 are done with: refill their stages (all of them have been waited: waited > w1)
      while (issued < nwin && issued < w0 + kStages) issue(issued++);
    } else {
      l = (uint32_t)__ldg(buf + off) | (uint32_t)__ldg(buf + off + 1) << 8 | (uint32_t)__ldg(buf + off + 2) << 16 |
          (uint32_t)__ldg(buf + off + 3) << 24;
    }
    off += 4;
    if (l > len - off) { err = 1; break; }
    scratch[n] = (uint32_t)(fr.offset + off);
    off += l;
    n++;
  }
  // no copy may still be in flight into this CTA's shared memory when it exits
  while (waited < issued) wait(waited++);
  p.frame_nrec[f] = n;
  p.frame_chain_err[f] = err;
}

// thread per record: locate its frame (binary search over the exclusive scan) and
// publish (offset, length, frame)
__global__ void k_frame_fill(const BatchParams p) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= batch_nrec(p)) return;
  uint32_t lo = 0, hi = p.nframes;  // last frame with frame_base <= r
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(p.frame_base + mid) <= r) lo = mid; else hi = mid;
  }
  // frames without records share a base with their successor: step to the owner
  while (lo + 1 < p.nframes && __ldg(p.frame_base + lo + 1) <= r) lo++;
  const uint32_t j = r - __ldg(p.frame_base + lo);
  const uint32_t off = p.rec_scratch[(p.frames[lo].offset >> 2) + j];
  const uint8_t* q = p.in + off - 4;
  p.rec_off[r] = off;
  p.rec_len[r] = (uint32_t)__ldg(q) | (uint32_t)__ldg(q + 1) << 8 | (uint32_t)__ldg(q + 2) << 16 | (uint32_t)__ldg(q + 3) << 24;
  p.rec_frame[r] = lo;
}

// exclusive scan of frame_nrec (single CTA, any nframes)
__global__ void k_frame_scan(const BatchParams p) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < p.nframes; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < p.nframes ? p.frame_nrec[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
      uint32_t t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < p.nframes) p.frame_base[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += sh[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) p.dyn->stats.nrec = carry;
}

// Everything the record count decides, planned on the device right after framing (the host never
// waits for it): does the record index hold the batch, how many warps of records there are, and where
// every fixed-width column of the main table lives inside the arena (dense, for exactly nrec rows).
__global__ void k_prepare(const BatchParams p) {
  DynState* d = p.dyn;
  if (threadIdx.x != 0) return;
  uint32_t nrec = d->stats.nrec;
  if (nrec > p.rec_cap) {
    d->rec_overflow = 1;
    d->overflow = 1;
    nrec = 0;
  }
  d->nwarps = (nrec + 31) / 32;
  const uint32_t rows = d->plan_valid ? d->plan_rows : nrec;
  uint64_t cur = 0;
  for (uint32_t c = 0; c < p.ncols; c++) {
    d->fix_off[c] = cur;
    cur += ((uint64_t)rows * p.col_es[c] + 255) & ~255ull;
    d->aux_off[c] = cur;
    cur += ((uint64_t)rows * p.col_aux[c] + 255) & ~255ull;
  }
  d->var_base = cur;
  if (cur > p.arena_bytes) d->overflow = 1;
}

// ---- slot scan: one CTA per slot, exclusive scan over the warps ----------------------
// Each thread owns a contiguous chunk of the slot's per-warp totals: chunk sums, ONE block-wide scan of
// the 1024 chunk sums (shuffles), then the chunk is written out.  (The first version scanned 1024
// elements at a time with a 10-step shared-memory ladder: 45 us per launch for 19.5 k warps, all of it
// barrier latency on an otherwise idle GPU.)
__global__ void __launch_bounds__(1024) k_slot_scan(const BatchParams p, uint64_t* totals_out) {
  __shared__ uint64_t wsum[32];
  const uint32_t slot = blockIdx.x, n = p.dyn->nwarps;
  const uint32_t* tot = p.warp_tot + (size_t)slot * p.warp_cap;
  uint32_t* out = p.warp_base + (size_t)slot * p.warp_cap;
  const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
  uint64_t sum = 0;
  for (uint32_t i = lo; i < hi; i++) sum += tot[i];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t inc = sum;  // inclusive scan inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint64_t w = lane < (blockDim.x >> 5) ? wsum[lane] : 0, winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint64_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    wsum[lane] = winc - w;  // exclusive base of each warp
    if (lane == 31) totals_out[slot] = winc;
  }
  __syncthreads();
  uint64_t run = wsum[wid] + inc - sum;
  for (uint32_t i = lo; i < hi; i++) {
    const uint32_t v = tot[i];
    out[i] = (uint32_t)run;
    run += v;
  }
}

// ---- finish: stream placement + LC ranks (one CTA, 256 threads) --------------------
__global__ void k_finish(const BatchParams p) {
  DynState* d = p.dyn;
  __shared__ uint64_t slot_total[MAX_SLOTS];
  if (threadIdx.x < MAX_SLOTS) slot_total[threadIdx.x] = threadIdx.x < p.nslots ? d->stats.var_total[threadIdx.x] : 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t nrows = slot_total[0];
    d->stats.nrows = (uint32_t)nrows;
    uint64_t cur = d->var_base;
    for (uint32_t v = 0; v < p.nstreams; v++) {
      StreamSpec s = p.streams[v];
      uint64_t elems = s.slot_elems >= 0 ? slot_total[s.slot_elems] : nrows;
      uint64_t bytes = s.slot_bytes >= 0 ? slot_total[s.slot_bytes] : elems * s.elem_size;
      d->var_pos_off[v] = cur;
      if (s.has_pos && !p.native_str) cur += (elems * 16 + 255) & ~255ull;
      d->var_buf_off[v] = cur;
      cur += (bytes + 255) & ~255ull;
      d->stats.var_total[v] = elems;
      d->stats.var_bytes[v] = bytes;
    }
    d->var_end = cur;
    if (p.native_str) {
      // native strings: the Position arrays (one per value of every string stream) are device-only
      // working state of the flow_tag stage; they sit behind everything that is copied to the host
      for (uint32_t v = 0; v < p.nstreams; v++) {
        const StreamSpec s = p.streams[v];
        if (s.slot_bytes < 0) continue;  // not a string stream
        d->var_pos_off[v] = cur;
        cur += (d->stats.var_total[v] * 16 + 255) & ~255ull;
      }
    }
    d->dev_end = cur;
    if (cur > p.arena_bytes) d->overflow = 1;
  }
  // rank of each LowCardinality code by first appearance (ch-go builds the
  // dictionary in Append order)
  for (uint32_t sp = 0; sp < p.n_lc; sp++) {
    uint32_t mine = d->stats.lc_first[sp][threadIdx.x];
    uint32_t rank = 0;
    if (mine != NO_BAD) {
      for (int c = 0; c < 256; c++) rank += d->stats.lc_first[sp][c] < mine;
    }
    d->lc_rank[sp][threadIdx.x] = (uint8_t)rank;
  }
}

// end of a batch: advance the device-resident counters exactly once
__global__ void k_commit(const BatchParams p, int is_l7) {
  DynState* d = p.dyn;
  if ((d->stats.need_repair && !p.repair) || d->overflow) return;
  p.persist->id_counter += d->stats.nrows;
  if (is_l7 && d->stats.last_trace_id_index != 0xffffffffffffffffull && !p.cfg_trace_disabled)
    p.persist->last_trace_id_index = d->stats.last_trace_id_index;
}


void launch_frame_count(const BatchParams& p, cudaStream_t s) {
  if (!p.nframes) return;
  k_frame_walk<<<p.nframes, 32, 0, s>>>(p);
  k_frame_scan<<<1, 1024, 0, s>>>(p);
}
void launch_prepare(const BatchParams& p, cudaStream_t s) { k_prepare<<<1, 32, 0, s>>>(p); }
void launch_frame_fill(const BatchParams& p, cudaStream_t s) {
  if (!p.nframes) return;
  k_frame_fill<<<rec_grid(p), 256, 0, s>>>(p);
}
void launch_slot_scan(const BatchParams& p, cudaStream_t s) {
  // totals per SLOT land in stats.var_total (scratch until k_finish rewrites it per stream)
  k_slot_scan<<<p.nslots, 1024, 0, s>>>(p, p.dyn->stats.var_total);
}
void launch_slot_scan_to(const BatchParams& p, uint32_t nslots, uint64_t* totals_out, cudaStream_t s) {
  k_slot_scan<<<nslots, 1024, 0, s>>>(p, totals_out);
}
void launch_finish(const BatchParams& p, cudaStream_t s) { k_finish<<<1, 256, 0, s>>>(p); }
void launch_commit(const BatchParams& p, int is_l7, cudaStream_t s) { k_commit<<<1, 1, 0, s>>>(p, is_l7); }

}  // namespace dfgpu
