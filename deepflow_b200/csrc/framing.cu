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
// One thread walks one frame's {u32 LE len, pb}* chain (the chain is inherently serial:
// each length says where the next one is).  The walk counts the records and drops each
// record's payload offset into a scratch slot derived from the frame's own byte range
// (a frame of `len` bytes holds at most len/4 records), so the per-record arrays can
// be filled afterwards by a fully parallel kernel instead of a second serial walk.
// The dependent loads are ~DRAM-latency bound; an L2 prefetch a few records ahead
// (records are a few hundred bytes) turns most of them into L2 hits.
__global__ void k_frame_walk(const BatchParams p) {
  uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.nframes) return;
  const dfgpu_frame fr = p.frames[f];
  const uint8_t* buf = p.in + fr.offset;
  uint32_t* scratch = p.rec_scratch + (fr.offset >> 2);
  uint32_t len = fr.len, off = 0, n = 0, err = 0;
  constexpr uint32_t kFramePrefetch = 4096;  // bytes of look-ahead (~9 records: > DRAM latency / time per record)
  uint32_t pf = 0;                           // next line not prefetched yet
  while (off < len) {  // !IsEnd()
    if (len - off < 4) { err = 1; break; }  // offset(+4) out of buf len
    uint32_t l = (uint32_t)__ldg(buf + off) | (uint32_t)__ldg(buf + off + 1) << 8 |
                 (uint32_t)__ldg(buf + off + 2) << 16 | (uint32_t)__ldg(buf + off + 3) << 24;
    off += 4;
    if (l > len - off) { err = 1; break; }
    // read-ahead: every 128-byte line this record's successor-at-distance will need its header from.
    // One line per record (the first version) covered only ~1 in 3.4 lines of a 439-byte record, so
    // most headers still paid DRAM latency.
    for (uint32_t x = pf; x < off + l + kFramePrefetch && x < len; x += 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(buf + x));
    pf = max(pf, ((off + l + kFramePrefetch) + 127u) & ~127u);
    scratch[n] = (uint32_t)(fr.offset + off);
    off += l;
    n++;
  }
  p.frame_nrec[f] = n;
  p.frame_chain_err[f] = err;
}

// thread per record: locate its frame (binary search over the exclusive scan) and
// publish (offset, length, frame)
__global__ void k_frame_fill(const BatchParams p) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.nrec) return;
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

// ---- slot scan: one CTA per slot, exclusive scan over the warps ----------------------
__global__ void k_slot_scan(const BatchParams p, uint64_t* totals_out) {
  __shared__ uint32_t sh[1024];
  __shared__ uint64_t carry;
  uint32_t slot = blockIdx.x;
  const uint32_t* tot = p.warp_tot + (size_t)slot * p.nwarps;
  uint32_t* out = p.warp_base + (size_t)slot * p.nwarps;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < p.nwarps; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < p.nwarps ? tot[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
      uint32_t t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < p.nwarps) out[i] = (uint32_t)carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += sh[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) totals_out[slot] = carry;
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
    uint64_t cur = p.var_base;
    for (uint32_t v = 0; v < p.nstreams; v++) {
      StreamSpec s = p.streams[v];
      uint64_t elems = s.slot_elems >= 0 ? slot_total[s.slot_elems] : nrows;
      uint64_t bytes = s.slot_bytes >= 0 ? slot_total[s.slot_bytes] : elems * s.elem_size;
      d->var_pos_off[v] = cur;
      if (s.has_pos) cur += (elems * 16 + 255) & ~255ull;
      d->var_buf_off[v] = cur;
      cur += (bytes + 255) & ~255ull;
      d->stats.var_total[v] = elems;
      d->stats.var_bytes[v] = bytes;
    }
    d->var_end = cur;
    d->overflow = cur > p.arena_bytes ? 1 : 0;
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
  k_frame_walk<<<(p.nframes + 31) / 32, 32, 0, s>>>(p);
  k_frame_scan<<<1, 1024, 0, s>>>(p);
}
void launch_frame_fill(const BatchParams& p, cudaStream_t s) {
  if (!p.nframes || !p.nrec) return;
  k_frame_fill<<<(p.nrec + 255) / 256, 256, 0, s>>>(p);
}
void launch_slot_scan(const BatchParams& p, cudaStream_t s) {
  if (!p.nrec) return;
  // totals per SLOT land in stats.var_total (scratch until k_finish rewrites it per stream)
  k_slot_scan<<<p.nslots, 1024, 0, s>>>(p, p.dyn->stats.var_total);
}
void launch_slot_scan_to(const BatchParams& p, uint32_t nslots, uint64_t* totals_out, cudaStream_t s) {
  k_slot_scan<<<nslots, 1024, 0, s>>>(p, totals_out);
}
void launch_finish(const BatchParams& p, cudaStream_t s) { k_finish<<<1, 256, 0, s>>>(p); }
void launch_commit(const BatchParams& p, int is_l7, cudaStream_t s) { k_commit<<<1, 1, 0, s>>>(p, is_l7); }

}  // namespace dfgpu
