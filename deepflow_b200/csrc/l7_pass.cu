/*
 * l7_pass.cu — the L7 decode kernel: AppProtoLogsData record -> the fixed-width columns of its
 * l7_flow_log row + the references the strings / SmartEncoding kernels work from.
 *
 * One thread per record, a warp = 32 consecutive records, persistent warps striding over the batch
 * (the record count lives on the device).  The record is decoded ONCE by the table-driven tokenizer
 * (pb_tok.cuh) into shared-memory words; the row is then built by every lane of the warp together
 * (l7_row.cuh) and each column leaves as one coalesced, full-sector store — no slot buffer in global
 * memory, no second decode.  Replaces, fused:
 *   decode          gogo Unmarshal via SimpleDecoder.ReadPB (simple_codec.go:229-248)
 *   row build       ProtoLogToL7FlowLog  flow_log/log_data/l7_flow_log.go:638-644 (l7_row.cuh)
 *   column pack     L7FlowLog.AppendToColumnBlock l7_column_block.go:167-208,376-422 (fixed-width part)
 *   counters        updateCounter decoder.go:498-525
 * What follows on the same stream: scans -> l7_strings.cu (string / array columns, trace_id_index)
 * -> l7_kg.cu (SmartEncoding) -> l7.cu (flow_tag).
 */
#include "l7_dev.h"

namespace dfgpu {

namespace {

constexpr int A_THREADS = 128;

// the thread's column of the shared-memory words + its lane of the warp's reference scratch
struct SlotSink {
  uint32_t* w;               // s_words + tid, stride A_THREADS
  unsigned long long* refs;  // scratch + lane, stride 32
  unsigned long long* elems;  // this record's block of element references (global, L7E_COUNT entries)
  uint32_t mask;             // references present
  __device__ __forceinline__ void u32(uint32_t s, uint32_t v) { w[s * A_THREADS] = v; }
  __device__ __forceinline__ void u64(uint32_t s, uint32_t lo, uint32_t hi) {
    w[s * A_THREADS] = lo;
    w[(s + 1) * A_THREADS] = hi;
  }
  __device__ __forceinline__ void add(uint32_t s, uint32_t v) { w[s * A_THREADS] += v; }
  __device__ __forceinline__ uint32_t word(uint32_t s) const { return w[s * A_THREADS]; }
  __device__ __forceinline__ void ref(uint32_t r, uint32_t off, uint32_t len) {
    refs[r * 32] = (unsigned long long)len << 32 | off;
    mask |= 1u << r;
  }
  __device__ __forceinline__ void velem(uint32_t, uint32_t, uint32_t, uint32_t) {}
  __device__ __forceinline__ void elem(uint32_t base, uint32_t i, uint32_t off, uint32_t len) {
    if (i < (base < L7E_METRICS_NAMES ? L7E_CAP_ATTR : base < L7E_METRICS_VALUES ? L7E_CAP_METRICS : 1u))
      elems[base + i] = (unsigned long long)len << 32 | off;
  }
};

struct SlotView {
  const uint32_t* wp;
  const unsigned long long* refs;
  uint32_t mask;
  __device__ __forceinline__ uint32_t w(int s) const { return wp[s * A_THREADS]; }
  __device__ __forceinline__ uint64_t w64(int s) const { return (uint64_t)w(s + 1) << 32 | w(s); }
  __device__ __forceinline__ SRef32 ref(int r) const {
    unsigned long long v = 0;
    if ((mask >> r) & 1u) v = __ldcg(refs + r * 32);
    return SRef32{(uint32_t)v, (uint32_t)(v >> 32)};
  }
};

// Row sink of the decode kernel: fixed-width values go straight to their column (lane = row: coalesced),
// references to the per-record tables, sizes into per-warp totals.
struct RowEmit {
  const BatchParams& p;
  const L7Dev* dv;
  const uint64_t* s_fix;  // shared copies of DynState::fix_off / aux_off
  const uint64_t* s_aux;
  bool ok, in_range;
  uint32_t rec, row, wtile, lane, nat, meta;
  template <class T>
  __device__ __forceinline__ void fix(int col, T v) {
    if (ok) reinterpret_cast<T*>(p.arena + s_fix[col])[row] = v;
  }
  __device__ __forceinline__ void ip6(int col, SRef32 r) {
    // ckdb.AppendIPv6 (ckdb.go:46-56): nil -> zeros, short -> zero padded
    uint64_t lo = 0, hi = 0;
    const uint32_t n = r.len > 16 ? 16 : r.len;
    for (uint32_t i = 0; i < n; i++) {
      const uint64_t v = (uint64_t)__ldg(p.in + r.off + i) << ((i & 7) * 8);
      if (i < 8) lo |= v; else hi |= v;
    }
    if (ok) reinterpret_cast<ulonglong2*>(p.arena + s_fix[col])[row] = make_ulonglong2(lo, hi);
  }
  template <class T>
  __device__ __forceinline__ void nullable(int col, bool has, T v) {
    if (ok) {
      reinterpret_cast<T*>(p.arena + s_fix[col])[row] = has ? v : T(0);
      (p.arena + s_aux[col])[row] = has ? 0 : 1;
    }
  }
  __device__ __forceinline__ void put_ref(int idx, SRef32 r) {
    if (in_range) dv->refs[(size_t)idx * dv->rec_cap + rec] = ok ? ((unsigned long long)r.len << 32 | r.off) : 0ull;
  }
  __device__ __forceinline__ void str(int i, SRef32 r) {
    put_ref(i, r);
    const uint32_t n = ok ? r.len + (nat ? tk_vsize(r.len) : 0u) : 0u;
    const uint32_t tot = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) p.warp_tot[(size_t)(L7S_STR0 + i) * p.warp_cap + wtile] = tot;
  }
  __device__ __forceinline__ void aux(int k, SRef32 r) { put_ref(L7REF_EXT_RANGE + k, r); }
  __device__ __forceinline__ void flags(bool has_ext, bool slow) {
    meta |= (has_ext ? L7META_HAS_EXT : 0u) | (slow ? L7META_SLOW_ARRAYS : 0u);
  }
  __device__ __forceinline__ void kg(uint32_t vtap, uint32_t tap_side, uint32_t proto, uint32_t port_dst, bool v6,
                                     bool vip0, bool vip1, uint32_t ip0, uint32_t ip1, uint32_t epc0, uint32_t epc1,
                                     uint64_t mac0, uint64_t mac1, uint32_t gpid0, uint32_t gpid1, uint32_t pod0,
                                     uint32_t pod1, uint32_t pid0, uint32_t pid1, SRef32 ip6s, SRef32 ip6d) {
    put_ref(L7REF_IP6_SRC, ip6s);
    put_ref(L7REF_IP6_DST, ip6d);
    if (!in_range) return;
    uint32_t* k = dv->kgin + rec;
    const size_t st = dv->rec_cap;
    k[L7KG_VTAP_SIDE_PROTO * st] = (vtap & 0xffffu) | (tap_side & 0xffu) << 16 | (proto & 0xffu) << 24;
    k[L7KG_PORT_FLAGS * st] = (port_dst & 0xffffu) | (v6 ? 1u << 16 : 0u) | (vip0 ? 1u << 17 : 0u) | (vip1 ? 1u << 18 : 0u);
    k[L7KG_IP_SRC * st] = ip0;
    k[L7KG_IP_DST * st] = ip1;
    k[L7KG_EPC_SRC * st] = epc0;
    k[L7KG_EPC_DST * st] = epc1;
    k[L7KG_MAC_SRC_LO * st] = (uint32_t)mac0;
    k[L7KG_MAC_SRC_HI * st] = (uint32_t)(mac0 >> 32);
    k[L7KG_MAC_DST_LO * st] = (uint32_t)mac1;
    k[L7KG_MAC_DST_HI * st] = (uint32_t)(mac1 >> 32);
    k[L7KG_GPID_0 * st] = gpid0;
    k[L7KG_GPID_1 * st] = gpid1;
    k[L7KG_POD_0 * st] = pod0;
    k[L7KG_POD_1 * st] = pod1;
    k[L7KG_PID_0 * st] = pid0;
    k[L7KG_PID_1 * st] = pid1;
  }
};

}  // namespace

#ifndef L7_DECODE_MIN_BLOCKS
#define L7_DECODE_MIN_BLOCKS 5
#endif
__global__ void __launch_bounds__(A_THREADS, L7_DECODE_MIN_BLOCKS) k_l7_decode(const __grid_constant__ BatchParams p) {
  __shared__ uint32_t s_tbl[M7_COUNT * L7_MAXF];
  __shared__ uint32_t s_words[W7_COUNT * A_THREADS];
  __shared__ uint64_t s_fix[L7C_COUNT], s_aux[L7C_COUNT];
  DynState* dyn = p.dyn;
  const L7Dev* dv = reinterpret_cast<const L7Dev*>(p.ext);
  for (uint32_t i = threadIdx.x; i < M7_COUNT * L7_MAXF; i += A_THREADS) s_tbl[i] = dv->tok_tbl[i];
  for (uint32_t i = threadIdx.x; i < L7C_COUNT; i += A_THREADS) {
    s_fix[i] = dyn->fix_off[i];
    s_aux[i] = dyn->aux_off[i];
  }
  __syncthreads();
  if (dyn->overflow) return;
  const uint32_t nrec = batch_nrec(p), nwarps = dyn->nwarps;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * (A_THREADS / 32) + (threadIdx.x >> 5), total_warps = gridDim.x * (A_THREADS / 32);
  const uint32_t nat = p.native_str ? 1u : 0u;
  uint32_t* my_words = s_words + threadIdx.x;
  unsigned long long* my_refs = reinterpret_cast<unsigned long long*>(dv->raw_scratch) + (size_t)gwarp * (R7_COUNT * 32) + lane;
  TokWin win;
  win.init(p.in, p.in_bytes);

  for (uint32_t wtile = gwarp; wtile < nwarps; wtile += total_warps) {
    const uint32_t rec = wtile * 32 + lane;
    const bool in_range = rec < nrec;
#pragma unroll
    for (int i = 0; i < W7_COUNT; i++) my_words[i * A_THREADS] = 0;
    SlotSink sink{my_words, my_refs, reinterpret_cast<unsigned long long*>(dv->elems) + (size_t)(in_range ? rec : 0) * L7E_COUNT, 0u};
    bool ok = false;
    uint32_t frame = 0;
    if (in_range) {
      frame = p.rec_frame[rec];
      if (!(p.repair && rec >= p.frame_bad[frame])) {
        ok = tk_decode<L7_MAXF>(win, p.rec_off[rec], p.rec_len[rec], s_tbl, sink, nat ? ~0u : 0u);
        if (!ok) {
          if (!p.repair) {
            atomicMin(&p.frame_bad[frame], rec);
            dyn->stats.need_repair = 1;
          }
          // nothing of a failed record may reach the row builder (its references are garbage)
          for (int i = 0; i < W7_COUNT; i++) my_words[i * A_THREADS] = 0;
          sink.mask = 0;
        }
      }
    }
    // AppProtoLogsData.IsValid() (datatype/pb/flow_log.go:82-89) cannot fail: Base and Base.Head are
    // pre-allocated.  A record has a row unless it (or an earlier record of its frame) failed to decode;
    // rows keep the input order: row = record index - records dropped in earlier frames.
    const uint32_t row = ok ? rec - p.frame_drop[frame] : 0xffffffffu;
    const SlotView sl{my_words, my_refs, sink.mask};
    RowEmit em{p, dv, s_fix, s_aux, ok, in_range, rec, row, wtile, lane, nat, ok ? L7META_OK : 0u};
    // genID l4_flow_log.go:949-953: the counter is pre-incremented per row in arrival order
    const L7RowInfo ri = l7_build_row(dv->tb, p.in, sl, em, nat, p.dict.analyzer_id, p.persist->id_counter + row + 1);
    if (ok) reinterpret_cast<uint16_t*>(p.arena + s_fix[L7C_TEAM_ID])[row] = (uint16_t)p.frames[frame].team_id;

    // ---- per-warp totals of everything the scans place -----------------------------------------------
#define WTOT(s, v)                                                        \
  {                                                                       \
    const uint32_t t_ = __reduce_add_sync(0xffffffffu, ok ? (v) : 0u);    \
    if (lane == 0) p.warp_tot[(size_t)(s) * p.warp_cap + wtile] = t_;     \
  }
    WTOT(L7S_ROWS, 1u)
    WTOT(L7S_AN_ELEMS, ri.z_an_e)
    WTOT(L7S_AN_BYTES, ri.z_an_b)
    WTOT(L7S_AV_ELEMS, ri.z_av_e)
    WTOT(L7S_AV_BYTES, ri.z_av_b)
    WTOT(L7S_MN_ELEMS, ri.z_mn_e)
    WTOT(L7S_MN_BYTES, ri.z_mn_b)
    WTOT(L7S_MV_ELEMS, ri.z_mv_e)
#undef WTOT
    if (in_range) {
      uint32_t* a = dv->arr + rec;
      const size_t st = dv->rec_cap;
      a[L7ARR_AN_E * st] = ok ? ri.z_an_e : 0u;
      a[L7ARR_AN_B * st] = ok ? ri.z_an_b : 0u;
      a[L7ARR_AV_E * st] = ok ? ri.z_av_e : 0u;
      a[L7ARR_AV_B * st] = ok ? ri.z_av_b : 0u;
      a[L7ARR_MN_E * st] = ok ? ri.z_mn_e : 0u;
      a[L7ARR_MN_B * st] = ok ? ri.z_mn_b : 0u;
      a[L7ARR_MV_E * st] = ok ? ri.z_mv_e : 0u;
      const uint32_t tside = ok ? ri.tside : 0u;
      dv->meta[rec] = em.meta | tside << 8;
      dv->thash[rec] = ri.trace_hash;
      dv->rec_row[rec] = row;
    }
    // lastTraceIdIndex carry (ParseTraceIdIndex :296-310): last record of the warp carrying a trace id
    {
      const bool has_trace = ok && ri.trace_len != 0;
      const uint32_t trace_mask = __ballot_sync(0xffffffffu, has_trace);
      const int last = trace_mask ? 31 - __clz(trace_mask) : 0;
      const uint64_t h = __shfl_sync(0xffffffffu, ri.trace_hash, last);
      if (lane == 0) {
        p.warp_tot[(size_t)L7S_TRACE_HAS * p.warp_cap + wtile] = trace_mask ? 1 : 0;
        p.warp_tot[(size_t)L7S_TRACE_LO * p.warp_cap + wtile] = (uint32_t)h;
        p.warp_tot[(size_t)L7S_TRACE_HI * p.warp_cap + wtile] = (uint32_t)(h >> 32);
      }
    }
    lc_note(dyn->stats.lc_first[LC7_TAPSIDE], ok ? ri.tside : 0u, rec, ok);
    // updateCounter decoder.go:498-525 (one atomic per class per warp)
    {
      int cls = -1;
      if (ok) {
        switch (ri.l7_protocol) {
          case 20: case 21: cls = 0; break;
          case 120: cls = 1; break;
          case 60: case 61: cls = 2; break;
          case 80: cls = 3; break;
          case 40: cls = 4; break;
          case 101: cls = 5; break;
        }
      }
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const uint32_t m = __ballot_sync(0xffffffffu, cls == c);
        if (lane == 0 && m) atomicAdd(&dyn->stats.l7_counts[c], __popc(m));
      }
    }
  }
}

// In repair mode (a record failed to decode in the first attempt) the frames' cut points are known:
// rows dropped before each frame (exclusive scan) and the row count the layout is planned for.
__global__ void __launch_bounds__(1024) k_l7_repair_plan(const BatchParams p) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < p.nframes; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    uint32_t v = 0;
    if (i < p.nframes) {
      const uint32_t n = p.frame_nrec[i], b = p.frame_base[i], bad = p.frame_bad[i];
      const uint32_t kept = bad == NO_BAD ? n : min(n, bad - b);
      v = n - kept;
    }
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
      const uint32_t t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < p.nframes) p.frame_drop[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += sh[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t nrec = p.dyn->stats.nrec;
    p.dyn->plan_rows = nrec > carry ? nrec - carry : 0;
    p.dyn->plan_valid = 1;
  }
}

// "last valid" carry of lastTraceIdIndex across warps (single CTA): the carry INTO warp i is the hash of
// the last warp j < i that saw a trace id, else the value persisted by the previous batch.  Chunk per
// thread + one block-wide max-scan of "last valid warp of my chunk" (same shape as k_slot_scan).
__global__ void __launch_bounds__(1024) k_l7_trace_scan(const BatchParams p) {
  __shared__ int wmax[32];
  const uint32_t n = p.dyn->nwarps;
  const uint32_t* has = p.warp_tot + (size_t)L7S_TRACE_HAS * p.warp_cap;
  const uint32_t* lo_in = p.warp_tot + (size_t)L7S_TRACE_LO * p.warp_cap;
  const uint32_t* hi_in = p.warp_tot + (size_t)L7S_TRACE_HI * p.warp_cap;
  uint32_t* olo = p.warp_base + (size_t)L7S_TRACE_LO * p.warp_cap;
  uint32_t* ohi = p.warp_base + (size_t)L7S_TRACE_HI * p.warp_cap;
  const unsigned long long carry0 = p.persist->last_trace_id_index;
  const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
  int last = -1;
  for (uint32_t i = lo; i < hi; i++)
    if (has[i]) last = (int)i;
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = last;  // inclusive max-scan inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o && t > inc) inc = t;
  }
  if (lane == 31) wmax[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < (blockDim.x >> 5) ? wmax[lane] : -1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o && t > w) w = t;
    }
    const int excl = __shfl_up_sync(0xffffffffu, w, 1);
    wmax[lane] = lane ? excl : -1;  // last valid warp before this warp of threads
    if (lane == 31)
      p.dyn->stats.last_trace_id_index = w >= 0 ? (unsigned long long)hi_in[w] << 32 | lo_in[w] : carry0;
  }
  __syncthreads();
  int prev = __shfl_up_sync(0xffffffffu, inc, 1);  // last valid before my chunk, inside my warp
  if (lane == 0) prev = -1;
  int run = prev > wmax[wid] ? prev : wmax[wid];
  for (uint32_t i = lo; i < hi; i++) {
    const unsigned long long c = run >= 0 ? (unsigned long long)hi_in[run] << 32 | lo_in[run] : carry0;
    olo[i] = (uint32_t)c;
    ohi[i] = (uint32_t)(c >> 32);
    if (has[i]) run = (int)i;
  }
}

void launch_l7_repair_plan(const BatchParams& p, cudaStream_t s) { k_l7_repair_plan<<<1, 1024, 0, s>>>(p); }

void launch_l7_decode(const BatchParams& p, unsigned grid, cudaStream_t s) {
  static bool carve = false;
  if (!carve) {  // 41 KB of shared memory per CTA, 5 CTAs per SM: ask for the largest shared-memory carve-out
    cudaFuncSetAttribute(k_l7_decode, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carve = true;
  }
  k_l7_decode<<<grid, A_THREADS, 0, s>>>(p);
  k_l7_trace_scan<<<1, 1024, 0, s>>>(p);
}
unsigned l7_decode_warps_per_cta() { return A_THREADS / 32; }

}  // namespace dfgpu
