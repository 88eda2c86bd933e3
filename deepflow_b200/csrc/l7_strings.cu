/*
 * l7_strings.cu — the variable-length columns of l7_flow_log, from the references the decode kernel left:
 *   22 per-row String / LowCardinality(String) columns  (Position + bytes, or uvarint + bytes in native mode)
 *   attribute_names / attribute_values / metrics_names / metrics_values arrays (l7_flow_log.go:411-433)
 *   trace_id_index (ParseTraceIdIndex :296-310), observation_point keys, the always-empty events column
 *   (L7FlowLog.AppendToColumnBlock l7_column_block.go:376-422).
 * Work item = (warp of 32 records, column task); persistent warps stride over the items, so the 24 tasks
 * of one tile run next to each other (the tile's wire bytes are read by neighbouring warps while they sit
 * in L1/L2) and every warp has a short, uniform job: one coalesced reference load, one warp scan, 32
 * Positions stored as one 512-byte segment, then the copies.
 */
#include "l7_dev.h"

namespace dfgpu {

namespace {

constexpr int B_TASKS = 23;  // 0..21 strings, 22 trace index / observation point / events (arrays: k_l7_arrays)

// String bytes -> column buffer.  Source and destination have arbitrary alignment.  Every source byte
// comes out of an aligned 8-byte block (a block is only touched when it holds a byte of the string, and
// the first blocks are requested together); bytes are then STORED from registers.  ncu, round 1: the
// byte-by-byte load/store loop for strings < 16 B was 24 % of the EMIT pass's stall samples, the head /
// tail byte loops of longer strings another 8 %.
__device__ __forceinline__ void store_bytes(uint8_t* dst, unsigned long long w, uint32_t n) {  // n <= 8
  for (uint32_t i = 0; i < n; i++) {
    dst[i] = (uint8_t)w;
    w >>= 8;
  }
}
// uvarint length prefix of the ClickHouse native String encoding (dfgpu_cfg.native_strings)
__device__ __forceinline__ void put_varint(uint8_t* dst, uint32_t n) {
  while (n >= 0x80u) {
    *dst++ = (uint8_t)(n | 0x80u);
    n >>= 7;
  }
  *dst = (uint8_t)n;
}
__device__ __noinline__ void copy_bytes(uint8_t* dst, const uint8_t* src, uint32_t n) {  // n > 0
  const unsigned long long* sb = reinterpret_cast<const unsigned long long*>((uintptr_t)src & ~(uintptr_t)7);
  const uint32_t so = (uint32_t)(uintptr_t)src & 7u, sh = so * 8u;
  const uint32_t nblk = (so + n + 7u) >> 3;  // blocks holding source bytes (>= 1)
  const unsigned long long b0 = __ldg(sb);
  const unsigned long long b1 = nblk > 1 ? __ldg(sb + 1) : 0ull;
  const unsigned long long w0 = sh ? (b0 >> sh) | (b1 << (64 - sh)) : b0;  // source bytes 0..7
  if (n <= 16) {
    const unsigned long long b2 = nblk > 2 ? __ldg(sb + 2) : 0ull;
    const unsigned long long w1 = sh ? (b1 >> sh) | (b2 << (64 - sh)) : b1;  // source bytes 8..15
    store_bytes(dst, w0, n < 8 ? n : 8u);
    if (n > 8) store_bytes(dst + 8, w1, n - 8);
    return;
  }
  // byte-align the destination from the first word, then move 8 bytes per load/store pair
  const uint32_t head = (uint32_t)(0u - (uint32_t)(uintptr_t)dst) & 7u;
  store_bytes(dst, w0, head);
  const uint8_t* s = src + head;
  uint8_t* d = dst + head;
  const uint32_t m = n - head;
  const unsigned long long* sb2 = reinterpret_cast<const unsigned long long*>((uintptr_t)s & ~(uintptr_t)7);
  const uint32_t sh2 = ((uint32_t)(uintptr_t)s & 7u) * 8u;
  unsigned long long cur = sb2 == sb ? b0 : b1;  // head <= 7: s lies in block 0 or 1
  uint32_t k = 0;
  for (; k + 8 <= m; k += 8) {
    // the next block is touched only if it holds a valid source byte
    const bool need = sh2 != 0 || k + 8 < m;
    const unsigned long long nxt = need ? __ldg(sb2 + 1) : 0ull;
    const unsigned long long w = sh2 ? (cur >> sh2) | (nxt << (64 - sh2)) : cur;
    *reinterpret_cast<unsigned long long*>(d + k) = w;
    cur = nxt;
    sb2++;
  }
  const uint32_t r = m - k;  // 0..7 tail bytes; `cur` is the block holding source byte k
  if (r) {
    const bool need = sh2 + r * 8u > 64u;  // the tail runs into the next block
    const unsigned long long nxt = need ? __ldg(sb2 + 1) : 0ull;
    store_bytes(d + k, sh2 ? (cur >> sh2) | (nxt << (64 - sh2)) : cur, r);
  }
}
struct L7EmitSink {
  dfgpu_pos* an_pos; uint8_t* an_buf; uint64_t an_off;  // attribute_names
  dfgpu_pos* av_pos; uint8_t* av_buf; uint64_t av_off;  // attribute_values
  dfgpu_pos* mn_pos; uint8_t* mn_buf; uint64_t mn_off;  // metrics_names
  uint64_t* mv;                                          // metrics_values (f64 bit patterns)
  uint32_t nat;                                          // native strings: uvarint length before every element
  __device__ __forceinline__ void put(dfgpu_pos*& pos, uint8_t* buf, uint64_t& off, const uint8_t* s, uint32_t n) {
    if (nat) {
      put_varint(buf + off, n);
      off += tk_vsize(n);
    }
    *pos++ = dfgpu_pos{(int64_t)off, (int64_t)(off + n)};
    if (n) copy_bytes(buf + off, s, n);
    off += n;
  }
  __device__ __forceinline__ void attr_name(const uint8_t* s, uint32_t n) { put(an_pos, an_buf, an_off, s, n); }
  __device__ __forceinline__ void attr_value(const uint8_t* s, uint32_t n) { put(av_pos, av_buf, av_off, s, n); }
  __device__ __forceinline__ void metric_name(const uint8_t* s, uint32_t n) { put(mn_pos, mn_buf, mn_off, s, n); }
  __device__ __forceinline__ void metric_value(uint64_t v) { *mv++ = v; }
};

// streams fields 16..19 of ONE ExtendedInfo occurrence (already validated by pb_interp)
__device__ __forceinline__ void walk_ext_range(const uint8_t* p, uint32_t n, L7EmitSink& sink) {
  Rd r{p, 0, n};
  while (r.i < r.l) {
    uint32_t f, wt;
    if (!rd_tag(r, f, wt)) return;
    if (wt == 2) {
      uint32_t len;
      if (!rd_len(r, len)) return;
      const uint8_t* s = r.p + r.i;
      if (f == 16) sink.attr_name(s, len);
      else if (f == 17) sink.attr_value(s, len);
      else if (f == 18) sink.metric_name(s, len);
      else if (f == 19) {
        uint32_t end = r.i + len;
        while (r.i < end && r.l - r.i >= 8) {
          uint64_t v = 0;
          for (int k = 0; k < 8; k++) v |= (uint64_t)ldb(r.p + r.i + k) << (8 * k);
          sink.metric_value(v);
          r.i += 8;
        }
        continue;
      }
      r.i += len;
    } else if (wt == 1 && f == 19) {
      if (r.l - r.i < 8) return;
      uint64_t v = 0;
      for (int k = 0; k < 8; k++) v |= (uint64_t)ldb(r.p + r.i + k) << (8 * k);
      sink.metric_value(v);
      r.i += 8;
    } else {
      if (!rd_skip(r, wt)) return;
    }
  }
}

// rare: ext_info split over several occurrences.  Kept out of line so that the whole switch-based
// decoder does not sit in the hot kernel's instruction stream (ncu: 31 % of the EMIT pass's stall
// samples were instruction-fetch misses on a 230 KB kernel).
__device__ __noinline__ void l7_arrays_slow(const uint8_t* base, uint32_t len, L7EmitSink& sink) {
  L7Rec scratch;
  memset(&scratch, 0, sizeof scratch);
  dec_app_proto_log(base, len, scratch, sink);
}


__device__ __forceinline__ int l7_str_col(const int i) {
  switch (i) {
    case L7V_BIZ_CODE: return L7C_BIZ_CODE;
    case L7V_BIZ_SCENARIO: return L7C_BIZ_SCENARIO;
    case L7V_BIZ_RESPONSE_CODE: return L7C_BIZ_RESPONSE_CODE;
    case L7V_KNAME_0: return L7C_PROCESS_KNAME_0;
    case L7V_KNAME_1: return L7C_PROCESS_KNAME_1;
    case L7V_BIZ_PROTOCOL: return L7C_BIZ_PROTOCOL;
    case L7V_VERSION: return L7C_VERSION;
    case L7V_REQUEST_TYPE: return L7C_REQUEST_TYPE;
    case L7V_REQUEST_DOMAIN: return L7C_REQUEST_DOMAIN;
    case L7V_REQUEST_RESOURCE: return L7C_REQUEST_RESOURCE;
    case L7V_ENDPOINT: return L7C_ENDPOINT;
    case L7V_RESPONSE_EXCEPTION: return L7C_RESPONSE_EXCEPTION;
    case L7V_RESPONSE_RESULT: return L7C_RESPONSE_RESULT;
    case L7V_HTTP_PROXY_CLIENT: return L7C_HTTP_PROXY_CLIENT;
    case L7V_X_REQUEST_ID_0: return L7C_X_REQUEST_ID_0;
    case L7V_X_REQUEST_ID_1: return L7C_X_REQUEST_ID_1;
    case L7V_TRACE_ID: return L7C_TRACE_ID;
    case L7V_TRACE_ID_2: return L7C_TRACE_ID_2;
    case L7V_SPAN_ID: return L7C_SPAN_ID;
    case L7V_PARENT_SPAN_ID: return L7C_PARENT_SPAN_ID;
    case L7V_APP_SERVICE: return L7C_APP_SERVICE;
    default: return L7C_APP_INSTANCE;
  }
}

}  // namespace

#ifndef L7_STR_MIN_BLOCKS
#define L7_STR_MIN_BLOCKS 6
#endif
#ifndef L7_ARR_MIN_BLOCKS
#define L7_ARR_MIN_BLOCKS 3
#endif
// ARRAYS = 0: the string / misc tasks (light: 40 registers);  ARRAYS = 1: the array columns, one thread per
// record walking its ext_info (heavy: its own kernel so that the light one keeps its occupancy)
template <int ARRAYS>
__global__ void __launch_bounds__(256, ARRAYS ? L7_ARR_MIN_BLOCKS : L7_STR_MIN_BLOCKS) k_l7_strings(const __grid_constant__ BatchParams p) {
  __shared__ uint8_t s_strcol[22];
  DynState* dyn = p.dyn;
  const L7Dev* dv = reinterpret_cast<const L7Dev*>(p.ext);
  if (threadIdx.x < 22) s_strcol[threadIdx.x] = (uint8_t)l7_str_col((int)threadIdx.x);
  __syncthreads();
  if ((dyn->stats.need_repair && !p.repair) || dyn->overflow) return;
  const uint32_t nrec = batch_nrec(p), nwarps = dyn->nwarps;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), total_warps = gridDim.x * (blockDim.x >> 5);
  const uint32_t nat = p.native_str ? 1u : 0u;
  const size_t cap = dv->rec_cap;
  const uint64_t nitems = (uint64_t)nwarps * (ARRAYS ? 1 : B_TASKS);
  for (uint64_t item = gwarp; item < nitems; item += total_warps) {
    __syncwarp();
    const uint32_t wtile = ARRAYS ? (uint32_t)item : (uint32_t)(item / B_TASKS);
    const uint32_t task = ARRAYS ? 0xffu : (uint32_t)(item % B_TASKS);
    const uint32_t rec = wtile * 32 + lane;
    const bool in_range = rec < nrec;
    const uint32_t row = in_range ? dv->rec_row[rec] : 0xffffffffu;
    const bool ok = row != 0xffffffffu;
#define WBASE(s, e) ((uint64_t)p.warp_base[(size_t)(s) * p.warp_cap + wtile] + (e))
    if (!ARRAYS && task < 22) {
      // ---- one per-row string column: Position + bytes ------------------------------------------------
      const unsigned long long ref = ok ? __ldg(reinterpret_cast<const unsigned long long*>(dv->refs) + (size_t)task * cap + rec) : 0ull;
      const uint32_t len = (uint32_t)(ref >> 32), roff = (uint32_t)ref;
      const uint32_t vs = nat ? tk_vsize(len) : 0u;  // length prefix
      const uint32_t n = ok ? len + vs : 0u;
      const uint64_t off = WBASE(L7S_STR0 + task, warp_excl_scan(n));
      if (ok) {
        // Position of the value's bytes: the column's aux region, or (native) the device-only array the
        // flow_tag stage reads
        dfgpu_pos* pos = reinterpret_cast<dfgpu_pos*>(p.arena + (nat ? dyn->var_pos_off[task] : dyn->aux_off[s_strcol[task]]));
        pos[row] = dfgpu_pos{(int64_t)(off + vs), (int64_t)(off + vs + len)};
        uint8_t* dst = p.arena + dyn->var_buf_off[task] + off;
        if (nat) put_varint(dst, len);
        if (len) copy_bytes(dst + vs, (roff & REF_CS) ? dv->tb.cs_blob + (roff & ~REF_CS) : p.in + roff, len);
      }
    } else if (ARRAYS) {
      // ---- arrays: cumulative END offsets per row, then the elements -------------------------------------
      const uint32_t* a = dv->arr + rec;
      const uint32_t z_an_e = ok ? a[L7ARR_AN_E * cap] : 0u, z_an_b = ok ? a[L7ARR_AN_B * cap] : 0u;
      const uint32_t z_av_e = ok ? a[L7ARR_AV_E * cap] : 0u, z_av_b = ok ? a[L7ARR_AV_B * cap] : 0u;
      const uint32_t z_mn_e = ok ? a[L7ARR_MN_E * cap] : 0u, z_mn_b = ok ? a[L7ARR_MN_B * cap] : 0u;
      const uint32_t z_mv_e = ok ? a[L7ARR_MV_E * cap] : 0u;
      const uint64_t an_e = WBASE(L7S_AN_ELEMS, warp_excl_scan(z_an_e)), an_b = WBASE(L7S_AN_BYTES, warp_excl_scan(z_an_b));
      const uint64_t av_e = WBASE(L7S_AV_ELEMS, warp_excl_scan(z_av_e)), av_b = WBASE(L7S_AV_BYTES, warp_excl_scan(z_av_b));
      const uint64_t mn_e = WBASE(L7S_MN_ELEMS, warp_excl_scan(z_mn_e)), mn_b = WBASE(L7S_MN_BYTES, warp_excl_scan(z_mn_b));
      const uint64_t mv_e = WBASE(L7S_MV_ELEMS, warp_excl_scan(z_mv_e));
      if (ok) {
        reinterpret_cast<uint64_t*>(p.arena + dyn->aux_off[L7C_ATTRIBUTE_NAMES])[row] = an_e + z_an_e;
        reinterpret_cast<uint64_t*>(p.arena + dyn->aux_off[L7C_ATTRIBUTE_VALUES])[row] = av_e + z_av_e;
        reinterpret_cast<uint64_t*>(p.arena + dyn->aux_off[L7C_METRICS_NAMES])[row] = mn_e + z_mn_e;
        reinterpret_cast<uint64_t*>(p.arena + dyn->aux_off[L7C_METRICS_VALUES])[row] = mv_e + z_mv_e;
        if (z_an_e | z_av_e | z_mn_e | z_mv_e) {
          const uint32_t meta = dv->meta[rec];
          const unsigned long long* R = reinterpret_cast<const unsigned long long*>(dv->refs) + rec;
          L7EmitSink sink;
          sink.nat = nat;
          sink.an_pos = reinterpret_cast<dfgpu_pos*>(p.arena + dyn->var_pos_off[L7V_ATTR_NAMES]) + an_e;
          sink.an_buf = p.arena + dyn->var_buf_off[L7V_ATTR_NAMES];
          sink.an_off = an_b;
          sink.av_pos = reinterpret_cast<dfgpu_pos*>(p.arena + dyn->var_pos_off[L7V_ATTR_VALUES]) + av_e;
          sink.av_buf = p.arena + dyn->var_buf_off[L7V_ATTR_VALUES];
          sink.av_off = av_b;
          sink.mn_pos = reinterpret_cast<dfgpu_pos*>(p.arena + dyn->var_pos_off[L7V_METRICS_NAMES]) + mn_e;
          sink.mn_buf = p.arena + dyn->var_buf_off[L7V_METRICS_NAMES];
          sink.mn_off = mn_b;
          sink.mv = reinterpret_cast<uint64_t*>(p.arena + dyn->var_buf_off[L7V_METRICS_VALUES]) + mv_e;
          const unsigned long long ext = R[(size_t)L7REF_EXT_RANGE * cap];
        if ((uint32_t)(ext >> 32) || !(meta & L7META_SLOW_ARRAYS)) {
          // Fast path: the decode kernel left a reference to every element (l7_tok.h L7E_*): no second walk
          // over the record.  Pseudo attributes first, in the reference's order (:418-429).
          uint32_t np = 0;
          if (meta & L7META_HAS_EXT) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const unsigned long long r = R[(size_t)(L7REF_UA + k) * cap];
              if ((uint32_t)(r >> 32)) {
                const SRef nm = cs_ref(dv, CS_ATTR_UA + k);
                sink.attr_name(nm.p, nm.n);
                sink.attr_value(p.in + (uint32_t)r, (uint32_t)(r >> 32));
                np++;
              }
            }
          }
          const unsigned long long* E = reinterpret_cast<const unsigned long long*>(dv->elems) + (size_t)rec * L7E_COUNT;
          const uint32_t n_an = z_an_e - np, n_av = z_av_e - np;
          for (uint32_t i = 0; i < n_an; i++) {
            const unsigned long long r = __ldg(E + L7E_ATTR_NAMES + i);
            sink.attr_name(p.in + (uint32_t)r, (uint32_t)(r >> 32));
          }
          for (uint32_t i = 0; i < n_av; i++) {
            const unsigned long long r = __ldg(E + L7E_ATTR_VALUES + i);
            sink.attr_value(p.in + (uint32_t)r, (uint32_t)(r >> 32));
          }
          for (uint32_t i = 0; i < z_mn_e; i++) {
            const unsigned long long r = __ldg(E + L7E_METRICS_NAMES + i);
            sink.metric_name(p.in + (uint32_t)r, (uint32_t)(r >> 32));
          }
          if (z_mv_e) {  // one packed run of doubles
            const unsigned long long r = __ldg(E + L7E_METRICS_VALUES);
            const uint8_t* q = p.in + (uint32_t)r;
            for (uint32_t i = 0; i < z_mv_e; i++) {
              uint64_t v = 0;
              for (int k = 0; k < 8; k++) v |= (uint64_t)ldb(q + 8 * i + k) << (8 * k);
              sink.metric_value(v);
            }
          }
        } else {
          // ext_info split over several occurrences / more elements than a reference block holds (merge
          // semantics): pseudo attributes, then a full re-walk, out of line, on a copy so that the address of
          // `sink` itself never escapes
          if (meta & L7META_HAS_EXT) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
              const unsigned long long r = R[(size_t)(L7REF_UA + k) * cap];
              if ((uint32_t)(r >> 32)) {
                const SRef nm = cs_ref(dv, CS_ATTR_UA + k);
                sink.attr_name(nm.p, nm.n);
                sink.attr_value(p.in + (uint32_t)r, (uint32_t)(r >> 32));
              }
            }
          }
          L7EmitSink cold = sink;
          l7_arrays_slow(p.in + p.rec_off[rec], p.rec_len[rec], cold);
        }
      }
      }
    } else {
      // ---- trace_id_index (ParseTraceIdIndex :296-310): hash, or previous valid hash + 1 ------------------
      const unsigned long long tref = ok ? __ldg(reinterpret_cast<const unsigned long long*>(dv->refs) + (size_t)L7V_TRACE_ID * cap + rec) : 0ull;
      const bool has_trace = ok && (uint32_t)(tref >> 32) != 0;
      const uint64_t my_hash = has_trace ? dv->thash[rec] : 0ull;
      const uint32_t trace_mask = __ballot_sync(0xffffffffu, has_trace);
      const uint64_t carry = (uint64_t)p.warp_base[(size_t)L7S_TRACE_HI * p.warp_cap + wtile] << 32 |
                             p.warp_base[(size_t)L7S_TRACE_LO * p.warp_cap + wtile];
      const uint32_t before = trace_mask & ((1u << lane) - 1);
      const int src = before ? 31 - __clz(before) : 0;
      uint64_t prev = __shfl_sync(0xffffffffu, my_hash, src);
      if (!before) prev = carry;
      if (ok) {
        const uint64_t trace_index = p.cfg_trace_disabled ? 0 : (has_trace ? my_hash : prev + 1);
        reinterpret_cast<uint64_t*>(p.arena + dyn->fix_off[L7C_TRACE_ID_INDEX])[row] = trace_index;
        const uint32_t tside = (dv->meta[rec] >> 8) & 0xffu;
        (p.arena + dyn->fix_off[L7C_OBSERVATION_POINT])[row] = dyn->lc_rank[LC7_TAPSIDE][tside];
        // events: always empty on this path (only the OTel import fills it)
        if (nat) (p.arena + dyn->var_buf_off[L7V_EVENTS])[row] = 0;  // uvarint(0)
        else reinterpret_cast<dfgpu_pos*>(p.arena + dyn->aux_off[L7C_EVENTS])[row] = dfgpu_pos{0, 0};
      }
    }
#undef WBASE
  }
}

void launch_l7_strings(const BatchParams& p, unsigned grid, cudaStream_t s) {
  k_l7_strings<0><<<grid, 256, 0, s>>>(p);
  k_l7_strings<1><<<grid / 2, 256, 0, s>>>(p);
}

}  // namespace dfgpu
