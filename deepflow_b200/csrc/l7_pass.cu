/*
 * l7_pass.cu — the L7 flow-log kernel: AppProtoLogsData record -> one l7_flow_log
 * row.  One thread per record, a warp = 32 consecutive records (coalesced column
 * stores).  The SIZE pass decodes ONCE through the table-driven interpreter
 * (pb_interp.cuh) straight into a slot-major global buffer and sizes the variable
 * columns; the EMIT pass builds the row from those slots and copies the strings.
 * Replaces, fused:
 *   decode          gogo Unmarshal via SimpleDecoder.ReadPB (simple_codec.go:229-248)
 *   row build       ProtoLogToL7FlowLog  flow_log/log_data/l7_flow_log.go:638-644,
 *                   L7FlowLog.Fill :312-342, L7Base.Fill :514-597, fillL7FlowLog :363-452,
 *                   fillTraceIds :344-360, fillExceptionDesc :463-488 (exception_desc.go),
 *                   L7Protocol.String datatype/flow.go:623-803,
 *                   ParseTraceIdIndex :296-310 -> DJBHash libs/utils/hash.go:37-43
 *   SmartEncoding   KnowledgeGraph.FillL7 :599-613, QueryProcessInfo :589-596 -> its own kernel, l7_kg.cu
 *   column pack     L7FlowLog.AppendToColumnBlock l7_column_block.go:167-208,376-422
 *   counters        updateCounter decoder.go:498-525
 */
#include "l7_dev.h"

namespace dfgpu {

namespace {

__device__ __forceinline__ SRef cs_tbl(const L7Dev* d, uint32_t base, int32_t idx, int32_t n) {
  if (idx >= 0 && idx < n) return cs_ref(d, base + (uint32_t)idx);
  return SRef{nullptr, 0};
}
// DJBHash (libs/utils/hash.go:37-43) over n bytes; the bytes come out of aligned 8-byte blocks (one load
// per 8 bytes instead of one per byte; a block is only touched when it holds a byte of the string)
__device__ __forceinline__ uint64_t djb_hash(uint64_t h, const uint8_t* p, uint32_t n) {
  if (!n) return h;
  const unsigned long long* sb = reinterpret_cast<const unsigned long long*>((uintptr_t)p & ~(uintptr_t)7);
  uint32_t o = (uint32_t)(uintptr_t)p & 7u;
  unsigned long long cur = __ldg(sb) >> (o * 8);
  for (uint32_t i = 0; i < n; i++) {
    h = ((h << 5) + h) + (uint64_t)(cur & 0xff);
    cur >>= 8;
    if (++o == 8 && i + 1 < n) {
      o = 0;
      cur = __ldg(++sb);
    }
  }
  return h;
}
__device__ bool str_eq(SRef s, const char* lit) {
  uint32_t i = 0;
  for (; lit[i]; i++)
    if (i >= s.n || ldb(s.p + i) != (uint8_t)lit[i]) return false;
  return i == s.n;
}
template <class T>
__device__ __forceinline__ void st(const BatchParams& p, int col, uint32_t row, T v) {
  reinterpret_cast<T*>(p.arena + p.fix_off[col])[row] = v;
}
__device__ __forceinline__ void st_ip6(const BatchParams& p, int col, uint32_t row, SRef s) {
  uint64_t lo = 0, hi = 0;  // bytes 0..7 / 8..15 (no byte-indexed local array: that would live in local memory)
  const uint32_t n = s.n > 16 ? 16 : s.n;
  for (uint32_t i = 0; i < n; i++) {
    const uint64_t v = (uint64_t)ldb(s.p + i) << ((i & 7) * 8);
    if (i < 8) lo |= v; else hi |= v;
  }
  reinterpret_cast<ulonglong2*>(p.arena + p.fix_off[col])[row] = make_ulonglong2(lo, hi);
}
template <class T>
__device__ __forceinline__ void st_nullable(const BatchParams& p, int col, uint32_t row, bool has, T v) {
  reinterpret_cast<T*>(p.arena + p.fix_off[col])[row] = has ? v : T(0);
  (p.arena + p.aux_off[col])[row] = has ? 0 : 1;
}
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
template <class S>
__device__ __forceinline__ SRef sref(const uint8_t* base, const S& sl, int i) {
  const uint64_t v = sl[i];
  return SRef{base + (uint32_t)v, (uint32_t)(v >> 32)};
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
      off += pb_vsize(n);
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

// The derived (post-Fill) view of one record: the scalars of L7FlowLog.Fill + fillL7FlowLog
// (l7_flow_log.go:363-452).  Scalars only — the 22 output strings are produced one at a time by
// l7_str() where they are sized / copied, so nothing here is indexed dynamically and the whole view
// stays in registers (ncu, round 1: an SRef[22] + size arrays cost a 1016-byte stack frame per thread,
// 154 MB of live frames > L2, and 2x the algorithmic DRAM traffic).
struct L7Hdr {
  bool has_req, has_resp, has_ext, has_trace, two_ids;
  uint8_t type, is_tls, is_async, is_reversed, l7_protocol, response_status;
  bool has_request_id, has_response_code, has_request_length, has_response_length;
  uint32_t request_id, sql_rows;
  int32_t response_code, request_length, response_length;
  uint32_t n_pseudo, pseudo_name_bytes, pseudo_value_bytes;
  SRef exc;       // response_exception (wire value or the protocol's description of the code)
  SRef trace_id;  // first of trace_ids, else the legacy trace_id
  uint64_t trace_hash;
};

template <class S>
__device__ __forceinline__ void l7_hdr(const L7Dev* dv, const uint8_t* base, const S& sl, L7Hdr& o, const uint32_t nat) {
  const SRef none{nullptr, 0};
  o.has_req = sl[S7_REQ_CNT] != 0;
  o.has_resp = sl[S7_RESP_CNT] != 0;
  o.has_ext = sl[S7_EXT_CNT] != 0;
  o.has_trace = sl[S7_TRACE_CNT] != 0;
  const uint32_t flags = sl_u32(sl, S7_FLAGS);
  o.type = (uint8_t)sl_u32(sl, S7_HEAD_MSG_TYPE);
  o.is_tls = (flags & 1) ? 1 : 0;
  o.is_async = (flags & 2) ? 1 : 0;
  o.is_reversed = (flags & 4) ? 1 : 0;
  o.l7_protocol = (uint8_t)sl_u32(sl, S7_HEAD_PROTO);
  o.response_status = 5;  // STATUS_UNKNOWN
  o.request_length = sl_i32(sl, S7_REQ_LEN);
  o.response_length = sl_i32(sl, S7_RESP_LEN);
  o.sql_rows = sl_u32(sl, S7_ROW_EFFECT);
  o.has_request_length = o.has_response_length = o.has_response_code = o.has_request_id = false;
  o.response_code = 0;
  o.request_id = 0;
  o.exc = none;
  if (o.has_req && o.request_length != -1 && o.type != 1) o.has_request_length = true;  // :379-387
  if (o.has_resp) {  // :389-409
    o.response_status = (uint8_t)sl_u32(sl, S7_RESP_STATUS);
    if (o.type != 0) {
      const int32_t code = sl_i32(sl, S7_RESP_CODE);
      o.response_code = code;
      SRef exc = sref(base, sl, S7_EXCEPTION);
      if (exc.n == 0 && (o.response_status == 3 || o.response_status == 4)) {  // fillExceptionDesc
        switch (o.l7_protocol) {
          case 20:
          case 21: exc = cs_tbl(dv, dv->cs_http, (int32_t)(uint16_t)code, DFGPU_EXC_HTTP_LEN); break;
          case 120: {
            uint16_t c = (uint16_t)code;
            if (c < DFGPU_EXC_DNS_LEN) exc = cs_ref(dv, dv->cs_dns + c);
            else if (c >= 22 && c <= 3840) exc = cs_ref(dv, CS_DNS_AVAILABLE);
            else if (c >= 3841 && c <= 4095) exc = cs_ref(dv, CS_DNS_PRIVATE);
            else exc = cs_ref(dv, CS_DNS_AVAILABLE);
            break;
          }
          case 40: exc = cs_tbl(dv, dv->cs_dubbo, (int32_t)(uint16_t)code, DFGPU_EXC_DUBBO_LEN); break;
          case 100: {
            int16_t c = (int16_t)code;
            if (c == -1) exc = cs_ref(dv, CS_KAFKA_UNKNOWN);
            else if (c > 0 && c < DFGPU_EXC_KAFKA_LEN) exc = cs_ref(dv, dv->cs_kafka + c);
            break;
          }
          case 101: {
            uint16_t c = (uint16_t)code;
            bool v5 = str_eq(sref(base, sl, S7_VERSION), "5");
            if (!v5) { if (c > 0 && c < DFGPU_EXC_MQTT_V3_LEN) exc = cs_ref(dv, dv->cs_mqtt3 + c); }
            else { if (c > 0 && c < DFGPU_EXC_MQTT_V5_LEN) exc = cs_ref(dv, dv->cs_mqtt5 + c); }
            break;
          }
          default: break;  // h.ResponseException = l.Resp.Exception (empty)
        }
      }
      o.exc = exc;
      if (o.response_code != -32768) o.has_response_code = true;
      if (o.response_length != -1) o.has_response_length = true;
    }
  }
  o.n_pseudo = o.pseudo_name_bytes = o.pseudo_value_bytes = 0;
  if (o.has_ext) {  // :411-433
    o.request_id = sl_u32(sl, S7_REQUEST_ID);
    if (o.request_id != 0) o.has_request_id = true;
    uint32_t n;
    // nat: every element carries its uvarint length in the stream (names are < 128 bytes: one byte)
    const uint32_t pv = nat ? ~0u : 0u;
    if ((n = sl_len(sl, S7_USER_AGENT))) { o.n_pseudo++; o.pseudo_name_bytes += 15 + nat; o.pseudo_value_bytes += n + (pb_vsize(n) & pv); }
    if ((n = sl_len(sl, S7_REFERER))) { o.n_pseudo++; o.pseudo_name_bytes += 12 + nat; o.pseudo_value_bytes += n + (pb_vsize(n) & pv); }
    if ((n = sl_len(sl, S7_RPC_SERVICE))) { o.n_pseudo++; o.pseudo_name_bytes += 11 + nat; o.pseudo_value_bytes += n + (pb_vsize(n) & pv); }
  }
  SRef trace_id = none;
  o.two_ids = false;
  if (o.has_trace) {  // :434-438 + fillTraceIds :344-360
    const uint32_t n_ids = (uint32_t)sl[S7_TRACE_IDS_N];
    if (n_ids > 0) trace_id = sref(base, sl, S7_TRACE_IDS_0);
    o.two_ids = n_ids > 1;
    if (trace_id.n == 0) trace_id = sref(base, sl, S7_TRACE_ID);
  }
  o.trace_id = trace_id;
  // the SIZE pass hashes the trace id (ParseTraceIdIndex :296-310) and leaves the value for the EMIT pass
  o.trace_hash = trace_id.n == 0 ? 0 : S::kHashSaved ? sl.saved_hash() : djb_hash(17, trace_id.p, trace_id.n);
  switch (o.l7_protocol) {  // :438-451
    case 100:
      if (o.has_req) {
        SRef rt = sref(base, sl, S7_REQ_TYPE);
        bool supported = str_eq(rt, "Fetch") || str_eq(rt, "Produce") || str_eq(rt, "JoinGroup") ||
                         str_eq(rt, "LeaveGroup") || str_eq(rt, "SyncGroup");
        if (o.response_code == 0 && !supported) {
          o.response_status = 2;
          o.has_response_code = false;
        }
        o.has_request_id = true;
      }
      break;
    case 43: o.has_request_id = true; break;
  }
}

// Output string `i` (stream order L7V_*) of the row.  `i` is a compile-time constant at every call
// site (fully unrolled loops), so the switch folds away.
template <class S>
__device__ __forceinline__ SRef l7_str(const int i, const L7Dev* dv, const uint8_t* base, const S& sl,
                                       const L7Hdr& h) {
  const SRef none{nullptr, 0};
  switch (i) {
    case L7V_BIZ_CODE: return sref(base, sl, S7_BIZ_CODE);
    case L7V_BIZ_SCENARIO: return sref(base, sl, S7_BIZ_SCENARIO);
    case L7V_BIZ_RESPONSE_CODE: return sref(base, sl, S7_BIZ_RESPONSE_CODE);
    case L7V_KNAME_0: return sref(base, sl, S7_KNAME_0);
    case L7V_KNAME_1: return sref(base, sl, S7_KNAME_1);
    case L7V_BIZ_PROTOCOL:
      if (h.has_ext && sl_len(sl, S7_PROTOCOL_STR)) return sref(base, sl, S7_PROTOCOL_STR);
      return cs_ref(dv, dv->proto_cs[h.l7_protocol][h.is_tls]);
    case L7V_VERSION: return sref(base, sl, S7_VERSION);
    case L7V_REQUEST_TYPE: return h.has_req ? sref(base, sl, S7_REQ_TYPE) : none;
    case L7V_REQUEST_DOMAIN: return h.has_req ? sref(base, sl, S7_DOMAIN) : none;
    case L7V_REQUEST_RESOURCE: return h.has_req ? sref(base, sl, S7_RESOURCE) : none;
    case L7V_ENDPOINT: return h.has_req ? sref(base, sl, S7_ENDPOINT) : none;
    case L7V_RESPONSE_EXCEPTION: return h.exc;
    case L7V_RESPONSE_RESULT: return (h.has_resp && h.type != 0) ? sref(base, sl, S7_RESULT) : none;
    case L7V_HTTP_PROXY_CLIENT: return h.has_ext ? sref(base, sl, S7_CLIENT_IP) : none;
    case L7V_X_REQUEST_ID_0: return h.has_ext ? sref(base, sl, S7_X_REQUEST_ID_0) : none;
    case L7V_X_REQUEST_ID_1: return h.has_ext ? sref(base, sl, S7_X_REQUEST_ID_1) : none;
    case L7V_TRACE_ID: return h.trace_id;
    case L7V_TRACE_ID_2: return (h.has_trace && h.two_ids) ? sref(base, sl, S7_TRACE_IDS_1) : none;
    case L7V_SPAN_ID: return h.has_trace ? sref(base, sl, S7_SPAN_ID) : none;
    case L7V_PARENT_SPAN_ID: return h.has_trace ? sref(base, sl, S7_PARENT_SPAN_ID) : none;
    case L7V_APP_SERVICE: return h.has_ext ? sref(base, sl, S7_SERVICE_NAME) : none;
    default: return none;  // app_instance: only the OTel import fills it
  }
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

// The slots of one record inside the slot-major buffer the SIZE pass wrote: the EMIT pass reads each
// slot where it needs it (one coalesced load across the warp) instead of copying all of them into a
// per-thread array (ncu: 7 GB of local-memory traffic per 625 k records with the private copy).
// Only slots the record actually carried are stored; a presence bitmap (rows S7_COUNT and S7_COUNT+1 of
// the buffer, bit 63 of the second word = the record decoded) says which — absent slots read as 0
// without touching memory, and the SIZE pass does not have to zero 760 bytes per record first.
struct SlotsG {
  const uint64_t* p;
  size_t stride;
  uint64_t m0, m1;
  __device__ __forceinline__ SlotsG(const uint64_t* col, size_t st, bool valid) : p(col), stride(st) {
    m0 = valid ? __ldg(p + (size_t)S7_COUNT * stride) : 0ull;
    m1 = valid ? __ldg(p + (size_t)(S7_COUNT + 1) * stride) : 0ull;
  }
  __device__ __forceinline__ bool ok() const { return (m1 >> 63) != 0; }
  static constexpr bool kHashSaved = true;  // row S7_COUNT + 2: DJB hash of the record's trace id
  __device__ __forceinline__ uint64_t saved_hash() const { return __ldg(p + (size_t)(S7_COUNT + 2) * stride); }
  __device__ __forceinline__ void save_hash(uint64_t) const {}
  __device__ __forceinline__ uint64_t operator[](int i) const {
    const uint64_t m = i < 64 ? m0 >> i : m1 >> (i - 64);
    return (m & 1) ? __ldg(p + (size_t)i * stride) : 0ull;
  }
};

// The same column, read-write, for the pass that fills it.
struct SlotsW {
  uint64_t* p;
  size_t stride;
  uint64_t m0 = 0, m1 = 0;
  struct Ref {
    SlotsW* s;
    int i;
    __device__ __forceinline__ operator uint64_t() const { return s->get(i); }
    __device__ __forceinline__ Ref& operator=(uint64_t v) {
      s->p[(size_t)i * s->stride] = v;
      if (i < 64) s->m0 |= 1ull << i; else s->m1 |= 1ull << (i - 64);
      return *this;
    }
    __device__ __forceinline__ Ref& operator+=(uint64_t v) { return *this = s->get(i) + v; }
  };
  __device__ __forceinline__ uint64_t get(int i) const {
    const uint64_t m = i < 64 ? m0 >> i : m1 >> (i - 64);
    return (m & 1) ? p[(size_t)i * stride] : 0ull;
  }
  __device__ __forceinline__ Ref operator[](int i) { return Ref{this, i}; }
  __device__ __forceinline__ uint64_t operator[](int i) const { return get(i); }
  __device__ __forceinline__ void clear() { m0 = m1 = 0; }
  static constexpr bool kHashSaved = false;
  __device__ __forceinline__ uint64_t saved_hash() const { return 0; }
  __device__ __forceinline__ void save_hash(uint64_t h) const { p[(size_t)(S7_COUNT + 2) * stride] = h; }
  __device__ __forceinline__ void finish(bool ok) {
    p[(size_t)S7_COUNT * stride] = m0;
    p[(size_t)(S7_COUNT + 1) * stride] = m1 | (uint64_t)(ok ? 1 : 0) << 63;
  }
};
static_assert(S7_COUNT <= 64 + 63, "presence bitmap: two words, bit 63 of the second is the ok flag");

// Everything after the decode: row build, sizes (SIZE pass) or column stores (EMIT pass).
// S is SlotsW (SIZE) or SlotsG (EMIT).  Called convergently by every lane of the warp (warp scans);
// lanes without a row carry ok = false and contribute zero sizes.
template <int EMIT, class S>
__device__ __forceinline__ void l7_tail(const BatchParams& p, const L7Dev* dv, DynState* dyn, const uint32_t rec,
                                        const uint32_t warp, const uint32_t lane, const bool ok, const uint32_t frame,
                                        const uint8_t* base, const uint32_t len, const S& sl) {
  // AppProtoLogsData.IsValid() (datatype/pb/flow_log.go:82-89) cannot fail: Base and
  // Base.Head are pre-allocated.
  const uint32_t nat = p.native_str ? 1u : 0u;  // ClickHouse native String encoding (dfgpu_cfg.native_strings)
  L7Hdr w;
  l7_hdr(dv, base, sl, w, nat);

  const uint32_t n_an = (uint32_t)sl[S7_ATTR_NAMES], b_an = (uint32_t)(sl[S7_ATTR_NAMES] >> 32);
  const uint32_t n_av = (uint32_t)sl[S7_ATTR_VALUES], b_av = (uint32_t)(sl[S7_ATTR_VALUES] >> 32);
  const uint32_t n_mn = (uint32_t)sl[S7_METRICS_NAMES], b_mn = (uint32_t)(sl[S7_METRICS_NAMES] >> 32);
  const uint32_t n_mv = (uint32_t)sl[S7_METRICS_VALUES];
  const uint32_t z_rows = ok ? 1u : 0u;
  const uint32_t z_an_e = ok ? w.n_pseudo + n_an : 0u, z_an_b = ok ? w.pseudo_name_bytes + b_an : 0u;
  const uint32_t z_av_e = ok ? w.n_pseudo + n_av : 0u, z_av_b = ok ? w.pseudo_value_bytes + b_av : 0u;
  const uint32_t z_mn_e = ok ? n_mn : 0u, z_mn_b = ok ? b_mn : 0u, z_mv_e = ok ? n_mv : 0u;
  const uint32_t tap_side = sl_u32(sl, S7_TAP_SIDE);
  const uint32_t tside = ok ? tapside_code(tap_side) : 0;
  const bool has_trace = ok && w.trace_id.n != 0;
  const uint32_t trace_mask = __ballot_sync(0xffffffffu, has_trace);
  const bool wr = lane == 0 && warp < p.nwarps;

  if (!EMIT) {
#define WTOT(s, v)                                                 \
  {                                                                \
    const uint32_t t_ = warp_sum(v);                               \
    if (wr) p.warp_tot[(size_t)(s) * p.nwarps + warp] = t_;        \
  }
    WTOT(L7S_ROWS, z_rows)
#pragma unroll
    for (int i = 0; i < 22; i++) {
      const uint32_t sn = l7_str(i, dv, base, sl, w).n;
      const uint32_t n = ok ? sn + (nat ? pb_vsize(sn) : 0u) : 0u;
      WTOT(L7S_STR0 + i, n)
    }
    WTOT(L7S_AN_ELEMS, z_an_e)
    WTOT(L7S_AN_BYTES, z_an_b)
    WTOT(L7S_AV_ELEMS, z_av_e)
    WTOT(L7S_AV_BYTES, z_av_b)
    WTOT(L7S_MN_ELEMS, z_mn_e)
    WTOT(L7S_MN_BYTES, z_mn_b)
    WTOT(L7S_MV_ELEMS, z_mv_e)
#undef WTOT
    // last record of the warp carrying a trace id: feeds the lastTraceIdIndex carry
    int last = trace_mask ? 31 - __clz(trace_mask) : 0;
    uint64_t h = __shfl_sync(0xffffffffu, w.trace_hash, last);
    if (wr) {
      p.warp_tot[(size_t)L7S_TRACE_HAS * p.nwarps + warp] = trace_mask ? 1 : 0;
      p.warp_tot[(size_t)L7S_TRACE_LO * p.nwarps + warp] = (uint32_t)h;
      p.warp_tot[(size_t)L7S_TRACE_HI * p.nwarps + warp] = (uint32_t)(h >> 32);
    }
    lc_note(dyn->stats.lc_first[LC7_TAPSIDE], tside, rec, ok);
    if (has_trace) sl.save_hash(w.trace_hash);
    return;
  }

  // ---- EMIT ---------------------------------------------------------------------------
  const uint32_t wb = warp < p.nwarps ? warp : 0;
#define WBASE(s, e) ((uint64_t)p.warp_base[(size_t)(s) * p.nwarps + wb] + (e))
  const uint32_t row = (uint32_t)WBASE(L7S_ROWS, warp_excl_scan(z_rows));
  // trace_id_index (ParseTraceIdIndex :296-310): hash, or previous valid hash + 1
  uint64_t trace_index = 0;
  {
    uint64_t carry = (uint64_t)p.warp_base[(size_t)L7S_TRACE_HI * p.nwarps + wb] << 32 |
                     p.warp_base[(size_t)L7S_TRACE_LO * p.nwarps + wb];
    uint32_t before = trace_mask & ((1u << lane) - 1);
    int src = before ? 31 - __clz(before) : 0;
    uint64_t prev = __shfl_sync(0xffffffffu, w.trace_hash, src);
    if (!before) prev = carry;
    trace_index = p.cfg_trace_disabled ? 0 : (w.trace_id.n != 0 ? w.trace_hash : prev + 1);
  }
  // updateCounter decoder.go:498-525 (one atomic per class per warp)
  {
    int cls = -1;
    if (ok) {
      switch (w.l7_protocol) {
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
      uint32_t m = __ballot_sync(0xffffffffu, cls == c);
      if (lane == 0 && m) atomicAdd(&dyn->stats.l7_counts[c], __popc(m));
    }
  }
  // SmartEncoding (KnowledgeGraph.FillL7: the 39 KG columns, team_id, gprocess_id_0/1) runs right after
  // this pass with one lane per (record, side): l7_kg.cu.  It needs the row of every record
  // (0xffffffff: no row, the record is skipped).
  if (rec < p.nrec) dv->rec_row[rec] = ok ? row : 0xffffffffu;

  if (ok) {
    const bool v6 = sl_u32(sl, S7_IS_IPV6) == 1;
    const uint64_t end_time = sl[S7_END_TIME];
    const uint32_t time_s = (uint32_t)(end_time / 1000000000ull);
    const uint32_t vtap_id = sl_u32(sl, S7_VTAP_ID);
    // ---- L7Base.Fill :514-582 ---------------------------------------------------------------
    const uint64_t tap_port = sl[S7_TAP_PORT];
    st<uint32_t>(p, L7C_TIME, row, time_s);
    st<uint32_t>(p, L7C_IP4_0, row, v6 ? 0u : sl_u32(sl, S7_IP_SRC));
    st<uint32_t>(p, L7C_IP4_1, row, v6 ? 0u : sl_u32(sl, S7_IP_DST));
    st_ip6(p, L7C_IP6_0, row, v6 ? sref(base, sl, S7_IP6_SRC) : SRef{nullptr, 0});
    st_ip6(p, L7C_IP6_1, row, v6 ? sref(base, sl, S7_IP6_DST) : SRef{nullptr, 0});
    st<uint8_t>(p, L7C_IS_IPV4, row, v6 ? 0 : 1);
    st<uint8_t>(p, L7C_PROTOCOL, row, (uint8_t)sl_u32(sl, S7_PROTOCOL));
    st<uint16_t>(p, L7C_CLIENT_PORT, row, (uint16_t)sl_u32(sl, S7_PORT_SRC));
    st<uint16_t>(p, L7C_SERVER_PORT, row, (uint16_t)sl_u32(sl, S7_PORT_DST));
    st<uint64_t>(p, L7C_FLOW_ID, row, sl[S7_FLOW_ID]);
    st<uint8_t>(p, L7C_CAPTURE_NETWORK_TYPE_ID, row, (uint8_t)sl_u32(sl, S7_TAP_TYPE));
    const uint8_t tap_port_type = (uint8_t)(tap_port >> 60);
    st<uint8_t>(p, L7C_NAT_SOURCE, row, (uint8_t)((tap_port >> 36) & 0xf));
    st<uint8_t>(p, L7C_CAPTURE_NIC_TYPE, row, tap_port_type);
    st<uint16_t>(p, L7C_SIGNAL_SOURCE, row, tap_port_type == 8 ? 4 : tap_port_type == 7 ? 3 : 0);
    st<uint8_t>(p, L7C_TUNNEL_TYPE, row, (uint8_t)((tap_port >> 32) & 0xf));
    st<uint32_t>(p, L7C_CAPTURE_NIC, row, (uint32_t)tap_port);
    st<uint8_t>(p, L7C_OBSERVATION_POINT, row, dyn->lc_rank[LC7_TAPSIDE][tside]);
    st<uint16_t>(p, L7C_AGENT_ID, row, (uint16_t)vtap_id);
    st<uint32_t>(p, L7C_REQ_TCP_SEQ, row, sl_u32(sl, S7_REQ_TCP_SEQ));
    st<uint32_t>(p, L7C_RESP_TCP_SEQ, row, sl_u32(sl, S7_RESP_TCP_SEQ));
    st<int64_t>(p, L7C_START_TIME, row, (int64_t)sl[S7_START_TIME] / 1000);
    st<int64_t>(p, L7C_END_TIME, row, (int64_t)end_time / 1000);
    st<uint8_t>(p, L7C_BIZ_TYPE, row, (uint8_t)sl_u32(sl, S7_BIZ_TYPE));
    st<int32_t>(p, L7C_PROCESS_ID_0, row, (int32_t)sl_u32(sl, S7_PID_0));
    st<int32_t>(p, L7C_PROCESS_ID_1, row, (int32_t)sl_u32(sl, S7_PID_1));
    st<uint64_t>(p, L7C_SYSCALL_TRACE_ID_REQUEST, row, sl[S7_SYS_REQ]);
    st<uint64_t>(p, L7C_SYSCALL_TRACE_ID_RESPONSE, row, sl[S7_SYS_RESP]);
    st<uint32_t>(p, L7C_SYSCALL_THREAD_0, row, sl_u32(sl, S7_THREAD_0));
    st<uint32_t>(p, L7C_SYSCALL_THREAD_1, row, sl_u32(sl, S7_THREAD_1));
    st<uint64_t>(p, L7C_SYSCALL_COROUTINE_0, row, sl[S7_CORO_0]);
    st<uint64_t>(p, L7C_SYSCALL_COROUTINE_1, row, sl[S7_CORO_1]);
    st<uint32_t>(p, L7C_SYSCALL_CAP_SEQ_0, row, sl_u32(sl, S7_CAP_SEQ_0));
    st<uint32_t>(p, L7C_SYSCALL_CAP_SEQ_1, row, sl_u32(sl, S7_CAP_SEQ_1));
    // ---- L7FlowLog block ----------------------------------------------------------------------
    {
      uint32_t count = p.persist->id_counter + row + 1;  // genID l4_flow_log.go:949-953
      st<uint64_t>(p, L7C__ID, row,
                   (uint64_t)time_s << 32 | (uint64_t)(p.dict.analyzer_id & 0x3ff) << 22 | ((uint64_t)count & 0x3fffff));
    }
    st<uint8_t>(p, L7C_L7_PROTOCOL, row, w.l7_protocol);
    st<uint8_t>(p, L7C_TYPE, row, w.type);
    st<uint8_t>(p, L7C_IS_TLS, row, w.is_tls);
    st<uint8_t>(p, L7C_IS_ASYNC, row, w.is_async);
    st<uint8_t>(p, L7C_IS_REVERSED, row, w.is_reversed);
    st_nullable<uint64_t>(p, L7C_REQUEST_ID, row, w.has_request_id, (uint64_t)w.request_id);
    st<uint8_t>(p, L7C_RESPONSE_STATUS, row, w.response_status);
    st_nullable<int32_t>(p, L7C_RESPONSE_CODE, row, w.has_response_code, w.response_code);
    st<uint64_t>(p, L7C_TRACE_ID_INDEX, row, trace_index);
    st_nullable<uint8_t>(p, L7C_SPAN_KIND, row, false, 0);
    st<uint64_t>(p, L7C_RESPONSE_DURATION, row, sl[S7_HEAD_RRT] / 1000);
    st_nullable<int64_t>(p, L7C_REQUEST_LENGTH, row, w.has_request_length, (int64_t)w.request_length);
    st_nullable<int64_t>(p, L7C_RESPONSE_LENGTH, row, w.has_response_length, (int64_t)w.response_length);
    st_nullable<uint64_t>(p, L7C_SQL_AFFECTED_ROWS, row, w.sql_rows != 0, (uint64_t)w.sql_rows);
    st<uint8_t>(p, L7C_DIRECTION_SCORE, row, (uint8_t)sl_u32(sl, S7_DIRECTION_SCORE));
    st<uint32_t>(p, L7C_CAPTURED_REQUEST_BYTE, row, sl_u32(sl, S7_CAP_REQ_BYTE));
    st<uint32_t>(p, L7C_CAPTURED_RESPONSE_BYTE, row, sl_u32(sl, S7_CAP_RESP_BYTE));
    // events: always empty on this path (only the OTel import fills it)
    if (nat) (p.arena + dyn->var_buf_off[L7V_EVENTS])[row] = 0;  // uvarint(0)
    else reinterpret_cast<dfgpu_pos*>(p.arena + p.aux_off[L7C_EVENTS])[row] = dfgpu_pos{0, 0};
  }

  // ---- the 22 per-row string columns: Pos + bytes (one string live at a time) ----------------
#pragma unroll
  for (int i = 0; i < 22; i++) {
    const SRef s = l7_str(i, dv, base, sl, w);
    const uint32_t vs = nat ? pb_vsize(s.n) : 0u;  // length prefix
    const uint32_t n = ok ? s.n + vs : 0u;
    const uint64_t off = WBASE(L7S_STR0 + i, warp_excl_scan(n));
    if (ok) {
      // Position of the value's bytes: the column's aux region, or (native) the device-only array the
      // flow_tag stage reads
      dfgpu_pos* pos = reinterpret_cast<dfgpu_pos*>(p.arena + (nat ? dyn->var_pos_off[i] : p.aux_off[l7_str_col(i)]));
      pos[row] = dfgpu_pos{(int64_t)(off + vs), (int64_t)(off + vs + s.n)};
      uint8_t* dst = p.arena + dyn->var_buf_off[i] + off;
      if (nat) put_varint(dst, s.n);
      if (s.n) copy_bytes(dst + vs, s.p, s.n);
    }
  }
  // ---- arrays: cumulative END offsets per row, then the elements ---------------------------
  const uint64_t an_e = WBASE(L7S_AN_ELEMS, warp_excl_scan(z_an_e)), an_b = WBASE(L7S_AN_BYTES, warp_excl_scan(z_an_b));
  const uint64_t av_e = WBASE(L7S_AV_ELEMS, warp_excl_scan(z_av_e)), av_b = WBASE(L7S_AV_BYTES, warp_excl_scan(z_av_b));
  const uint64_t mn_e = WBASE(L7S_MN_ELEMS, warp_excl_scan(z_mn_e)), mn_b = WBASE(L7S_MN_BYTES, warp_excl_scan(z_mn_b));
  const uint64_t mv_e = WBASE(L7S_MV_ELEMS, warp_excl_scan(z_mv_e));
#undef WBASE
  if (!ok) return;
  reinterpret_cast<uint64_t*>(p.arena + p.aux_off[L7C_ATTRIBUTE_NAMES])[row] = an_e + z_an_e;
  reinterpret_cast<uint64_t*>(p.arena + p.aux_off[L7C_ATTRIBUTE_VALUES])[row] = av_e + z_av_e;
  reinterpret_cast<uint64_t*>(p.arena + p.aux_off[L7C_METRICS_NAMES])[row] = mn_e + z_mn_e;
  reinterpret_cast<uint64_t*>(p.arena + p.aux_off[L7C_METRICS_VALUES])[row] = mv_e + z_mv_e;
  if (z_an_e | z_av_e | z_mn_e | z_mv_e) {
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
    // pseudo attributes first, in the reference's order (:418-429)
    if (w.has_ext) {
      SRef ua = sref(base, sl, S7_USER_AGENT), rf = sref(base, sl, S7_REFERER), rpc = sref(base, sl, S7_RPC_SERVICE);
      if (ua.n) {
        SRef nm = cs_ref(dv, CS_ATTR_UA);
        sink.attr_name(nm.p, nm.n);
        sink.attr_value(ua.p, ua.n);
      }
      if (rf.n) {
        SRef nm = cs_ref(dv, CS_ATTR_REFERER);
        sink.attr_name(nm.p, nm.n);
        sink.attr_value(rf.p, rf.n);
      }
      if (rpc.n) {
        SRef nm = cs_ref(dv, CS_ATTR_RPC);
        sink.attr_name(nm.p, nm.n);
        sink.attr_value(rpc.p, rpc.n);
      }
    }
    if (n_an | n_av | n_mn | n_mv) {
      if (sl[S7_EXT_CNT] == 1 && sl[S7_BASE_CNT] <= 1) {
        // the common case: one ext_info occurrence — walk just that range (inlined: the sink stays in
        // registers; it is the tail of the kernel)
        walk_ext_range(base + sl_off(sl, S7_EXT_RANGE), sl_len(sl, S7_EXT_RANGE), sink);
      } else {
        // ext_info split over several occurrences (merge semantics): full re-walk, out of line, on a
        // copy so that the address of `sink` itself never escapes
        L7EmitSink cold = sink;
        l7_arrays_slow(base, len, cold);
      }
    }
  }
}

#ifndef L7_PREFETCH
#define L7_PREFETCH 0  // bit 0: SIZE pass, bit 1: EMIT pass
#endif
#ifndef L7_EMIT_MIN_BLOCKS
#define L7_EMIT_MIN_BLOCKS 4
#endif
template <int EMIT>
// SIZE pass: 40 registers / 6 CTAs per SM measured fastest (0.73 ms per 625 k records; 4 and 8 CTAs: +20 % / +12 %);
// EMIT pass: 64 registers / 4 CTAs (0.67 ms; 3 / 5 / 6 CTAs per SM: 0.78 / 0.80 / 0.83 ms before the last round of
// changes).
__global__ void __launch_bounds__(256, EMIT ? L7_EMIT_MIN_BLOCKS : 6) k_l7_pass(const __grid_constant__ BatchParams p) {
  const uint32_t rec = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = rec >> 5, lane = threadIdx.x & 31;
  const bool in_range = rec < p.nrec;
  DynState* dyn = p.dyn;
  const L7Dev* dv = reinterpret_cast<const L7Dev*>(p.ext);
  if (EMIT && (dyn->stats.need_repair && !p.repair)) return;
  if (EMIT && dyn->overflow) return;
  bool ok = false;
  uint32_t frame = 0;
  const uint8_t* base = p.in;
  uint32_t len = 0;
  if (in_range) {
    frame = p.rec_frame[rec];
    base = p.in + p.rec_off[rec];
    len = p.rec_len[rec];
  }
#if L7_PREFETCH
  // all 128-byte lines of the record are requested before the walk touches the first: the walk is a
  // chain of dependent reads, this turns all but the first line's miss into an L2 hit
  if (L7_PREFETCH & (EMIT ? 2 : 1))
    for (uint32_t x = 0; x < len && x < 2048u; x += 128u)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + x));
#endif
  if (EMIT) {
    // decoded once, by the SIZE pass
    const SlotsG sl(dv->slotbuf + rec, dv->slot_stride, in_range);
    ok = sl.ok();
    l7_tail<1>(p, dv, dyn, rec, warp, lane, ok, frame, base, len, sl);
    return;
  }
  __shared__ uint32_t s_tbl[M7_COUNT * L7_MAXF];
  for (uint32_t i = threadIdx.x; i < M7_COUNT * L7_MAXF; i += blockDim.x) s_tbl[i] = dv->pb_tbl[i];
  __syncthreads();
  // The interpreter writes its slots straight into the slot-major buffer the EMIT pass reads
  // (same interleaving as a per-thread local array, without the zero / fill / copy-out round trip).
  // Threads past the last record own a spare column too (slot_stride >= nrec + 4096 > grid size).
  SlotsW sl{dv->slotbuf + rec, dv->slot_stride};
  if (in_range) {
    if (!(p.repair && rec >= p.frame_bad[frame])) {
      ok = pb_interp(base, len, s_tbl, L7_MAXF, sl, p.native_str ? ~0u : 0u);
      if (!ok) {
        if (!p.repair) {
          atomicMin(&p.frame_bad[frame], rec);
          dyn->stats.need_repair = 1;
        }
        sl.clear();
      }
    }
    sl.finish(ok);
  }
  l7_tail<0>(p, dv, dyn, rec, warp, lane, ok, frame, base, len, sl);
}

// "last valid" carry of lastTraceIdIndex across warps (single CTA): the carry INTO warp i is the hash of
// the last warp j < i that saw a trace id, else the value persisted by the previous batch.  Chunk per
// thread + one block-wide max-scan of "last valid warp of my chunk" (same shape as k_slot_scan).
__global__ void __launch_bounds__(1024) k_l7_trace_scan(const BatchParams p) {
  __shared__ int wmax[32];
  const uint32_t n = p.nwarps;
  const uint32_t* has = p.warp_tot + (size_t)L7S_TRACE_HAS * n;
  const uint32_t* lo_in = p.warp_tot + (size_t)L7S_TRACE_LO * n;
  const uint32_t* hi_in = p.warp_tot + (size_t)L7S_TRACE_HI * n;
  uint32_t* olo = p.warp_base + (size_t)L7S_TRACE_LO * n;
  uint32_t* ohi = p.warp_base + (size_t)L7S_TRACE_HI * n;
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

void launch_l7_pass(const BatchParams& p, int emit, cudaStream_t s) {
  if (!emit) {
    if (p.nrec) k_l7_pass<0><<<(p.nrec + 255) / 256, 256, 0, s>>>(p);
    k_l7_trace_scan<<<1, 1024, 0, s>>>(p);
  } else if (p.nrec) {
    k_l7_pass<1><<<(p.nrec + 255) / 256, 256, 0, s>>>(p);  // api.cu launches l7_kg.cu right after
  }
}

}  // namespace dfgpu
