/*
 * l7.cu — the L7 flow-log kernels: AppProtoLogsData record -> one l7_flow_log row,
 * plus flow_tag discovery.  One thread per record (warp = 32 consecutive records).
 * Replaces, fused:
 *   decode          gogo Unmarshal via SimpleDecoder.ReadPB (simple_codec.go:229-248)
 *   row build       ProtoLogToL7FlowLog  flow_log/log_data/l7_flow_log.go:638-644,
 *                   L7FlowLog.Fill :312-342, L7Base.Fill :514-597, fillL7FlowLog :363-452,
 *                   fillTraceIds :344-360, fillExceptionDesc :463-488 (exception_desc.go),
 *                   L7Protocol.String datatype/flow.go:623-803,
 *                   ParseTraceIdIndex :296-310 -> DJBHash libs/utils/hash.go:37-43
 *   SmartEncoding   KnowledgeGraph.FillL7 :599-613 -> kg.cuh ; QueryProcessInfo :589-596
 *   column pack     L7FlowLog.AppendToColumnBlock l7_column_block.go:167-208,376-422
 *   flow_tag        GenerateNewFlowTags :648-745 with lru.Cache AddOrGet/Add semantics
 *                   (libs/lru/lru.go:44-76), FlowTag.AppendToColumnBlock
 *                   flow_tag_column_block.go:79-101, AppServiceTagWriter.Write
 *                   app_service_tag_writer.go:92-113
 *   counters        updateCounter decoder.go:498-525
 */
#include "ctx.h"
#include "l7_decode.cuh"
#include "schema_l7.h"

#include "../../include/dfgpu_exception_tables.h"

namespace dfgpu {

// ---- constant strings the path can emit (protocol names, pseudo attribute names,
// exception descriptions): one device blob, addressed by id -------------------------
enum : uint32_t {
  CS_EMPTY = 0, CS_ATTR_UA, CS_ATTR_REFERER, CS_ATTR_RPC, CS_FT_APP_SERVICE, CS_FT_ENDPOINT,
  CS_FT_APP_INSTANCE, CS_PROTO_NA, CS_PROTO_TLS, CS_DNS_AVAILABLE, CS_DNS_PRIVATE, CS_KAFKA_UNKNOWN,
  CS_FIXED_COUNT
};

struct FtEntry {  // one slot of the device tag set (32 B = one HBM sector)
  unsigned long long fp_lo, fp_hi;  // 128-bit fingerprint of the lru key; 0/0 = empty
  unsigned long long first_pos;     // (epoch << 48) | (2^48-1 - position): atomicMax keeps the FIRST
                                    // candidate of the NEWEST batch; 0 = none
  uint32_t time;                    // value stored by lru Add(): time of the last emission
  uint32_t state;                   // 0 = inserted but never emitted, 1 = emitted
};
static_assert(sizeof(FtEntry) == 32, "FtEntry");

struct L7Dev {
  const uint8_t* cs_blob;
  const uint32_t* cs_off;  // [n+1]
  uint32_t cs_http, cs_dns, cs_dubbo, cs_kafka, cs_mqtt3, cs_mqtt5;  // base ids of the tables
  uint16_t proto_cs[256][2];  // L7Protocol.String(isTLS) -> cs id
  FtEntry* ft;
  uint32_t ft_mask;
  uint32_t ft_timeout;
  uint32_t epoch;
  uint16_t org_id;
  uint8_t* ft_arena;
  uint64_t ft_arena_bytes;
};

struct L7Host {  // host mirror kept in ctx->ft
  L7Dev dev{};
  L7Dev* d_dev = nullptr;
  uint8_t* d_blob = nullptr;
  uint32_t* d_off = nullptr;
  FtEntry* d_ft = nullptr;
  uint8_t* d_ft_arena = nullptr;
  size_t ft_arena_cap = 0;
  uint32_t epoch = 0;
};

namespace {

__device__ __forceinline__ SRef cs_ref(const L7Dev* d, uint32_t id) {
  uint32_t a = d->cs_off[id], b = d->cs_off[id + 1];
  return SRef{d->cs_blob + a, b - a};
}
__device__ __forceinline__ SRef cs_tbl(const L7Dev* d, uint32_t base, int32_t idx, int32_t n) {
  if (idx >= 0 && idx < n) return cs_ref(d, base + (uint32_t)idx);
  return SRef{nullptr, 0};
}

__device__ __forceinline__ uint64_t djb_hash(uint64_t h, const uint8_t* p, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) h = ((h << 5) + h) + (uint64_t)ldb(p + i);
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
  uint32_t w[4] = {0, 0, 0, 0};
  uint8_t* b = reinterpret_cast<uint8_t*>(w);
  uint32_t n = s.n > 16 ? 16 : s.n;
  for (uint32_t i = 0; i < n; i++) b[i] = ldb(s.p + i);
  reinterpret_cast<uint4*>(p.arena + p.fix_off[col])[row] = make_uint4(w[0], w[1], w[2], w[3]);
}
template <class T>
__device__ __forceinline__ void st_nullable(const BatchParams& p, int col, uint32_t row, bool has, T v) {
  reinterpret_cast<T*>(p.arena + p.fix_off[col])[row] = has ? v : T(0);
  (p.arena + p.aux_off[col])[row] = has ? 0 : 1;
}
__device__ __forceinline__ void copy_bytes(uint8_t* dst, const uint8_t* src, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) dst[i] = ldb(src + i);
}

struct L7EmitSink {
  dfgpu_pos* an_pos; uint8_t* an_buf; uint64_t an_off;  // attribute_names
  dfgpu_pos* av_pos; uint8_t* av_buf; uint64_t av_off;  // attribute_values
  dfgpu_pos* mn_pos; uint8_t* mn_buf; uint64_t mn_off;  // metrics_names
  uint64_t* mv;                                          // metrics_values (f64 bit patterns)
  __device__ __forceinline__ void put(dfgpu_pos*& pos, uint8_t* buf, uint64_t& off, const uint8_t* s, uint32_t n) {
    *pos++ = dfgpu_pos{(int64_t)off, (int64_t)(off + n)};
    copy_bytes(buf + off, s, n);
    off += n;
  }
  __device__ __forceinline__ void attr_name(const uint8_t* s, uint32_t n) { put(an_pos, an_buf, an_off, s, n); }
  __device__ __forceinline__ void attr_value(const uint8_t* s, uint32_t n) { put(av_pos, av_buf, av_off, s, n); }
  __device__ __forceinline__ void metric_name(const uint8_t* s, uint32_t n) { put(mn_pos, mn_buf, mn_off, s, n); }
  __device__ __forceinline__ void metric_value(uint64_t v) { *mv++ = v; }
};

// The derived (post-Fill) view of one record: every output string and nullable.
struct L7Row {
  SRef str[22];  // streams 0..21
  uint8_t type, is_tls, is_async, is_reversed, l7_protocol, response_status;
  bool has_request_id, has_response_code, has_request_length, has_response_length, has_sql_rows;
  uint64_t request_id, sql_rows;
  int32_t response_code;
  int64_t request_length, response_length;
  uint32_t n_pseudo, pseudo_name_bytes, pseudo_value_bytes;
  bool trace_nonempty;
  uint64_t trace_hash;
};

// L7FlowLog.Fill + fillL7FlowLog, strings only referenced (never copied here)
__device__ void l7_derive(const L7Dev* dv, const L7Rec& r, L7Row& o) {
  const SRef none{nullptr, 0};
#pragma unroll
  for (int i = 0; i < 22; i++) o.str[i] = none;
  o.type = (uint8_t)r.head_msg_type;
  o.is_tls = (r.flags & 1) ? 1 : 0;
  o.is_async = (r.flags & 2) ? 1 : 0;
  o.is_reversed = (r.flags & 4) ? 1 : 0;
  o.l7_protocol = (uint8_t)r.head_proto;
  o.str[L7V_BIZ_CODE] = r.biz_code;
  o.str[L7V_BIZ_SCENARIO] = r.biz_scenario;
  o.str[L7V_BIZ_RESPONSE_CODE] = r.biz_response_code;
  o.str[L7V_KNAME_0] = r.process_kname_0;
  o.str[L7V_KNAME_1] = r.process_kname_1;
  if (r.has_ext && r.protocol_str.n) o.str[L7V_BIZ_PROTOCOL] = r.protocol_str;
  else o.str[L7V_BIZ_PROTOCOL] = cs_ref(dv, dv->proto_cs[o.l7_protocol][o.is_tls]);
  o.str[L7V_VERSION] = r.version;
  o.response_status = 5;  // STATUS_UNKNOWN
  o.request_length = (int64_t)r.req_len;
  o.response_length = (int64_t)r.resp_len;
  o.sql_rows = (uint64_t)r.row_effect;
  o.has_sql_rows = o.sql_rows != 0;
  o.has_request_length = o.has_response_length = o.has_response_code = o.has_request_id = false;
  o.response_code = 0;
  o.request_id = 0;
  if (r.has_req) {  // :379-387
    o.str[L7V_REQUEST_DOMAIN] = r.domain;
    o.str[L7V_REQUEST_RESOURCE] = r.resource;
    o.str[L7V_REQUEST_TYPE] = r.req_type;
    if (o.request_length != -1 && o.type != 1) o.has_request_length = true;
    o.str[L7V_ENDPOINT] = r.endpoint;
  }
  if (r.has_resp) {  // :389-409
    o.response_status = (uint8_t)r.resp_status;
    if (o.type != 0) {
      o.str[L7V_RESPONSE_RESULT] = r.result;
      o.response_code = r.resp_code;
      SRef exc = r.exception;
      if (exc.n == 0 && (o.response_status == 3 || o.response_status == 4)) {  // fillExceptionDesc
        int32_t code = r.resp_code;
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
            bool v5 = str_eq(r.version, "5");
            if (!v5) { if (c > 0 && c < DFGPU_EXC_MQTT_V3_LEN) exc = cs_ref(dv, dv->cs_mqtt3 + c); }
            else { if (c > 0 && c < DFGPU_EXC_MQTT_V5_LEN) exc = cs_ref(dv, dv->cs_mqtt5 + c); }
            break;
          }
          default: break;  // h.ResponseException = l.Resp.Exception (empty)
        }
      }
      o.str[L7V_RESPONSE_EXCEPTION] = exc;
      if (o.response_code != -32768) o.has_response_code = true;
      if (o.response_length != -1) o.has_response_length = true;
    }
  }
  o.n_pseudo = o.pseudo_name_bytes = o.pseudo_value_bytes = 0;
  if (r.has_ext) {  // :411-433
    o.request_id = (uint64_t)r.request_id;
    if (o.request_id != 0) o.has_request_id = true;
    o.str[L7V_APP_SERVICE] = r.service_name;
    o.str[L7V_X_REQUEST_ID_0] = r.x_request_id_0;
    o.str[L7V_X_REQUEST_ID_1] = r.x_request_id_1;
    o.str[L7V_HTTP_PROXY_CLIENT] = r.client_ip;
    if (r.http_user_agent.n) { o.n_pseudo++; o.pseudo_name_bytes += 15; o.pseudo_value_bytes += r.http_user_agent.n; }
    if (r.http_referer.n) { o.n_pseudo++; o.pseudo_name_bytes += 12; o.pseudo_value_bytes += r.http_referer.n; }
    if (r.rpc_service.n) { o.n_pseudo++; o.pseudo_name_bytes += 11; o.pseudo_value_bytes += r.rpc_service.n; }
  }
  SRef trace_id = none;
  if (r.has_trace) {  // :434-438 + fillTraceIds :344-360
    o.str[L7V_SPAN_ID] = r.span_id;
    if (r.n_trace_ids > 0) trace_id = r.trace_ids0;
    if (r.n_trace_ids > 1) o.str[L7V_TRACE_ID_2] = r.trace_ids1;
    if (trace_id.n == 0) trace_id = r.trace_id;
    o.str[L7V_PARENT_SPAN_ID] = r.parent_span_id;
  }
  o.str[L7V_TRACE_ID] = trace_id;
  o.trace_nonempty = trace_id.n != 0;
  o.trace_hash = o.trace_nonempty ? djb_hash(17, trace_id.p, trace_id.n) : 0;
  switch (o.l7_protocol) {  // :438-451
    case 100:
      if (r.has_req) {
        bool supported = str_eq(r.req_type, "Fetch") || str_eq(r.req_type, "Produce") || str_eq(r.req_type, "JoinGroup") ||
                         str_eq(r.req_type, "LeaveGroup") || str_eq(r.req_type, "SyncGroup");
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

}  // namespace

template <int EMIT>
__global__ void __launch_bounds__(256) k_l7_pass(const __grid_constant__ BatchParams p) {
  const uint32_t rec = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = rec >> 5, lane = threadIdx.x & 31;
  const bool in_range = rec < p.nrec;
  DynState* dyn = p.dyn;
  const L7Dev* dv = reinterpret_cast<const L7Dev*>(p.ext);
  if (EMIT && (dyn->stats.need_repair && !p.repair)) return;
  if (EMIT && dyn->overflow) return;

  L7Rec r;
  memset(&r, 0, sizeof r);
  bool ok = false;
  uint32_t frame = 0;
  const uint8_t* base = nullptr;
  uint32_t len = 0;
  if (in_range) {
    frame = p.rec_frame[rec];
    base = p.in + p.rec_off[rec];
    len = p.rec_len[rec];
    if (p.repair && rec >= p.frame_bad[frame]) {
      ok = false;
    } else {
      L7NoSink ns;
      ok = dec_app_proto_log(base, len, r, ns);
      if (!ok && !p.repair) {
        atomicMin(&p.frame_bad[frame], rec);
        dyn->stats.need_repair = 1;
      }
    }
  }
  // AppProtoLogsData.IsValid() (datatype/pb/flow_log.go:82-89) cannot fail: Base and
  // Base.Head are pre-allocated.
  L7Row w;
  l7_derive(dv, r, w);

  uint32_t sz[L7S_COUNT];
  sz[L7S_ROWS] = ok ? 1u : 0u;
#pragma unroll
  for (int i = 0; i < 22; i++) sz[L7S_STR0 + i] = ok ? w.str[i].n : 0u;
  sz[L7S_AN_ELEMS] = ok ? w.n_pseudo + r.n_attr_names : 0u;
  sz[L7S_AN_BYTES] = ok ? w.pseudo_name_bytes + r.b_attr_names : 0u;
  sz[L7S_AV_ELEMS] = ok ? w.n_pseudo + r.n_attr_values : 0u;
  sz[L7S_AV_BYTES] = ok ? w.pseudo_value_bytes + r.b_attr_values : 0u;
  sz[L7S_MN_ELEMS] = ok ? r.n_metrics_names : 0u;
  sz[L7S_MN_BYTES] = ok ? r.b_metrics_names : 0u;
  sz[L7S_MV_ELEMS] = ok ? r.n_metrics_values : 0u;
  const uint32_t tside = ok ? tapside_code(r.tap_side) : 0;
  const bool has_trace = ok && w.trace_nonempty;
  const uint32_t trace_mask = __ballot_sync(0xffffffffu, has_trace);

  if (!EMIT) {
#pragma unroll
    for (int s = 0; s < L7S_COUNT; s++) {
      uint32_t t = warp_sum(sz[s]);
      if (lane == 0 && warp < p.nwarps) p.warp_tot[(size_t)s * p.nwarps + warp] = t;
    }
    // last record of the warp carrying a trace id: feeds the lastTraceIdIndex carry
    int last = trace_mask ? 31 - __clz(trace_mask) : 0;
    uint64_t h = __shfl_sync(0xffffffffu, w.trace_hash, last);
    if (lane == 0 && warp < p.nwarps) {
      p.warp_tot[(size_t)L7S_TRACE_HAS * p.nwarps + warp] = trace_mask ? 1 : 0;
      p.warp_tot[(size_t)L7S_TRACE_LO * p.nwarps + warp] = (uint32_t)h;
      p.warp_tot[(size_t)L7S_TRACE_HI * p.nwarps + warp] = (uint32_t)(h >> 32);
    }
    lc_note(dyn->stats.lc_first[LC7_TAPSIDE], tside, rec, ok);
    return;
  }

  // ---- EMIT ---------------------------------------------------------------------------
  uint32_t ex[L7S_COUNT];
#pragma unroll
  for (int s = 0; s < L7S_COUNT; s++) ex[s] = warp_excl_scan(sz[s]);
  // trace_id_index (ParseTraceIdIndex :296-310): hash, or previous valid hash + 1
  uint64_t trace_index = 0;
  {
    const uint32_t wb0 = warp < p.nwarps ? warp : 0;
    uint64_t carry = (uint64_t)p.warp_base[(size_t)L7S_TRACE_HI * p.nwarps + wb0] << 32 |
                     p.warp_base[(size_t)L7S_TRACE_LO * p.nwarps + wb0];
    uint32_t before = trace_mask & ((1u << lane) - 1);
    int src = before ? 31 - __clz(before) : 0;
    uint64_t prev = __shfl_sync(0xffffffffu, w.trace_hash, src);
    if (!before) prev = carry;
    trace_index = p.cfg_trace_disabled ? 0 : (w.trace_nonempty ? w.trace_hash : prev + 1);
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
  if (!ok) return;
  const uint32_t wb = warp < p.nwarps ? warp : 0;
#define WBASE(s) ((uint64_t)p.warp_base[(size_t)(s) * p.nwarps + wb] + ex[s])
  const uint32_t row = (uint32_t)WBASE(L7S_ROWS);
  const dfgpu_frame fr = p.frames[frame];
  const bool v6 = r.is_ipv6 == 1;
  const uint32_t time_s = (uint32_t)(r.end_time / 1000000000ull);

  // ---- KnowledgeGraph.FillL7 :599-613 ---------------------------------------------------
  KG k;
  {
    KGInput in;
    in.v6 = v6;
    in.vip[0] = r.is_vip_interface_src == 1;
    in.vip[1] = r.is_vip_interface_dst == 1;
    in.l3_epc[0] = r.l3_epc_id_src;
    in.l3_epc[1] = r.l3_epc_id_dst;
    in.ip[0].ip4 = r.ip_src;
    in.ip[1].ip4 = r.ip_dst;
    in.ip[0].ip6len = r.ip6_src.n;
    in.ip[1].ip6len = r.ip6_dst.n;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      in.ip[0].ip6[i] = (uint32_t)i < r.ip6_src.n ? ldb(r.ip6_src.p + i) : 0;
      in.ip[1].ip6[i] = (uint32_t)i < r.ip6_dst.n ? ldb(r.ip6_dst.p + i) : 0;
    }
    in.mac[0] = r.mac_src;
    in.mac[1] = r.mac_dst;
    in.gpid[0] = r.gpid_0;
    in.gpid[1] = r.gpid_1;
    in.pod_id[0] = r.pod_id_0;
    in.pod_id[1] = r.pod_id_1;
    in.vtap_id = (uint16_t)r.vtap_id;
    in.port = (uint16_t)r.port_dst;
    in.tap_side = r.tap_side;
    in.protocol = (uint8_t)r.protocol;
    kg_fill(p.dict, in, k);
  }
  uint32_t gpid0 = r.gpid_0, gpid1 = r.gpid_1;
  if (r.process_id_0 != 0 && r.gpid_0 == 0) {  // :589-592
    gpid0 = t64_getz(p.dict, TID_VTAP_PID, (uint64_t)(uint16_t)r.vtap_id << 32 | (uint64_t)r.process_id_0);
    k.tag_source[0] |= TS_PROCESSID;
  }
  if (r.process_id_1 != 0 && r.gpid_1 == 0) {
    gpid1 = t64_getz(p.dict, TID_VTAP_PID, (uint64_t)(uint16_t)r.vtap_id << 32 | (uint64_t)r.process_id_1);
    k.tag_source[1] |= TS_PROCESSID;
  }
  st<uint16_t>(p, L7C_REGION_ID_0, row, k.region_id[0]);
  st<uint16_t>(p, L7C_REGION_ID_1, row, k.region_id[1]);
  st<uint16_t>(p, L7C_AZ_ID_0, row, k.az_id[0]);
  st<uint16_t>(p, L7C_AZ_ID_1, row, k.az_id[1]);
  st<uint16_t>(p, L7C_HOST_ID_0, row, k.host_id[0]);
  st<uint16_t>(p, L7C_HOST_ID_1, row, k.host_id[1]);
  st<uint8_t>(p, L7C_L3_DEVICE_TYPE_0, row, k.l3_device_type[0]);
  st<uint8_t>(p, L7C_L3_DEVICE_TYPE_1, row, k.l3_device_type[1]);
  st<uint32_t>(p, L7C_L3_DEVICE_ID_0, row, k.l3_device_id[0]);
  st<uint32_t>(p, L7C_L3_DEVICE_ID_1, row, k.l3_device_id[1]);
  st<uint32_t>(p, L7C_POD_NODE_ID_0, row, k.pod_node_id[0]);
  st<uint32_t>(p, L7C_POD_NODE_ID_1, row, k.pod_node_id[1]);
  st<uint16_t>(p, L7C_POD_NS_ID_0, row, k.pod_ns_id[0]);
  st<uint16_t>(p, L7C_POD_NS_ID_1, row, k.pod_ns_id[1]);
  st<uint32_t>(p, L7C_POD_GROUP_ID_0, row, k.pod_group_id[0]);
  st<uint32_t>(p, L7C_POD_GROUP_ID_1, row, k.pod_group_id[1]);
  st<uint32_t>(p, L7C_POD_ID_0, row, k.pod_id[0]);
  st<uint32_t>(p, L7C_POD_ID_1, row, k.pod_id[1]);
  st<uint16_t>(p, L7C_POD_CLUSTER_ID_0, row, k.pod_cluster_id[0]);
  st<uint16_t>(p, L7C_POD_CLUSTER_ID_1, row, k.pod_cluster_id[1]);
  st<int32_t>(p, L7C_L3_EPC_ID_0, row, k.l3_epc_id[0]);
  st<int32_t>(p, L7C_L3_EPC_ID_1, row, k.l3_epc_id[1]);
  st<int32_t>(p, L7C_EPC_ID_0, row, k.epc_id[0]);
  st<int32_t>(p, L7C_EPC_ID_1, row, k.epc_id[1]);
  st<uint16_t>(p, L7C_SUBNET_ID_0, row, k.subnet_id[0]);
  st<uint16_t>(p, L7C_SUBNET_ID_1, row, k.subnet_id[1]);
  st<uint32_t>(p, L7C_SERVICE_ID_0, row, k.service_id[0]);
  st<uint32_t>(p, L7C_SERVICE_ID_1, row, k.service_id[1]);
  st<uint32_t>(p, L7C_AUTO_INSTANCE_ID_0, row, k.auto_instance_id[0]);
  st<uint8_t>(p, L7C_AUTO_INSTANCE_TYPE_0, row, k.auto_instance_type[0]);
  st<uint32_t>(p, L7C_AUTO_SERVICE_ID_0, row, k.auto_service_id[0]);
  st<uint8_t>(p, L7C_AUTO_SERVICE_TYPE_0, row, k.auto_service_type[0]);
  st<uint32_t>(p, L7C_AUTO_INSTANCE_ID_1, row, k.auto_instance_id[1]);
  st<uint8_t>(p, L7C_AUTO_INSTANCE_TYPE_1, row, k.auto_instance_type[1]);
  st<uint32_t>(p, L7C_AUTO_SERVICE_ID_1, row, k.auto_service_id[1]);
  st<uint8_t>(p, L7C_AUTO_SERVICE_TYPE_1, row, k.auto_service_type[1]);
  st<uint8_t>(p, L7C_TAG_SOURCE_0, row, k.tag_source[0]);
  st<uint8_t>(p, L7C_TAG_SOURCE_1, row, k.tag_source[1]);
  st<uint16_t>(p, L7C_TEAM_ID, row, (uint16_t)fr.team_id);

  // ---- L7Base.Fill :514-582 ---------------------------------------------------------------
  st<uint32_t>(p, L7C_TIME, row, time_s);
  st<uint32_t>(p, L7C_IP4_0, row, v6 ? 0u : r.ip_src);
  st<uint32_t>(p, L7C_IP4_1, row, v6 ? 0u : r.ip_dst);
  st_ip6(p, L7C_IP6_0, row, v6 ? r.ip6_src : SRef{nullptr, 0});
  st_ip6(p, L7C_IP6_1, row, v6 ? r.ip6_dst : SRef{nullptr, 0});
  st<uint8_t>(p, L7C_IS_IPV4, row, v6 ? 0 : 1);
  st<uint8_t>(p, L7C_PROTOCOL, row, (uint8_t)r.protocol);
  st<uint16_t>(p, L7C_CLIENT_PORT, row, (uint16_t)r.port_src);
  st<uint16_t>(p, L7C_SERVER_PORT, row, (uint16_t)r.port_dst);
  st<uint64_t>(p, L7C_FLOW_ID, row, r.flow_id);
  st<uint8_t>(p, L7C_CAPTURE_NETWORK_TYPE_ID, row, (uint8_t)r.tap_type);
  const uint8_t tap_port_type = (uint8_t)(r.tap_port >> 60);
  st<uint8_t>(p, L7C_NAT_SOURCE, row, (uint8_t)((r.tap_port >> 36) & 0xf));
  st<uint8_t>(p, L7C_CAPTURE_NIC_TYPE, row, tap_port_type);
  st<uint16_t>(p, L7C_SIGNAL_SOURCE, row, tap_port_type == 8 ? 4 : tap_port_type == 7 ? 3 : 0);
  st<uint8_t>(p, L7C_TUNNEL_TYPE, row, (uint8_t)((r.tap_port >> 32) & 0xf));
  st<uint32_t>(p, L7C_CAPTURE_NIC, row, (uint32_t)r.tap_port);
  st<uint8_t>(p, L7C_OBSERVATION_POINT, row, dyn->lc_rank[LC7_TAPSIDE][tside]);
  st<uint16_t>(p, L7C_AGENT_ID, row, (uint16_t)r.vtap_id);
  st<uint32_t>(p, L7C_REQ_TCP_SEQ, row, r.req_tcp_seq);
  st<uint32_t>(p, L7C_RESP_TCP_SEQ, row, r.resp_tcp_seq);
  st<int64_t>(p, L7C_START_TIME, row, (int64_t)r.start_time / 1000);
  st<int64_t>(p, L7C_END_TIME, row, (int64_t)r.end_time / 1000);
  st<uint32_t>(p, L7C_GPROCESS_ID_0, row, gpid0);
  st<uint32_t>(p, L7C_GPROCESS_ID_1, row, gpid1);
  st<uint8_t>(p, L7C_BIZ_TYPE, row, (uint8_t)r.biz_type);
  st<int32_t>(p, L7C_PROCESS_ID_0, row, (int32_t)r.process_id_0);
  st<int32_t>(p, L7C_PROCESS_ID_1, row, (int32_t)r.process_id_1);
  st<uint64_t>(p, L7C_SYSCALL_TRACE_ID_REQUEST, row, r.syscall_trace_id_request);
  st<uint64_t>(p, L7C_SYSCALL_TRACE_ID_RESPONSE, row, r.syscall_trace_id_response);
  st<uint32_t>(p, L7C_SYSCALL_THREAD_0, row, r.syscall_trace_id_thread_0);
  st<uint32_t>(p, L7C_SYSCALL_THREAD_1, row, r.syscall_trace_id_thread_1);
  st<uint64_t>(p, L7C_SYSCALL_COROUTINE_0, row, r.syscall_coroutine_0);
  st<uint64_t>(p, L7C_SYSCALL_COROUTINE_1, row, r.syscall_coroutine_1);
  st<uint32_t>(p, L7C_SYSCALL_CAP_SEQ_0, row, r.syscall_cap_seq_0);
  st<uint32_t>(p, L7C_SYSCALL_CAP_SEQ_1, row, r.syscall_cap_seq_1);
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
  st_nullable<uint64_t>(p, L7C_REQUEST_ID, row, w.has_request_id, w.request_id);
  st<uint8_t>(p, L7C_RESPONSE_STATUS, row, w.response_status);
  st_nullable<int32_t>(p, L7C_RESPONSE_CODE, row, w.has_response_code, w.response_code);
  st<uint64_t>(p, L7C_TRACE_ID_INDEX, row, trace_index);
  st_nullable<uint8_t>(p, L7C_SPAN_KIND, row, false, 0);
  st<uint64_t>(p, L7C_RESPONSE_DURATION, row, r.head_rrt / 1000);
  st_nullable<int64_t>(p, L7C_REQUEST_LENGTH, row, w.has_request_length, w.request_length);
  st_nullable<int64_t>(p, L7C_RESPONSE_LENGTH, row, w.has_response_length, w.response_length);
  st_nullable<uint64_t>(p, L7C_SQL_AFFECTED_ROWS, row, w.has_sql_rows, w.sql_rows);
  st<uint8_t>(p, L7C_DIRECTION_SCORE, row, (uint8_t)r.direction_score);
  st<uint32_t>(p, L7C_CAPTURED_REQUEST_BYTE, row, r.captured_request_byte);
  st<uint32_t>(p, L7C_CAPTURED_RESPONSE_BYTE, row, r.captured_response_byte);

  // ---- the 22 per-row string columns: Pos + bytes -------------------------------------------
  {
    const int cols[22] = {L7C_BIZ_CODE, L7C_BIZ_SCENARIO, L7C_BIZ_RESPONSE_CODE, L7C_PROCESS_KNAME_0,
                          L7C_PROCESS_KNAME_1, L7C_BIZ_PROTOCOL, L7C_VERSION, L7C_REQUEST_TYPE,
                          L7C_REQUEST_DOMAIN, L7C_REQUEST_RESOURCE, L7C_ENDPOINT, L7C_RESPONSE_EXCEPTION,
                          L7C_RESPONSE_RESULT, L7C_HTTP_PROXY_CLIENT, L7C_X_REQUEST_ID_0, L7C_X_REQUEST_ID_1,
                          L7C_TRACE_ID, L7C_TRACE_ID_2, L7C_SPAN_ID, L7C_PARENT_SPAN_ID, L7C_APP_SERVICE,
                          L7C_APP_INSTANCE};
#pragma unroll
    for (int i = 0; i < 22; i++) {
      uint64_t off = WBASE(L7S_STR0 + i);
      uint32_t n = w.str[i].n;
      reinterpret_cast<dfgpu_pos*>(p.arena + p.aux_off[cols[i]])[row] = dfgpu_pos{(int64_t)off, (int64_t)(off + n)};
      if (n) copy_bytes(p.arena + dyn->var_buf_off[i] + off, w.str[i].p, n);
    }
    // events: always empty on this path (only the OTel import fills it)
    reinterpret_cast<dfgpu_pos*>(p.arena + p.aux_off[L7C_EVENTS])[row] = dfgpu_pos{0, 0};
  }
  // ---- arrays: cumulative END offsets per row, then the elements ---------------------------
  const uint64_t an_e = WBASE(L7S_AN_ELEMS), an_b = WBASE(L7S_AN_BYTES);
  const uint64_t av_e = WBASE(L7S_AV_ELEMS), av_b = WBASE(L7S_AV_BYTES);
  const uint64_t mn_e = WBASE(L7S_MN_ELEMS), mn_b = WBASE(L7S_MN_BYTES);
  const uint64_t mv_e = WBASE(L7S_MV_ELEMS);
  reinterpret_cast<uint64_t*>(p.arena + p.aux_off[L7C_ATTRIBUTE_NAMES])[row] = an_e + sz[L7S_AN_ELEMS];
  reinterpret_cast<uint64_t*>(p.arena + p.aux_off[L7C_ATTRIBUTE_VALUES])[row] = av_e + sz[L7S_AV_ELEMS];
  reinterpret_cast<uint64_t*>(p.arena + p.aux_off[L7C_METRICS_NAMES])[row] = mn_e + sz[L7S_MN_ELEMS];
  reinterpret_cast<uint64_t*>(p.arena + p.aux_off[L7C_METRICS_VALUES])[row] = mv_e + sz[L7S_MV_ELEMS];
  if (sz[L7S_AN_ELEMS] | sz[L7S_AV_ELEMS] | sz[L7S_MN_ELEMS] | sz[L7S_MV_ELEMS]) {
    L7EmitSink sink;
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
    if (r.has_ext) {
      if (r.http_user_agent.n) {
        SRef nm = cs_ref(dv, CS_ATTR_UA);
        sink.attr_name(nm.p, nm.n);
        sink.attr_value(r.http_user_agent.p, r.http_user_agent.n);
      }
      if (r.http_referer.n) {
        SRef nm = cs_ref(dv, CS_ATTR_REFERER);
        sink.attr_name(nm.p, nm.n);
        sink.attr_value(r.http_referer.p, r.http_referer.n);
      }
      if (r.rpc_service.n) {
        SRef nm = cs_ref(dv, CS_ATTR_RPC);
        sink.attr_name(nm.p, nm.n);
        sink.attr_value(r.rpc_service.p, r.rpc_service.n);
      }
    }
    if (r.n_attr_names | r.n_attr_values | r.n_metrics_names | r.n_metrics_values) {
      L7Rec scratch;
      memset(&scratch, 0, sizeof scratch);
      dec_app_proto_log(base, len, scratch, sink);  // second walk streams the repeated fields
    }
  }
#undef WBASE
}

// "last valid" carry of lastTraceIdIndex across warps (single CTA)
__global__ void k_l7_trace_scan(const BatchParams p) {
  __shared__ int sh[1024];
  __shared__ unsigned long long carry;
  const uint32_t* has = p.warp_tot + (size_t)L7S_TRACE_HAS * p.nwarps;
  const uint32_t* lo = p.warp_tot + (size_t)L7S_TRACE_LO * p.nwarps;
  const uint32_t* hi = p.warp_tot + (size_t)L7S_TRACE_HI * p.nwarps;
  uint32_t* olo = p.warp_base + (size_t)L7S_TRACE_LO * p.nwarps;
  uint32_t* ohi = p.warp_base + (size_t)L7S_TRACE_HI * p.nwarps;
  if (threadIdx.x == 0) carry = p.persist->last_trace_id_index;
  __syncthreads();
  for (uint32_t base = 0; base < p.nwarps; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    int v = (i < p.nwarps && has[i]) ? (int)threadIdx.x : -1;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
      int t = threadIdx.x >= d ? sh[threadIdx.x - d] : -1;
      __syncthreads();
      if (t > sh[threadIdx.x]) sh[threadIdx.x] = t;
      __syncthreads();
    }
    // exclusive: the carry INTO warp i is the last valid among warps < i
    int prev = threadIdx.x > 0 ? sh[threadIdx.x - 1] : -1;
    unsigned long long c = carry;
    if (prev >= 0) c = (unsigned long long)hi[base + prev] << 32 | lo[base + prev];
    if (i < p.nwarps) {
      olo[i] = (uint32_t)c;
      ohi[i] = (uint32_t)(c >> 32);
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) {
      int last = sh[threadIdx.x];
      if (last >= 0) carry = (unsigned long long)hi[base + last] << 32 | lo[base + last];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) p.dyn->stats.last_trace_id_index = carry;
}

void launch_l7_pass(const BatchParams& p, int emit, cudaStream_t s) {
  if (!emit) {
    if (p.nrec) k_l7_pass<0><<<(p.nrec + 255) / 256, 256, 0, s>>>(p);
    k_l7_trace_scan<<<1, 1024, 0, s>>>(p);
  } else if (p.nrec) {
    k_l7_pass<1><<<(p.nrec + 255) / 256, 256, 0, s>>>(p);
  }
}

void l7_streams(BatchParams& bp) {
  bp.nslots = L7S_COUNT;
  bp.nstreams = L7V_COUNT;
  bp.n_lc = 1;
  for (int v = 0; v < 22; v++) bp.streams[v] = StreamSpec{-1, (int8_t)(L7S_STR0 + v), 1, 0};
  bp.streams[L7V_ATTR_NAMES] = StreamSpec{L7S_AN_ELEMS, L7S_AN_BYTES, 1, 1};
  bp.streams[L7V_ATTR_VALUES] = StreamSpec{L7S_AV_ELEMS, L7S_AV_BYTES, 1, 1};
  bp.streams[L7V_METRICS_NAMES] = StreamSpec{L7S_MN_ELEMS, L7S_MN_BYTES, 1, 1};
  bp.streams[L7V_METRICS_VALUES] = StreamSpec{L7S_MV_ELEMS, -1, 8, 0};
  bp.streams[L7V_EVENTS] = StreamSpec{-1, -1, 0, 0};
}

int l7_lc_space(int) { return LC7_TAPSIDE; }
const std::vector<std::string>& l7_lc_names(dfgpu_ctx*, int) {
  static const std::vector<std::string> tap(kTapSideNames, kTapSideNames + 18);
  return tap;
}

// =====================================================================================
// flow_tag discovery (columnar, runs on the freshly packed l7 columns)
// =====================================================================================
namespace {

// slots of the flow_tag stage scans
enum { FT_FV_ROWS = 0, FT_FV_NAME_B = 1, FT_FV_VAL_B = 2, FT_F_ROWS = 3, FT_F_NAME_B = 4, FT_AS_ROWS = 5, FT_AS_B = 6, FT_SLOTS = 7 };
// ft_off[] entries
enum {
  FO_FV_TIME = 0, FO_FV_VPC, FO_FV_PODNS, FO_FV_FTYPE, FO_FV_FVTYPE, FO_FV_TEAM, FO_FV_TABLE, FO_FV_NAME_POS,
  FO_FV_VAL_POS, FO_FV_COUNT, FO_FV_NAME_BUF, FO_FV_VAL_BUF,
  FO_F_TIME, FO_F_VPC, FO_F_PODNS, FO_F_FTYPE, FO_F_FVTYPE, FO_F_TEAM, FO_F_TABLE, FO_F_NAME_POS, FO_F_NAME_BUF,
  FO_AS_TIME, FO_AS_TABLE, FO_AS_SVC_POS, FO_AS_INST_POS, FO_AS_TEAM, FO_AS_BUF, FO_COUNT
};

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

// 128-bit fingerprint of a lru key: two independent 64-bit mixes over the bytes
struct Fp {
  uint64_t a, b;
  __device__ __forceinline__ void init(uint64_t kind) {
    a = 0x9E3779B97F4A7C15ull ^ kind;
    b = 0xC2B2AE3D27D4EB4Full + kind * 0x165667B19E3779F9ull;
  }
  __device__ __forceinline__ void byte(uint8_t c) {
    a = (a ^ c) * 0x100000001B3ull;
    b = (b + c) * 0xFF51AFD7ED558CCDull;
    b ^= b >> 29;
  }
  __device__ __forceinline__ void bytes(const uint8_t* p, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) byte(p[i]);
    // length terminator keeps ("ab","c") and ("a","bc") apart
    byte(0xff);
    byte((uint8_t)n);
    byte((uint8_t)(n >> 8));
  }
  __device__ __forceinline__ void u64(uint64_t v) {
    for (int i = 0; i < 8; i++) byte((uint8_t)(v >> (8 * i)));
  }
  __device__ __forceinline__ U128 done() {
    U128 r{mix64(a ^ (b >> 32)), mix64(b ^ (a << 1))};
    if (r.lo == 0) r.lo = 1;  // both halves non-zero: 0 marks an empty / half-published slot
    if (r.hi == 0) r.hi = 1;
    return r;
  }
};

// find-or-insert; returns the slot or nullptr when the table is full
__device__ FtEntry* ft_upsert(const L7Dev* d, U128 fp, uint32_t* full_flag) {
  uint32_t s = (uint32_t)(fp.lo ^ (fp.hi >> 17)) & d->ft_mask;
  // a healthy table (load <= 1/8) resolves in a few probes; a long run means it is
  // saturated: give up, flag it, and let fetch report the capacity error
  for (uint32_t probes = 0; probes < 512; probes++) {
    FtEntry* e = d->ft + s;
    unsigned long long lo, hi;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(e) : "memory");
    if (lo == fp.lo && hi == fp.hi) return e;
    if ((lo | hi) == 0) {
      U128 old = cas128(e, U128{0, 0}, fp);
      if ((old.lo | old.hi) == 0 || (old.lo == fp.lo && old.hi == fp.hi)) return e;
    } else if (lo == 0 || hi == 0) {
      continue;  // torn read of a concurrent 128-bit publish: re-read this slot
    }
    s = (s + 1) & d->ft_mask;
  }
  *full_flag = 1;
  return nullptr;
}
__device__ FtEntry* ft_find(const L7Dev* d, U128 fp) {
  uint32_t s = (uint32_t)(fp.lo ^ (fp.hi >> 17)) & d->ft_mask;
  for (uint32_t probes = 0; probes < 512; probes++) {
    FtEntry* e = d->ft + s;
    if (e->fp_lo == fp.lo && e->fp_hi == fp.hi) return e;
    if ((e->fp_lo | e->fp_hi) == 0) return nullptr;
    s = (s + 1) & d->ft_mask;
  }
  return nullptr;
}

// lru semantics (l7_flow_log.go:693-701): emit when the key is new, or when
// old + CacheFlushTimeout < time (uint32 arithmetic)
__device__ __forceinline__ bool ft_candidate(const FtEntry* e, uint32_t now, uint32_t timeout) {
  if (e->state == 0) return true;
  return (uint32_t)(e->time + timeout) < now;
}

enum { KIND_FIELD_VALUE = 1, KIND_FIELD = 2, KIND_APP_SERVICE = 3 };

struct RowView {  // one packed l7 row, as the flow_tag stage sees it
  uint32_t time;  // uint32(EndTime_us / 1e6)
  uint32_t time_as;  // L7Base.Time (app service tag)
  int32_t vpc[2];
  uint16_t pod_ns[2];
  uint16_t team;
  uint32_t n_sides;
  uint64_t an_lo, an_hi, av_lo, av_hi, mn_lo, mn_hi;  // element ranges
  const dfgpu_pos *an_pos, *av_pos, *mn_pos;
  const uint8_t *an_buf, *av_buf, *mn_buf;
  SRef app_service, endpoint;
};

__device__ void load_row(const BatchParams& p, uint32_t row, RowView& v) {
  const DynState* dyn = p.dyn;
  auto fix = [&](int col) { return p.arena + p.fix_off[col]; };
  auto aux = [&](int col) { return p.arena + p.aux_off[col]; };
  int64_t end_us = reinterpret_cast<const int64_t*>(fix(L7C_END_TIME))[row];
  v.time = (uint32_t)(end_us / 1000000);
  v.time_as = reinterpret_cast<const uint32_t*>(fix(L7C_TIME))[row];
  v.vpc[0] = reinterpret_cast<const int32_t*>(fix(L7C_L3_EPC_ID_0))[row];
  v.vpc[1] = reinterpret_cast<const int32_t*>(fix(L7C_L3_EPC_ID_1))[row];
  v.pod_ns[0] = reinterpret_cast<const uint16_t*>(fix(L7C_POD_NS_ID_0))[row];
  v.pod_ns[1] = reinterpret_cast<const uint16_t*>(fix(L7C_POD_NS_ID_1))[row];
  v.team = reinterpret_cast<const uint16_t*>(fix(L7C_TEAM_ID))[row];
  v.n_sides = (v.vpc[0] == v.vpc[1] && v.pod_ns[0] == v.pod_ns[1]) ? 1 : 2;
  const uint64_t* ao = reinterpret_cast<const uint64_t*>(aux(L7C_ATTRIBUTE_NAMES));
  const uint64_t* vo = reinterpret_cast<const uint64_t*>(aux(L7C_ATTRIBUTE_VALUES));
  const uint64_t* mo = reinterpret_cast<const uint64_t*>(aux(L7C_METRICS_NAMES));
  v.an_lo = row ? ao[row - 1] : 0; v.an_hi = ao[row];
  v.av_lo = row ? vo[row - 1] : 0; v.av_hi = vo[row];
  v.mn_lo = row ? mo[row - 1] : 0; v.mn_hi = mo[row];
  v.an_pos = reinterpret_cast<const dfgpu_pos*>(p.arena + dyn->var_pos_off[L7V_ATTR_NAMES]);
  v.av_pos = reinterpret_cast<const dfgpu_pos*>(p.arena + dyn->var_pos_off[L7V_ATTR_VALUES]);
  v.mn_pos = reinterpret_cast<const dfgpu_pos*>(p.arena + dyn->var_pos_off[L7V_METRICS_NAMES]);
  v.an_buf = p.arena + dyn->var_buf_off[L7V_ATTR_NAMES];
  v.av_buf = p.arena + dyn->var_buf_off[L7V_ATTR_VALUES];
  v.mn_buf = p.arena + dyn->var_buf_off[L7V_METRICS_NAMES];
  dfgpu_pos sp = reinterpret_cast<const dfgpu_pos*>(aux(L7C_APP_SERVICE))[row];
  dfgpu_pos ep = reinterpret_cast<const dfgpu_pos*>(aux(L7C_ENDPOINT))[row];
  v.app_service = SRef{p.arena + dyn->var_buf_off[L7V_APP_SERVICE] + sp.start, (uint32_t)(sp.end - sp.start)};
  v.endpoint = SRef{p.arena + dyn->var_buf_off[L7V_ENDPOINT] + ep.start, (uint32_t)(ep.end - ep.start)};
}

// i-th (name, value) of attributeNames/attributeValues = h.Attribute* ++ the three
// extra fields (:660-663); n_names/n_values are the lengths BEFORE the append.
__device__ __forceinline__ SRef attr_name_at(const L7Dev* d, const RowView& v, uint32_t i) {
  uint32_t n = (uint32_t)(v.an_hi - v.an_lo);
  if (i < n) {
    dfgpu_pos q = v.an_pos[v.an_lo + i];
    return SRef{v.an_buf + q.start, (uint32_t)(q.end - q.start)};
  }
  return cs_ref(d, CS_FT_APP_SERVICE + (i - n));
}
__device__ __forceinline__ SRef attr_value_at(const RowView& v, uint32_t i) {
  uint32_t n = (uint32_t)(v.av_hi - v.av_lo);
  if (i < n) {
    dfgpu_pos q = v.av_pos[v.av_lo + i];
    return SRef{v.av_buf + q.start, (uint32_t)(q.end - q.start)};
  }
  uint32_t e = i - n;
  if (e == 0) return v.app_service;
  if (e == 1) return v.endpoint;
  return SRef{nullptr, 0};  // app_instance: always "" on this path
}

__device__ __forceinline__ U128 key_fp(const L7Dev* d, uint32_t kind, SRef name, SRef value, int32_t vpc,
                                       uint16_t pod_ns, uint8_t field_type, uint16_t team) {
  Fp f;
  f.init(kind);
  f.bytes(name.p, name.n);
  f.bytes(value.p, value.n);
  f.u64((uint64_t)(uint32_t)vpc << 32 | (uint64_t)pod_ns << 16 | (uint64_t)field_type << 8);
  f.u64((uint64_t)d->org_id << 16 | team);
  return f.done();
}

__device__ __forceinline__ unsigned long long pos_key(const L7Dev* d, uint32_t row, uint32_t side, uint32_t metric,
                                                      uint32_t i) {
  unsigned long long pos = (unsigned long long)(row & 0x3ffffffu) << 22 | (unsigned long long)side << 21 |
                           (unsigned long long)metric << 20 | (i & 0xfffffu);
  return (unsigned long long)(d->epoch & 0xffffu) << 48 | (0xffffffffffffull - pos);
}

// The three phases share one walk over a row's keys.
//   PHASE 0: value / app-service keys: upsert + candidate atomicMin
//   PHASE 1: decide value winners (-> counts), set their time; field + metric keys upsert + atomicMin
//   PHASE 2: decide field winners (-> counts), set their time
//   PHASE 3: write rows
template <int PHASE>
__global__ void __launch_bounds__(256) k_l7_flow_tag(const __grid_constant__ BatchParams p) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = row >> 5, lane = threadIdx.x & 31;
  DynState* dyn = p.dyn;
  const L7Dev* d = reinterpret_cast<const L7Dev*>(p.ext);
  if ((dyn->stats.need_repair && !p.repair) || dyn->overflow) return;
  if (PHASE == 3 && dyn->stats.ft_overflow) return;
  const uint32_t nrows = dyn->stats.nrows;
  const uint32_t nwarps = p.nwarps;
  const bool active = row < nrows;
  uint32_t c_fv = 0, b_fvn = 0, b_fvv = 0, c_f = 0, b_fn = 0, c_as = 0, b_as = 0;

  // PHASE 3 output cursors
  uint64_t o_fv = 0, o_fvn = 0, o_fvv = 0, o_f = 0, o_fn = 0, o_as = 0, o_asb = 0;
  if (PHASE == 3 && warp < nwarps) {
    // exclusive offsets = warp base + intra-warp exclusive scan of this thread's counts,
    // recomputed below after the counting walk; see the two-walk structure.
  }

  RowView v;
  const uint32_t timeout = d->ft_timeout;
  uint8_t* fa = d->ft_arena;
  const uint64_t* fo = dyn->stats.ft_off;

  // A PHASE-3 thread walks twice: first to count (for its offsets), then to write.
  for (int walk = 0; walk < (PHASE == 3 ? 2 : 1); walk++) {
    const bool writing = PHASE == 3 && walk == 1;
    if (PHASE == 3 && walk == 1) {
      uint32_t e_fv = warp_excl_scan(c_fv), e_fvn = warp_excl_scan(b_fvn), e_fvv = warp_excl_scan(b_fvv);
      uint32_t e_f = warp_excl_scan(c_f), e_fn = warp_excl_scan(b_fn), e_as = warp_excl_scan(c_as),
               e_asb = warp_excl_scan(b_as);
      const uint32_t wb = warp < nwarps ? warp : 0;
      o_fv = (uint64_t)p.warp_base[(size_t)FT_FV_ROWS * nwarps + wb] + e_fv;
      o_fvn = (uint64_t)p.warp_base[(size_t)FT_FV_NAME_B * nwarps + wb] + e_fvn;
      o_fvv = (uint64_t)p.warp_base[(size_t)FT_FV_VAL_B * nwarps + wb] + e_fvv;
      o_f = (uint64_t)p.warp_base[(size_t)FT_F_ROWS * nwarps + wb] + e_f;
      o_fn = (uint64_t)p.warp_base[(size_t)FT_F_NAME_B * nwarps + wb] + e_fn;
      o_as = (uint64_t)p.warp_base[(size_t)FT_AS_ROWS * nwarps + wb] + e_as;
      o_asb = (uint64_t)p.warp_base[(size_t)FT_AS_B * nwarps + wb] + e_asb;
    }
    if (!active) continue;
    if (walk == 0) load_row(p, row, v);
    const uint32_t n_names = (uint32_t)(v.an_hi - v.an_lo), n_values = (uint32_t)(v.av_hi - v.av_lo);
    const uint32_t tot_names = n_names + 3, tot_values = n_values + 3;
    const uint32_t min_len = tot_names > tot_values ? tot_values : tot_names;  // :666-674
    for (uint32_t side = 0; side < v.n_sides; side++) {
      for (uint32_t i = 0; i < min_len; i++) {
        SRef name = attr_name_at(d, v, i), value = attr_value_at(v, i);
        U128 fpv = key_fp(d, KIND_FIELD_VALUE, name, value, v.vpc[side], v.pod_ns[side], 0, v.team);
        unsigned long long pk = pos_key(d, row, side, 0, i);
        if (PHASE == 0) {
          FtEntry* e = ft_upsert(d, fpv, &dyn->stats.ft_full);
          if (e && ft_candidate(e, v.time, timeout)) atomicMax(&e->first_pos, pk);
          continue;
        }
        FtEntry* e = ft_find(d, fpv);
        if (!e || e->first_pos != pk) continue;  // not the emitting occurrence
        if (PHASE == 1) {
          e->time = v.time;   // lru Add(key, time)
          e->state = 1;
        }
        if (PHASE == 1 || (PHASE == 3 && walk == 0)) {
          c_fv++;
          b_fvn += name.n;
          b_fvv += value.n;
        }
        if (writing) {
          reinterpret_cast<uint32_t*>(fa + fo[FO_FV_TIME])[o_fv] = v.time;
          reinterpret_cast<int32_t*>(fa + fo[FO_FV_VPC])[o_fv] = v.vpc[side];
          reinterpret_cast<uint16_t*>(fa + fo[FO_FV_PODNS])[o_fv] = v.pod_ns[side];
          (fa + fo[FO_FV_FTYPE])[o_fv] = 0;   // "tag"
          (fa + fo[FO_FV_FVTYPE])[o_fv] = 0;  // "string" (:82-88: FieldType == tag)
          reinterpret_cast<uint16_t*>(fa + fo[FO_FV_TEAM])[o_fv] = v.team;
          (fa + fo[FO_FV_TABLE])[o_fv] = 0;
          reinterpret_cast<dfgpu_pos*>(fa + fo[FO_FV_NAME_POS])[o_fv] = dfgpu_pos{(int64_t)o_fvn, (int64_t)(o_fvn + name.n)};
          reinterpret_cast<dfgpu_pos*>(fa + fo[FO_FV_VAL_POS])[o_fv] = dfgpu_pos{(int64_t)o_fvv, (int64_t)(o_fvv + value.n)};
          reinterpret_cast<uint64_t*>(fa + fo[FO_FV_COUNT])[o_fv] = 1;
          for (uint32_t c = 0; c < name.n; c++) (fa + fo[FO_FV_NAME_BUF])[o_fvn + c] = name.p[c];
          for (uint32_t c = 0; c < value.n; c++) (fa + fo[FO_FV_VAL_BUF])[o_fvv + c] = value.p[c];
          o_fv++;
          o_fvn += name.n;
          o_fvv += value.n;
        }
        // the tag-only key is only consulted when the value row was emitted (:703-722)
        if (i >= n_names) continue;
        U128 fpf = key_fp(d, KIND_FIELD, name, SRef{nullptr, 0}, v.vpc[side], v.pod_ns[side], 0, v.team);
        if (PHASE == 1) {
          FtEntry* ef = ft_upsert(d, fpf, &dyn->stats.ft_full);
          if (ef && ft_candidate(ef, v.time, timeout)) atomicMax(&ef->first_pos, pk);
          continue;
        }
        FtEntry* ef = ft_find(d, fpf);
        if (!ef || ef->first_pos != pk) continue;
        if (PHASE == 2) {
          ef->time = v.time;
          ef->state = 1;
        }
        if (PHASE == 2 || (PHASE == 3 && walk == 0)) {
          c_f++;
          b_fn += name.n;
        }
        if (writing) {
          reinterpret_cast<uint32_t*>(fa + fo[FO_F_TIME])[o_f] = v.time;
          reinterpret_cast<int32_t*>(fa + fo[FO_F_VPC])[o_f] = v.vpc[side];
          reinterpret_cast<uint16_t*>(fa + fo[FO_F_PODNS])[o_f] = v.pod_ns[side];
          (fa + fo[FO_F_FTYPE])[o_f] = 0;
          (fa + fo[FO_F_FVTYPE])[o_f] = 0;
          reinterpret_cast<uint16_t*>(fa + fo[FO_F_TEAM])[o_f] = v.team;
          (fa + fo[FO_F_TABLE])[o_f] = 0;
          reinterpret_cast<dfgpu_pos*>(fa + fo[FO_F_NAME_POS])[o_f] = dfgpu_pos{(int64_t)o_fn, (int64_t)(o_fn + name.n)};
          for (uint32_t c = 0; c < name.n; c++) (fa + fo[FO_F_NAME_BUF])[o_fn + c] = name.p[c];
          o_f++;
          o_fn += name.n;
        }
      }
      // metrics names: FieldType = metrics, value "" (:726-742)
      const uint32_t n_metrics = (uint32_t)(v.mn_hi - v.mn_lo);
      for (uint32_t i = 0; i < n_metrics; i++) {
        dfgpu_pos q = v.mn_pos[v.mn_lo + i];
        SRef name{v.mn_buf + q.start, (uint32_t)(q.end - q.start)};
        U128 fpf = key_fp(d, KIND_FIELD, name, SRef{nullptr, 0}, v.vpc[side], v.pod_ns[side], 1, v.team);
        unsigned long long pk = pos_key(d, row, side, 1, i);
        if (PHASE == 0) continue;
        if (PHASE == 1) {
          FtEntry* ef = ft_upsert(d, fpf, &dyn->stats.ft_full);
          if (ef && ft_candidate(ef, v.time, timeout)) atomicMax(&ef->first_pos, pk);
          continue;
        }
        FtEntry* ef = ft_find(d, fpf);
        if (!ef || ef->first_pos != pk) continue;
        if (PHASE == 2) {
          ef->time = v.time;
          ef->state = 1;
        }
        if (PHASE == 2 || (PHASE == 3 && walk == 0)) {
          c_f++;
          b_fn += name.n;
        }
        if (writing) {
          reinterpret_cast<uint32_t*>(fa + fo[FO_F_TIME])[o_f] = v.time;
          reinterpret_cast<int32_t*>(fa + fo[FO_F_VPC])[o_f] = v.vpc[side];
          reinterpret_cast<uint16_t*>(fa + fo[FO_F_PODNS])[o_f] = v.pod_ns[side];
          (fa + fo[FO_F_FTYPE])[o_f] = 1;   // "metrics"
          (fa + fo[FO_F_FVTYPE])[o_f] = 1;  // "float": empty value and FieldType != tag
          reinterpret_cast<uint16_t*>(fa + fo[FO_F_TEAM])[o_f] = v.team;
          (fa + fo[FO_F_TABLE])[o_f] = 0;
          reinterpret_cast<dfgpu_pos*>(fa + fo[FO_F_NAME_POS])[o_f] = dfgpu_pos{(int64_t)o_fn, (int64_t)(o_fn + name.n)};
          for (uint32_t c = 0; c < name.n; c++) (fa + fo[FO_F_NAME_BUF])[o_fn + c] = name.p[c];
          o_f++;
          o_fn += name.n;
        }
      }
    }
    // app_service tag (decoder.go:464-472): AppInstance is always "" here
    if (v.app_service.n) {
      Fp f;
      f.init(KIND_APP_SERVICE);
      f.bytes(v.app_service.p, v.app_service.n);
      f.u64((uint64_t)d->org_id << 16 | v.team);
      U128 fpa = f.done();
      unsigned long long pk = pos_key(d, row, 0, 0, 0);
      if (PHASE == 0) {
        FtEntry* e = ft_upsert(d, fpa, &dyn->stats.ft_full);
        if (e && ft_candidate(e, v.time_as, timeout)) atomicMax(&e->first_pos, pk);
      } else {
        FtEntry* e = ft_find(d, fpa);
        if (e && e->first_pos == pk) {
          if (PHASE == 1) {
            e->time = v.time_as;
            e->state = 1;
          }
          if (PHASE == 1 || (PHASE == 3 && walk == 0)) {
            c_as++;
            b_as += v.app_service.n;
          }
          if (writing) {
            reinterpret_cast<uint32_t*>(fa + fo[FO_AS_TIME])[o_as] = v.time_as;
            (fa + fo[FO_AS_TABLE])[o_as] = 0;
            reinterpret_cast<dfgpu_pos*>(fa + fo[FO_AS_SVC_POS])[o_as] =
                dfgpu_pos{(int64_t)o_asb, (int64_t)(o_asb + v.app_service.n)};
            reinterpret_cast<dfgpu_pos*>(fa + fo[FO_AS_INST_POS])[o_as] = dfgpu_pos{0, 0};
            reinterpret_cast<uint16_t*>(fa + fo[FO_AS_TEAM])[o_as] = v.team;
            for (uint32_t c = 0; c < v.app_service.n; c++) (fa + fo[FO_AS_BUF])[o_asb + c] = v.app_service.p[c];
            o_as++;
            o_asb += v.app_service.n;
          }
        }
      }
    }
  }
  if (PHASE == 1 || PHASE == 2) {
    // per-warp totals of the rows/bytes decided in this phase
    uint32_t vals[FT_SLOTS] = {c_fv, b_fvn, b_fvv, c_f, b_fn, c_as, b_as};
    const int lo = PHASE == 1 ? 0 : FT_F_ROWS, hi = PHASE == 1 ? FT_F_ROWS : FT_AS_ROWS;
#pragma unroll
    for (int s = 0; s < FT_SLOTS; s++) {
      bool mine = (s >= lo && s < hi) || (PHASE == 1 && s >= FT_AS_ROWS);
      if (!mine) continue;
      uint32_t t = warp_sum(vals[s]);
      if (lane == 0 && warp < nwarps) p.warp_tot[(size_t)s * nwarps + warp] = t;
    }
  }
}

__global__ void k_l7_ft_clear_pos(FtEntry* t, uint32_t mask) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= mask; i += gridDim.x * blockDim.x) t[i].first_pos = 0;
}

// placement of the three flow_tag tables inside the ft arena
__global__ void k_l7_ft_layout(const BatchParams p) {
  DynState* dyn = p.dyn;
  const L7Dev* d = reinterpret_cast<const L7Dev*>(p.ext);
  if ((dyn->stats.need_repair && !p.repair) || dyn->overflow) return;
  uint64_t* tot = dyn->stats.ft_tot;
  uint64_t* fo = dyn->stats.ft_off;
  uint64_t cur = 0;
  auto take = [&](int id, uint64_t bytes) {
    fo[id] = cur;
    cur += (bytes + 255) & ~255ull;
  };
  uint64_t nfv = tot[FT_FV_ROWS], nf = tot[FT_F_ROWS], nas = tot[FT_AS_ROWS];
  take(FO_FV_TIME, nfv * 4); take(FO_FV_VPC, nfv * 4); take(FO_FV_PODNS, nfv * 2); take(FO_FV_FTYPE, nfv);
  take(FO_FV_FVTYPE, nfv); take(FO_FV_TEAM, nfv * 2); take(FO_FV_TABLE, nfv); take(FO_FV_NAME_POS, nfv * 16);
  take(FO_FV_VAL_POS, nfv * 16); take(FO_FV_COUNT, nfv * 8); take(FO_FV_NAME_BUF, tot[FT_FV_NAME_B]);
  take(FO_FV_VAL_BUF, tot[FT_FV_VAL_B]);
  take(FO_F_TIME, nf * 4); take(FO_F_VPC, nf * 4); take(FO_F_PODNS, nf * 2); take(FO_F_FTYPE, nf);
  take(FO_F_FVTYPE, nf); take(FO_F_TEAM, nf * 2); take(FO_F_TABLE, nf); take(FO_F_NAME_POS, nf * 16);
  take(FO_F_NAME_BUF, tot[FT_F_NAME_B]);
  take(FO_AS_TIME, nas * 4); take(FO_AS_TABLE, nas); take(FO_AS_SVC_POS, nas * 16); take(FO_AS_INST_POS, nas * 16);
  take(FO_AS_TEAM, nas * 2); take(FO_AS_BUF, tot[FT_AS_B]);
  dyn->stats.ft_end = cur;
  dyn->stats.ft_overflow = cur > d->ft_arena_bytes ? 1 : 0;
}

}  // namespace

static int l7_ensure_state(dfgpu_ctx* ctx) {
  if (ctx->ft) return 0;
  L7Host* h = new L7Host();
  // ---- constant strings ----
  std::vector<std::string> cs(CS_FIXED_COUNT);
  cs[CS_ATTR_UA] = "http_user_agent";
  cs[CS_ATTR_REFERER] = "http_referer";
  cs[CS_ATTR_RPC] = "rpc_service";
  cs[CS_FT_APP_SERVICE] = "app_service";
  cs[CS_FT_ENDPOINT] = "endpoint";
  cs[CS_FT_APP_INSTANCE] = "app_instance";
  cs[CS_PROTO_NA] = "N/A";
  cs[CS_PROTO_TLS] = "TLS";
  cs[CS_DNS_AVAILABLE] = "available for assignment";
  cs[CS_DNS_PRIVATE] = "Private Use";
  cs[CS_KAFKA_UNKNOWN] = DFGPU_EXC_KAFKA_UNKNOWN_SERVER_ERROR;
  L7Dev& d = h->dev;
  memset(&d, 0, sizeof d);
  for (int i = 0; i < 256; i++) d.proto_cs[i][0] = d.proto_cs[i][1] = CS_PROTO_NA;
  struct PN { int id; const char* name; };
  static const PN protos[] = {  // L7Protocol.String, datatype/flow.go:623-803
      {20, "HTTP"}, {21, "HTTP2"}, {40, "Dubbo"}, {41, "gRPC"}, {43, "SofaRPC"}, {44, "FastCGI"}, {45, "bRPC"},
      {46, "Tars"}, {47, "SOME/IP"}, {48, "ISO-8583"}, {49, "TRIPLE"}, {50, "NetSign"}, {60, "MySQL"},
      {61, "PostgreSQL"}, {62, "Oracle"}, {63, "Dameng"}, {80, "Redis"}, {81, "MongoDB"}, {82, "Memcached"},
      {100, "Kafka"}, {101, "MQTT"}, {102, "AMQP"}, {103, "OpenWire"}, {104, "NATS"}, {105, "Pulsar"},
      {106, "ZMTP"}, {107, "RocketMQ"}, {108, "WebSphereMQ"}, {120, "DNS"}, {127, "Custom"}};
  for (auto& pn : protos) {
    d.proto_cs[pn.id][0] = (uint16_t)cs.size();
    cs.push_back(pn.name);
    d.proto_cs[pn.id][1] = (uint16_t)cs.size();
    cs.push_back(std::string(pn.name) + "_TLS");
  }
  d.proto_cs[121][0] = d.proto_cs[121][1] = CS_PROTO_TLS;
  auto add_table = [&](const char* const* t, int n) {
    uint32_t base = (uint32_t)cs.size();
    for (int i = 0; i < n; i++) cs.push_back(t[i]);
    return base;
  };
  d.cs_http = add_table(dfgpu_exc_http, DFGPU_EXC_HTTP_LEN);
  d.cs_dns = add_table(dfgpu_exc_dns, DFGPU_EXC_DNS_LEN);
  d.cs_dubbo = add_table(dfgpu_exc_dubbo, DFGPU_EXC_DUBBO_LEN);
  d.cs_kafka = add_table(dfgpu_exc_kafka, DFGPU_EXC_KAFKA_LEN);
  d.cs_mqtt3 = add_table(dfgpu_exc_mqtt_v3, DFGPU_EXC_MQTT_V3_LEN);
  d.cs_mqtt5 = add_table(dfgpu_exc_mqtt_v5, DFGPU_EXC_MQTT_V5_LEN);
  std::vector<uint8_t> blob;
  std::vector<uint32_t> off;
  for (auto& s : cs) {
    off.push_back((uint32_t)blob.size());
    blob.insert(blob.end(), s.begin(), s.end());
  }
  off.push_back((uint32_t)blob.size());
  CU(cudaMalloc(&h->d_blob, blob.size() + 16));
  CU(cudaMalloc(&h->d_off, off.size() * 4));
  CU(cudaMemcpy(h->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
  d.cs_blob = h->d_blob;
  d.cs_off = h->d_off;
  // ---- device tag set: never evicts, sized >= 8x the reference's three LRU capacities ----
  uint64_t want = (uint64_t)ctx->cfg.flow_tag_cache_max_size * 3 * 8;
  uint32_t n = 1u << 16;
  while (n < want && n < (1u << 28)) n <<= 1;
  CU(cudaMalloc(&h->d_ft, (size_t)n * sizeof(FtEntry)));
  CU(cudaMemset(h->d_ft, 0, (size_t)n * sizeof(FtEntry)));
  d.ft = h->d_ft;
  d.ft_mask = n - 1;
  d.ft_timeout = ctx->cfg.flow_tag_cache_flush_timeout;
  h->ft_arena_cap = 16u << 20;
  CU(cudaMalloc(&h->d_ft_arena, h->ft_arena_cap));
  d.ft_arena = h->d_ft_arena;
  d.ft_arena_bytes = h->ft_arena_cap;
  CU(cudaMalloc(&h->d_dev, sizeof(L7Dev)));
  ctx->ft = h;
  return 0;
}

// called by dfgpu_run before the passes: (re)publish the device-side L7 state
int l7_prepare(dfgpu_ctx* ctx, cudaStream_t s) {
  int rc = l7_ensure_state(ctx);
  if (rc) return rc;
  L7Host* h = reinterpret_cast<L7Host*>(ctx->ft);
  h->epoch++;
  if ((h->epoch & 0xffff) == 0) {
    // the 16-bit epoch inside first_pos wrapped: forget every stale candidate
    k_l7_ft_clear_pos<<<1024, 256, 0, s>>>(h->d_ft, h->dev.ft_mask);
    h->epoch = 1;
  }
  h->dev.epoch = h->epoch;
  h->dev.org_id = ctx->cur_org;
  h->dev.ft_arena = h->d_ft_arena;
  h->dev.ft_arena_bytes = h->ft_arena_cap;
  CU(cudaMemcpyAsync(h->d_dev, &h->dev, sizeof(L7Dev), cudaMemcpyHostToDevice, s));
  ctx->bp.ext = h->d_dev;
  return 0;
}

static void launch_ft_writer(dfgpu_ctx* ctx, cudaStream_t s) {
  BatchParams& bp = ctx->bp;
  if (!bp.nrec) return;
  k_l7_flow_tag<3><<<(bp.nrec + 255) / 256, 256, 0, s>>>(bp);
  ctx->launches++;
}

int l7_flow_tags(dfgpu_ctx* ctx, cudaStream_t s) {
  BatchParams& bp = ctx->bp;
  if (bp.nrec) {
    dim3 grid((bp.nrec + 255) / 256), block(256);
    k_l7_flow_tag<0><<<grid, block, 0, s>>>(bp);
    k_l7_flow_tag<1><<<grid, block, 0, s>>>(bp);
    k_l7_flow_tag<2><<<grid, block, 0, s>>>(bp);
    launch_slot_scan_to(bp, FT_SLOTS, ctx->d_dyn->stats.ft_tot, s);
    ctx->launches += 4;
  }
  k_l7_ft_layout<<<1, 1, 0, s>>>(bp);
  ctx->launches++;
  launch_ft_writer(ctx, s);
  launch_commit(bp, 1, s);
  ctx->launches++;
  CU(cudaGetLastError());
  return 0;
}

static void ft_lc_const(dfgpu_batch* b, dfgpu_col& col, const std::vector<std::string>& dict) {
  b->owned.emplace_back();
  b->owned_pos.emplace_back();
  auto& buf = b->owned.back();
  auto& pos = b->owned_pos.back();
  for (auto& s : dict) {
    pos.push_back(dfgpu_pos{(int64_t)buf.size(), (int64_t)(buf.size() + s.size())});
    buf.insert(buf.end(), s.begin(), s.end());
  }
  col.dict_data = buf.data();
  col.dict_pos = pos.data();
  col.dict_n = pos.size();
}

// LowCardinality dictionaries of the flow_tag tables hold at most two entries per
// column; keys are written as raw codes and remapped here (rows are few).
static void ft_lc_coded(dfgpu_batch* b, dfgpu_col& col, uint8_t* keys, uint64_t n, const char* name0,
                        const char* name1) {
  int first[2] = {-1, -1};
  for (uint64_t i = 0; i < n && (first[0] < 0 || first[1] < 0); i++)
    if (first[keys[i]] < 0) first[keys[i]] = (int)i;
  std::vector<std::string> dict;
  uint8_t remap[2] = {0, 0};
  if (first[0] >= 0 && (first[1] < 0 || first[0] < first[1])) {
    dict.push_back(name0);
    if (first[1] >= 0) { dict.push_back(name1); remap[1] = 1; }
  } else if (first[1] >= 0) {
    dict.push_back(name1);
    remap[1] = 0;
    if (first[0] >= 0) { dict.push_back(name0); remap[0] = 1; }
  }
  if (remap[0] != 0 || remap[1] != 1)
    for (uint64_t i = 0; i < n; i++) keys[i] = remap[keys[i]];
  ft_lc_const(b, col, dict);
}

int l7_fetch_tables(dfgpu_ctx* ctx, dfgpu_batch* b) {
  L7Host* h = reinterpret_cast<L7Host*>(ctx->ft);
  cudaStream_t s = ctx->stream;
  DynState* d = ctx->h_dyn;
  if (d->stats.ft_full)
    return fail(ctx, DFGPU_ENOMEM, "flow_tag device set is full (raise flow_tag_cache_max_size)");
  if (d->stats.ft_overflow) {
    // grow the ft arena and re-run only the writer: the emit decisions are already
    // recorded in the tag set and do not change
    size_t want = (size_t)d->stats.ft_end + (1u << 20);
    CU(cudaFree(h->d_ft_arena));
    h->d_ft_arena = nullptr;
    CU(cudaMalloc(&h->d_ft_arena, want));
    h->ft_arena_cap = want;
    h->dev.ft_arena = h->d_ft_arena;
    h->dev.ft_arena_bytes = want;
    CU(cudaMemcpyAsync(h->d_dev, &h->dev, sizeof(L7Dev), cudaMemcpyHostToDevice, s));
    uint32_t zero = 0;
    CU(cudaMemcpyAsync(&ctx->d_dyn->stats.ft_overflow, &zero, 4, cudaMemcpyHostToDevice, s));
    launch_ft_writer(ctx, s);
    CU(cudaStreamSynchronize(s));
  }
  uint64_t bytes = d->stats.ft_end;
  b->owned.emplace_back(bytes + 16);
  uint8_t* hb = b->owned.back().data();
  if (bytes) {
    CU(cudaMemcpyAsync(hb, h->d_ft_arena, bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  const uint64_t* fo = d->stats.ft_off;
  uint64_t nfv = d->stats.ft_tot[FT_FV_ROWS], nf = d->stats.ft_tot[FT_F_ROWS], nas = d->stats.ft_tot[FT_AS_ROWS];
  auto fixed = [&](dfgpu_batch::Table& t, const char* name, int ch, int es, uint64_t off, uint64_t n) {
    dfgpu_col c;
    memset(&c, 0, sizeof c);
    c.name = name;
    c.ch_type = ch;
    c.elem_size = es;
    c.nrows = n;
    c.data = hb + off;
    c.data_bytes = n * es;
    t.cols.push_back(c);
    return &t.cols.back();
  };
  auto strcol = [&](dfgpu_batch::Table& t, const char* name, int ch, uint64_t pos_off, uint64_t buf_off,
                    uint64_t buf_bytes, uint64_t n) {
    dfgpu_col c;
    memset(&c, 0, sizeof c);
    c.name = name;
    c.ch_type = ch;
    c.nrows = n;
    c.pos = (const dfgpu_pos*)(hb + pos_off);
    c.data = hb + buf_off;
    c.data_bytes = buf_bytes;
    t.cols.push_back(c);
  };
  for (int which = 0; which < 2; which++) {  // 0: custom_field, 1: custom_field_value
    b->tables.emplace_back();
    dfgpu_batch::Table& t = b->tables.back();
    t.cols.reserve(12);
    bool val = which == 1;
    uint64_t n = val ? nfv : nf;
    t.id = val ? DFGPU_TABLE_L7_CUSTOM_FIELD_VALUE : DFGPU_TABLE_L7_CUSTOM_FIELD;
    t.nrows = n;
    fixed(t, "time", DFGPU_CH_DATETIME, 4, fo[val ? FO_FV_TIME : FO_F_TIME], n);
    ft_lc_const(b, *fixed(t, "table", DFGPU_CH_LC_STR, 1, fo[val ? FO_FV_TABLE : FO_F_TABLE], n),
                n ? std::vector<std::string>{"l7_flow_log"} : std::vector<std::string>{});
    fixed(t, "vpc_id", DFGPU_CH_INT32, 4, fo[val ? FO_FV_VPC : FO_F_VPC], n);
    fixed(t, "pod_ns_id", DFGPU_CH_UINT16, 2, fo[val ? FO_FV_PODNS : FO_F_PODNS], n);
    ft_lc_coded(b, *fixed(t, "field_type", DFGPU_CH_LC_STR, 1, fo[val ? FO_FV_FTYPE : FO_F_FTYPE], n),
                hb + fo[val ? FO_FV_FTYPE : FO_F_FTYPE], n, "tag", "metrics");
    strcol(t, "field_name", DFGPU_CH_LC_STR_ROWS, fo[val ? FO_FV_NAME_POS : FO_F_NAME_POS],
           fo[val ? FO_FV_NAME_BUF : FO_F_NAME_BUF], d->stats.ft_tot[val ? FT_FV_NAME_B : FT_F_NAME_B], n);
    ft_lc_coded(b, *fixed(t, "field_value_type", DFGPU_CH_LC_STR, 1, fo[val ? FO_FV_FVTYPE : FO_F_FVTYPE], n),
                hb + fo[val ? FO_FV_FVTYPE : FO_F_FVTYPE], n, "string", "float");
    fixed(t, "team_id", DFGPU_CH_UINT16, 2, fo[val ? FO_FV_TEAM : FO_F_TEAM], n);
    if (val) {
      strcol(t, "field_value", DFGPU_CH_STR, fo[FO_FV_VAL_POS], fo[FO_FV_VAL_BUF], d->stats.ft_tot[FT_FV_VAL_B], n);
      fixed(t, "count", DFGPU_CH_UINT64, 8, fo[FO_FV_COUNT], n);
    }
  }
  {
    b->tables.emplace_back();
    dfgpu_batch::Table& t = b->tables.back();
    t.cols.reserve(8);
    t.id = DFGPU_TABLE_L7_APP_SERVICE;
    t.nrows = nas;
    fixed(t, "time", DFGPU_CH_DATETIME, 4, fo[FO_AS_TIME], nas);
    ft_lc_const(b, *fixed(t, "table", DFGPU_CH_LC_STR, 1, fo[FO_AS_TABLE], nas),
                nas ? std::vector<std::string>{"l7_flow_log"} : std::vector<std::string>{});
    strcol(t, "app_service", DFGPU_CH_LC_STR_ROWS, fo[FO_AS_SVC_POS], fo[FO_AS_BUF], d->stats.ft_tot[FT_AS_B], nas);
    strcol(t, "app_instance", DFGPU_CH_LC_STR_ROWS, fo[FO_AS_INST_POS], fo[FO_AS_BUF], 0, nas);
    fixed(t, "team_id", DFGPU_CH_UINT16, 2, fo[FO_AS_TEAM], nas);
  }
  b->counters.new_field_count = (int64_t)nf;
  b->counters.new_field_value_count = (int64_t)nfv;
  b->counters.rows[DFGPU_TABLE_L7_CUSTOM_FIELD] = (int64_t)nf;
  b->counters.rows[DFGPU_TABLE_L7_CUSTOM_FIELD_VALUE] = (int64_t)nfv;
  b->counters.rows[DFGPU_TABLE_L7_APP_SERVICE] = (int64_t)nas;
  return 0;
}

}  // namespace dfgpu

extern "C" void dfgpu_l7_free_state(dfgpu_ctx* ctx) {
  using namespace dfgpu;
  if (!ctx || !ctx->ft) return;
  L7Host* h = reinterpret_cast<L7Host*>(ctx->ft);
  cudaFree(h->d_blob);
  cudaFree(h->d_off);
  cudaFree(h->d_ft);
  cudaFree(h->d_ft_arena);
  cudaFree(h->d_dev);
  delete h;
  ctx->ft = nullptr;
}
