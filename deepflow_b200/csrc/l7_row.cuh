/*
 * l7_row.cuh — one decoded AppProtoLogsData record -> the values of its l7_flow_log row.
 * Pure per-record logic (no warp collectives, no CUDA types): the decode kernel (l7_pass.cu) runs it on
 * the device with an emitter that stores coalesced columns; tests/harness runs the same code on the host
 * against the oracle.  Restates
 *   ProtoLogToL7FlowLog  flow_log/log_data/l7_flow_log.go:638-644, L7FlowLog.Fill :312-342,
 *   L7Base.Fill :514-597, fillL7FlowLog :363-452, fillTraceIds :344-360, fillExceptionDesc :463-488
 *   (tables: exception_desc.go), L7Protocol.String datatype/flow.go:623-803,
 *   ParseTraceIdIndex :296-310 -> DJBHash libs/utils/hash.go:37-43,
 *   L7FlowLog.AppendToColumnBlock l7_column_block.go:167-208,376-422 (column order: schema_l7.h)
 */
#pragma once
#include "l7_tok.h"
#include "schema_l7.h"

#include "../../include/dfgpu.h"
#include "../../include/dfgpu_exception_tables.h"

namespace dfgpu {

// ---- constant strings the path can emit (protocol names, pseudo attribute names,
// exception descriptions): one device blob, addressed by id -------------------------
enum : uint32_t {
  CS_EMPTY = 0, CS_ATTR_UA, CS_ATTR_REFERER, CS_ATTR_RPC, CS_FT_APP_SERVICE, CS_FT_ENDPOINT,
  CS_FT_APP_INSTANCE, CS_PROTO_NA, CS_PROTO_TLS, CS_DNS_AVAILABLE, CS_DNS_PRIVATE, CS_KAFKA_UNKNOWN,
  CS_FIXED_COUNT
};

struct L7Tables {
  const uint8_t* cs_blob;
  const uint32_t* cs_off;  // [n+1]
  uint32_t cs_http, cs_dns, cs_dubbo, cs_kafka, cs_mqtt3, cs_mqtt5;  // base ids of the tables
  uint16_t proto_cs[256][2];  // L7Protocol.String(isTLS) -> cs id
};

// A string of the row: bytes [off, off + len) of the input buffer, or (REF_CS set) of the constant blob.
struct SRef32 {
  uint32_t off, len;
};
constexpr uint32_t REF_CS = 0x80000000u;
DF_HD const uint8_t* ref_ptr(const L7Tables& tb, const uint8_t* in, SRef32 r) {
  return (r.off & REF_CS) ? tb.cs_blob + (r.off & ~REF_CS) : in + r.off;
}
DF_HD SRef32 cs_ref32(const L7Tables& tb, uint32_t id) {
  const uint32_t a = tb.cs_off[id], b = tb.cs_off[id + 1];
  return SRef32{a | REF_CS, b - a};
}
DF_HD SRef32 cs_tbl32(const L7Tables& tb, uint32_t base, int32_t idx, int32_t n) {
  if (idx >= 0 && idx < n) return cs_ref32(tb, base + (uint32_t)idx);
  return SRef32{0, 0};
}

// canonical string ids of flow_metrics.TAPSideEnum.String() (tag.go:211-233); 0 = "" (unnamed)
DF_HD uint32_t tapside_code_hd(uint32_t v) {
  switch (v & 0xff) {
    case 0: return 1;  case 1: return 2;  case 2: return 3;  case 4: return 4;
    case 9: return 5;  case 10: return 6; case 17: return 7; case 18: return 8;
    case 25: return 9; case 26: return 10; case 33: return 11; case 34: return 12;
    case 41: return 13; case 42: return 14; case 49: return 15; case 50: return 16;
    case 48: return 17;
  }
  return 0;
}

// DJBHash (libs/utils/hash.go:37-43) over n input bytes
DF_HD uint64_t djb_hash32(uint64_t h, const uint8_t* p, uint32_t n) {
  if (!n) return h;
  const unsigned long long* sb = reinterpret_cast<const unsigned long long*>((uintptr_t)p & ~(uintptr_t)7);
  uint32_t o = (uint32_t)(uintptr_t)p & 7u;
  unsigned long long cur = DF_LDG(sb) >> (o * 8);
  for (uint32_t i = 0; i < n; i++) {
    h = ((h << 5) + h) + (uint64_t)(cur & 0xff);
    cur >>= 8;
    if (++o == 8 && i + 1 < n) {
      o = 0;
      cur = DF_LDG(++sb);
    }
  }
  return h;
}
DF_HD bool str_eq32(const uint8_t* p, uint32_t n, const char* lit) {
  uint32_t i = 0;
  for (; lit[i]; i++)
    if (i >= n || DF_LDG(p + i) != (uint8_t)lit[i]) return false;
  return i == n;
}

// What the kernel needs back for its warp-level work (everything else went through the emitter).
struct L7RowInfo {
  uint32_t tside;        // canonical tap-side code (LowCardinality observation_point)
  uint32_t l7_protocol;  // updateCounter class
  uint32_t trace_len;    // length of the row's trace id (0: none)
  uint64_t trace_hash;   // DJBHash(17, trace id) when trace_len != 0
  uint32_t z_an_e, z_an_b, z_av_e, z_av_b, z_mn_e, z_mn_b, z_mv_e;  // array element / byte counts
};

// S: decoded record.   uint32_t w(int word); uint64_t w64(int word); SRef32 ref(int raw)   (absent: 0)
// E: row sink.
//   template <class T> void fix(int col, T v);                 fixed-width column value
//   void ip6(int col, SRef32 r);                               16 bytes, zero padded (nil -> zeros)
//   template <class T> void nullable(int col, bool has, T v);
//   void str(int stream, SRef32 r);                            per-row string column (stream order L7V_*)
//   void aux(int which, SRef32 r);                             0 ext_info range, 1..3 user agent / referer / rpc
//   void kg(...)                                               SmartEncoding inputs (see l7_pass.cu)
// `nat`: ClickHouse native String encoding sizes (dfgpu_cfg.native_strings).
template <class S, class E>
DF_HD L7RowInfo l7_build_row(const L7Tables& tb, const uint8_t* in, const S& sl, E& e, const uint32_t nat,
                             const uint32_t analyzer_id, const uint32_t id_count /* id_counter + row + 1 */) {
  const SRef32 none{0, 0};
  L7RowInfo o;
  // every wire reference of the record, requested together (on the device they come from the warp's
  // scratch in L2: one latency for all of them instead of one per use); constant indices -> registers
  SRef32 rr[R7_COUNT];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < R7_COUNT; i++) rr[i] = sl.ref(i);
  const bool has_req = sl.w(W7_REQ_CNT) != 0, has_resp = sl.w(W7_RESP_CNT) != 0, has_ext = sl.w(W7_EXT_CNT) != 0,
             has_trace = sl.w(W7_TRACE_CNT) != 0;
  const uint32_t flags = sl.w(W7_FLAGS);
  const uint8_t type = (uint8_t)sl.w(W7_HEAD_MSG_TYPE);
  const uint8_t is_tls = (flags & 1) ? 1 : 0, is_async = (flags & 2) ? 1 : 0, is_reversed = (flags & 4) ? 1 : 0;
  const uint8_t l7_protocol = (uint8_t)sl.w(W7_HEAD_PROTO);
  uint8_t response_status = 5;  // STATUS_UNKNOWN
  const int32_t request_length = (int32_t)sl.w(W7_REQ_LEN), response_length = (int32_t)sl.w(W7_RESP_LEN);
  const uint32_t sql_rows = sl.w(W7_ROW_EFFECT);
  bool has_request_length = false, has_response_length = false, has_response_code = false, has_request_id = false;
  int32_t response_code = 0;
  uint32_t request_id = 0;
  SRef32 exc = none;
  if (has_req && request_length != -1 && type != 1) has_request_length = true;  // :379-387
  if (has_resp) {                                                                // :389-409
    response_status = (uint8_t)sl.w(W7_RESP_STATUS);
    if (type != 0) {
      const int32_t code = (int32_t)sl.w(W7_RESP_CODE);
      response_code = code;
      exc = rr[R7_EXCEPTION];
      if (exc.len == 0 && (response_status == 3 || response_status == 4)) {  // fillExceptionDesc :463-488
        switch (l7_protocol) {
          case 20:
          case 21: exc = cs_tbl32(tb, tb.cs_http, (int32_t)(uint16_t)code, DFGPU_EXC_HTTP_LEN); break;
          case 120: {
            const uint16_t c = (uint16_t)code;
            if (c < DFGPU_EXC_DNS_LEN) exc = cs_ref32(tb, tb.cs_dns + c);
            else if (c >= 3841 && c <= 4095) exc = cs_ref32(tb, CS_DNS_PRIVATE);
            else exc = cs_ref32(tb, CS_DNS_AVAILABLE);
            break;
          }
          case 40: exc = cs_tbl32(tb, tb.cs_dubbo, (int32_t)(uint16_t)code, DFGPU_EXC_DUBBO_LEN); break;
          case 100: {
            const int16_t c = (int16_t)code;
            if (c == -1) exc = cs_ref32(tb, CS_KAFKA_UNKNOWN);
            else if (c > 0 && c < DFGPU_EXC_KAFKA_LEN) exc = cs_ref32(tb, tb.cs_kafka + c);
            else exc = none;
            break;
          }
          case 101: {
            const uint16_t c = (uint16_t)code;
            const SRef32 ver = rr[R7_VERSION];
            const bool v5 = str_eq32(in + ver.off, ver.len, "5");
            exc = none;
            if (!v5) { if (c > 0 && c < DFGPU_EXC_MQTT_V3_LEN) exc = cs_ref32(tb, tb.cs_mqtt3 + c); }
            else { if (c > 0 && c < DFGPU_EXC_MQTT_V5_LEN) exc = cs_ref32(tb, tb.cs_mqtt5 + c); }
            break;
          }
          default: break;  // h.ResponseException = l.Resp.Exception (empty)
        }
      }
      if (response_code != -32768) has_response_code = true;
      if (response_length != -1) has_response_length = true;
    } else {
      exc = none;
    }
  }
  uint32_t n_pseudo = 0, pseudo_name_bytes = 0, pseudo_value_bytes = 0;
  SRef32 ua = none, rf = none, rpc = none;
  if (has_ext) {  // :411-433
    request_id = sl.w(W7_REQUEST_ID);
    if (request_id != 0) has_request_id = true;
    ua = rr[R7_USER_AGENT];
    rf = rr[R7_REFERER];
    rpc = rr[R7_RPC_SERVICE];
    // nat: every element carries its uvarint length in the stream (names are < 128 bytes: one byte)
    const uint32_t pv = nat ? ~0u : 0u;
    if (ua.len) { n_pseudo++; pseudo_name_bytes += 15 + nat; pseudo_value_bytes += ua.len + (tk_vsize(ua.len) & pv); }
    if (rf.len) { n_pseudo++; pseudo_name_bytes += 12 + nat; pseudo_value_bytes += rf.len + (tk_vsize(rf.len) & pv); }
    if (rpc.len) { n_pseudo++; pseudo_name_bytes += 11 + nat; pseudo_value_bytes += rpc.len + (tk_vsize(rpc.len) & pv); }
  }
  SRef32 trace_id = none;
  bool two_ids = false;
  if (has_trace) {  // :434-438 + fillTraceIds :344-360
    const uint32_t n_ids = sl.w(W7_TRACE_IDS_N);
    if (n_ids > 0) trace_id = rr[R7_TRACE_IDS_0];
    two_ids = n_ids > 1;
    if (trace_id.len == 0) trace_id = rr[R7_TRACE_ID];
  }
  if (trace_id.len == 0) trace_id = none;
  switch (l7_protocol) {  // :438-451
    case 100:
      if (has_req) {
        const SRef32 rt = rr[R7_REQ_TYPE];
        const uint8_t* q = in + rt.off;
        const bool supported = str_eq32(q, rt.len, "Fetch") || str_eq32(q, rt.len, "Produce") ||
                               str_eq32(q, rt.len, "JoinGroup") || str_eq32(q, rt.len, "LeaveGroup") ||
                               str_eq32(q, rt.len, "SyncGroup");
        if (response_code == 0 && !supported) {
          response_status = 2;
          has_response_code = false;
        }
        has_request_id = true;
      }
      break;
    case 43: has_request_id = true; break;
  }

  // ---- L7Base.Fill :514-582 -----------------------------------------------------------------
  const bool v6 = sl.w(W7_IS_IPV6) == 1;
  const uint64_t end_time = sl.w64(W7_END_TIME);
  const uint32_t time_s = (uint32_t)(end_time / 1000000000ull);
  const uint32_t vtap_id = sl.w(W7_VTAP_ID);
  const uint64_t tap_port = sl.w64(W7_TAP_PORT);
  e.template fix<uint32_t>(L7C_TIME, time_s);
  e.template fix<uint32_t>(L7C_IP4_0, v6 ? 0u : sl.w(W7_IP_SRC));
  e.template fix<uint32_t>(L7C_IP4_1, v6 ? 0u : sl.w(W7_IP_DST));
  e.ip6(L7C_IP6_0, v6 ? rr[R7_IP6_SRC] : none);
  e.ip6(L7C_IP6_1, v6 ? rr[R7_IP6_DST] : none);
  e.template fix<uint8_t>(L7C_IS_IPV4, v6 ? 0 : 1);
  e.template fix<uint8_t>(L7C_PROTOCOL, (uint8_t)sl.w(W7_PROTOCOL));
  e.template fix<uint16_t>(L7C_CLIENT_PORT, (uint16_t)sl.w(W7_PORT_SRC));
  e.template fix<uint16_t>(L7C_SERVER_PORT, (uint16_t)sl.w(W7_PORT_DST));
  e.template fix<uint64_t>(L7C_FLOW_ID, sl.w64(W7_FLOW_ID));
  e.template fix<uint8_t>(L7C_CAPTURE_NETWORK_TYPE_ID, (uint8_t)sl.w(W7_TAP_TYPE));
  const uint8_t tap_port_type = (uint8_t)(tap_port >> 60);
  e.template fix<uint8_t>(L7C_NAT_SOURCE, (uint8_t)((tap_port >> 36) & 0xf));
  e.template fix<uint8_t>(L7C_CAPTURE_NIC_TYPE, tap_port_type);
  e.template fix<uint16_t>(L7C_SIGNAL_SOURCE, tap_port_type == 8 ? 4 : tap_port_type == 7 ? 3 : 0);
  e.template fix<uint8_t>(L7C_TUNNEL_TYPE, (uint8_t)((tap_port >> 32) & 0xf));
  e.template fix<uint32_t>(L7C_CAPTURE_NIC, (uint32_t)tap_port);
  e.template fix<uint16_t>(L7C_AGENT_ID, (uint16_t)vtap_id);
  e.template fix<uint32_t>(L7C_REQ_TCP_SEQ, sl.w(W7_REQ_TCP_SEQ));
  e.template fix<uint32_t>(L7C_RESP_TCP_SEQ, sl.w(W7_RESP_TCP_SEQ));
  e.template fix<int64_t>(L7C_START_TIME, (int64_t)sl.w64(W7_START_TIME) / 1000);
  e.template fix<int64_t>(L7C_END_TIME, (int64_t)end_time / 1000);
  e.template fix<uint8_t>(L7C_BIZ_TYPE, (uint8_t)sl.w(W7_BIZ_TYPE));
  e.template fix<int32_t>(L7C_PROCESS_ID_0, (int32_t)sl.w(W7_PID_0));
  e.template fix<int32_t>(L7C_PROCESS_ID_1, (int32_t)sl.w(W7_PID_1));
  e.template fix<uint64_t>(L7C_SYSCALL_TRACE_ID_REQUEST, sl.w64(W7_SYS_REQ));
  e.template fix<uint64_t>(L7C_SYSCALL_TRACE_ID_RESPONSE, sl.w64(W7_SYS_RESP));
  e.template fix<uint32_t>(L7C_SYSCALL_THREAD_0, sl.w(W7_THREAD_0));
  e.template fix<uint32_t>(L7C_SYSCALL_THREAD_1, sl.w(W7_THREAD_1));
  e.template fix<uint64_t>(L7C_SYSCALL_COROUTINE_0, sl.w64(W7_CORO_0));
  e.template fix<uint64_t>(L7C_SYSCALL_COROUTINE_1, sl.w64(W7_CORO_1));
  e.template fix<uint32_t>(L7C_SYSCALL_CAP_SEQ_0, sl.w(W7_CAP_SEQ_0));
  e.template fix<uint32_t>(L7C_SYSCALL_CAP_SEQ_1, sl.w(W7_CAP_SEQ_1));
  // ---- L7FlowLog block ------------------------------------------------------------------------
  // genID l4_flow_log.go:949-953
  e.template fix<uint64_t>(L7C__ID, (uint64_t)time_s << 32 | (uint64_t)(analyzer_id & 0x3ff) << 22 | ((uint64_t)id_count & 0x3fffff));
  e.template fix<uint8_t>(L7C_L7_PROTOCOL, l7_protocol);
  e.template fix<uint8_t>(L7C_TYPE, type);
  e.template fix<uint8_t>(L7C_IS_TLS, is_tls);
  e.template fix<uint8_t>(L7C_IS_ASYNC, is_async);
  e.template fix<uint8_t>(L7C_IS_REVERSED, is_reversed);
  e.template nullable<uint64_t>(L7C_REQUEST_ID, has_request_id, (uint64_t)request_id);
  e.template fix<uint8_t>(L7C_RESPONSE_STATUS, response_status);
  e.template nullable<int32_t>(L7C_RESPONSE_CODE, has_response_code, response_code);
  e.template nullable<uint8_t>(L7C_SPAN_KIND, false, 0);
  e.template fix<uint64_t>(L7C_RESPONSE_DURATION, sl.w64(W7_HEAD_RRT) / 1000);
  e.template nullable<int64_t>(L7C_REQUEST_LENGTH, has_request_length, (int64_t)request_length);
  e.template nullable<int64_t>(L7C_RESPONSE_LENGTH, has_response_length, (int64_t)response_length);
  e.template nullable<uint64_t>(L7C_SQL_AFFECTED_ROWS, sql_rows != 0, (uint64_t)sql_rows);
  e.template fix<uint8_t>(L7C_DIRECTION_SCORE, (uint8_t)sl.w(W7_DIRECTION_SCORE));
  e.template fix<uint32_t>(L7C_CAPTURED_REQUEST_BYTE, sl.w(W7_CAP_REQ_BYTE));
  e.template fix<uint32_t>(L7C_CAPTURED_RESPONSE_BYTE, sl.w(W7_CAP_RESP_BYTE));

  // ---- the 22 per-row strings, stream order L7V_* ----------------------------------------------
  e.str(L7V_BIZ_CODE, rr[R7_BIZ_CODE]);
  e.str(L7V_BIZ_SCENARIO, rr[R7_BIZ_SCENARIO]);
  e.str(L7V_BIZ_RESPONSE_CODE, rr[R7_BIZ_RESPONSE_CODE]);
  e.str(L7V_KNAME_0, rr[R7_KNAME_0]);
  e.str(L7V_KNAME_1, rr[R7_KNAME_1]);
  {
    const SRef32 ps = has_ext ? rr[R7_PROTOCOL_STR] : none;
    e.str(L7V_BIZ_PROTOCOL, ps.len ? ps : cs_ref32(tb, tb.proto_cs[l7_protocol][is_tls]));
  }
  e.str(L7V_VERSION, rr[R7_VERSION]);
  e.str(L7V_REQUEST_TYPE, has_req ? rr[R7_REQ_TYPE] : none);
  e.str(L7V_REQUEST_DOMAIN, has_req ? rr[R7_DOMAIN] : none);
  e.str(L7V_REQUEST_RESOURCE, has_req ? rr[R7_RESOURCE] : none);
  e.str(L7V_ENDPOINT, has_req ? rr[R7_ENDPOINT] : none);
  e.str(L7V_RESPONSE_EXCEPTION, exc);
  e.str(L7V_RESPONSE_RESULT, (has_resp && type != 0) ? rr[R7_RESULT] : none);
  e.str(L7V_HTTP_PROXY_CLIENT, has_ext ? rr[R7_CLIENT_IP] : none);
  e.str(L7V_X_REQUEST_ID_0, has_ext ? rr[R7_X_REQUEST_ID_0] : none);
  e.str(L7V_X_REQUEST_ID_1, has_ext ? rr[R7_X_REQUEST_ID_1] : none);
  e.str(L7V_TRACE_ID, trace_id);
  e.str(L7V_TRACE_ID_2, (has_trace && two_ids) ? rr[R7_TRACE_IDS_1] : none);
  e.str(L7V_SPAN_ID, has_trace ? rr[R7_SPAN_ID] : none);
  e.str(L7V_PARENT_SPAN_ID, has_trace ? rr[R7_PARENT_SPAN_ID] : none);
  e.str(L7V_APP_SERVICE, has_ext ? rr[R7_SERVICE_NAME] : none);
  e.str(L7V_APP_INSTANCE, none);  // only the OTel import fills it

  // ---- arrays: what the strings kernel needs to stream them out -----------------------------------
  // aux 0: the ext_info range when ONE occurrence carries all repeated fields (the common case), len 0
  // with the slow flag otherwise (merge semantics: full re-walk)
  const uint32_t n_an = sl.w(W7_ATTR_NAMES), n_av = sl.w(W7_ATTR_VALUES), n_mn = sl.w(W7_METRICS_NAMES),
                 n_mv = sl.w(W7_METRICS_VALUES);
  const bool any_rep = (n_an | n_av | n_mn | n_mv) != 0;
  // fast path of the arrays kernel: the element references the tokenizer left are complete — one ext_info
  // occurrence, no more elements than a block holds, metrics_values as one packed run
  const bool single = sl.w(W7_EXT_CNT) == 1 && sl.w(W7_BASE_CNT) <= 1 && n_an <= L7E_CAP_ATTR && n_av <= L7E_CAP_ATTR &&
                      n_mn <= L7E_CAP_METRICS && sl.w(W7_METRICS_VALUES_OCC) == (n_mv ? 1u : 0u);
  e.aux(0, (any_rep && single) ? rr[R7_EXT_RANGE] : none);
  e.aux(1, ua);
  e.aux(2, rf);
  e.aux(3, rpc);
  e.flags(has_ext, any_rep && !single);
  o.z_an_e = n_pseudo + n_an;
  o.z_an_b = pseudo_name_bytes + sl.w(W7_ATTR_NAMES_B);
  o.z_av_e = n_pseudo + n_av;
  o.z_av_b = pseudo_value_bytes + sl.w(W7_ATTR_VALUES_B);
  o.z_mn_e = n_mn;
  o.z_mn_b = sl.w(W7_METRICS_NAMES_B);
  o.z_mv_e = n_mv;

  // ---- SmartEncoding inputs (KnowledgeGraph.FillL7 :599-613 runs in l7_kg.cu) ----------------------
  e.kg(vtap_id, sl.w(W7_TAP_SIDE), sl.w(W7_PROTOCOL), sl.w(W7_PORT_DST), v6, sl.w(W7_VIP_SRC) == 1,
       sl.w(W7_VIP_DST) == 1, sl.w(W7_IP_SRC), sl.w(W7_IP_DST), sl.w(W7_EPC_SRC), sl.w(W7_EPC_DST),
       sl.w64(W7_MAC_SRC), sl.w64(W7_MAC_DST), sl.w(W7_GPID_0), sl.w(W7_GPID_1), sl.w(W7_POD_0), sl.w(W7_POD_1),
       sl.w(W7_PID_0), sl.w(W7_PID_1), rr[R7_IP6_SRC], rr[R7_IP6_DST]);

  o.tside = tapside_code_hd(sl.w(W7_TAP_SIDE));
  o.l7_protocol = l7_protocol;
  o.trace_len = trace_id.len;
  // ParseTraceIdIndex :296-310: the hash of the row's trace id; rows without one take the previous
  // row's index + 1 (resolved by the scan over the batch)
  o.trace_hash = trace_id.len ? djb_hash32(17, ref_ptr(tb, in, trace_id), trace_id.len) : 0;
  return o;
}

}  // namespace dfgpu
