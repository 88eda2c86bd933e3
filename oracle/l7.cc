/*
 * ORACLE — test infrastructure only.
 * l7.cc: CPU restatement of the L7 flow-log row path
 *   sendProto / updateCounter     server/ingester/flow_log/decoder/decoder.go:474-525
 *   ProtoLogToL7FlowLog           server/ingester/flow_log/log_data/l7_flow_log.go:638-644
 *   L7FlowLog.Fill                :312-342 ; L7Base.Fill :514-597 ; FillL7 :599-613
 *   fillL7FlowLog                 :363-452 ; fillTraceIds :344-360 ; fillExceptionDesc :463-488
 *   ParseTraceIdIndex             :296-310 (hash mode: utils.GetTraceIdIndex utils.go:266-269)
 *   GenerateNewFlowTags           :648-745 ; lru.Cache lru.go:44-76
 *   AppServiceTagWriter.Write     flow_tag/app_service_tag_writer.go:92-113
 *   column blocks                 l7_column_block.go:114-158,167-208,376-422;
 *                                 flow_tag_column_block.go:50-101;
 *                                 app_service_tag_column_block.go:39-80
 * Pinned: column lists (machine-extracted from ToInput), L7Protocol.String (machine-extracted switch), exception
 * tables (generated from exception_desc.go), DJBHash, lru order.  Parity unpinned by the reference: the Fill functions,
 * GenerateNewFlowTags, AppendToColumnBlock — cross-checked by tests/test_l7_row_crosscheck_cpu.py and
 * tests/test_flow_tag_crosscheck_cpu.py.
 * Presence of req/resp/trace_info/ext_info follows the wire (a fresh
 * pb.AppProtoLogsData); the reference's pooled objects keep reset-but-non-nil
 * sub-messages after first use (datatype/pb/flow_log.go:37-77), which only
 * differs for records that omit a sub-message entirely.
 */
#include "l7.h"

#include "../include/dfgpu_exception_tables.h"

namespace oracle {

uint64_t DJBHash(uint64_t base, std::string_view s) {
  uint64_t hash = base;
  for (unsigned char ch : s) hash = ((hash << 5) + hash) + (uint64_t)ch;
  return hash;
}

std::string_view L7ProtocolString(uint8_t p, bool tls) {
  switch (p) {
    case 20: return tls ? "HTTP_TLS" : "HTTP";
    case 21: return tls ? "HTTP2_TLS" : "HTTP2";
    case 40: return tls ? "Dubbo_TLS" : "Dubbo";
    case 41: return tls ? "gRPC_TLS" : "gRPC";
    case 43: return tls ? "SofaRPC_TLS" : "SofaRPC";
    case 44: return tls ? "FastCGI_TLS" : "FastCGI";
    case 45: return tls ? "bRPC_TLS" : "bRPC";
    case 46: return tls ? "Tars_TLS" : "Tars";
    case 47: return tls ? "SOME/IP_TLS" : "SOME/IP";
    case 48: return tls ? "ISO-8583_TLS" : "ISO-8583";
    case 49: return tls ? "TRIPLE_TLS" : "TRIPLE";
    case 50: return tls ? "NetSign_TLS" : "NetSign";
    case 60: return tls ? "MySQL_TLS" : "MySQL";
    case 61: return tls ? "PostgreSQL_TLS" : "PostgreSQL";
    case 62: return tls ? "Oracle_TLS" : "Oracle";
    case 63: return tls ? "Dameng_TLS" : "Dameng";
    case 80: return tls ? "Redis_TLS" : "Redis";
    case 81: return tls ? "MongoDB_TLS" : "MongoDB";
    case 82: return tls ? "Memcached_TLS" : "Memcached";
    case 100: return tls ? "Kafka_TLS" : "Kafka";
    case 101: return tls ? "MQTT_TLS" : "MQTT";
    case 102: return tls ? "AMQP_TLS" : "AMQP";
    case 103: return tls ? "OpenWire_TLS" : "OpenWire";
    case 104: return tls ? "NATS_TLS" : "NATS";
    case 105: return tls ? "Pulsar_TLS" : "Pulsar";
    case 106: return tls ? "ZMTP_TLS" : "ZMTP";
    case 107: return tls ? "RocketMQ_TLS" : "RocketMQ";
    case 108: return tls ? "WebSphereMQ_TLS" : "WebSphereMQ";
    case 120: return tls ? "DNS_TLS" : "DNS";
    case 121: return "TLS";
    case 127: return tls ? "Custom_TLS" : "Custom";
  }
  return "N/A";
}

static std::string_view tbl(const char* const* t, int n, int idx) {
  if (idx >= 0 && idx < n) return t[idx];
  return "";
}
// exception_desc.go getters :32, :265, :341, :373, :440, :447
static std::string_view GetHTTPExceptionDesc(uint16_t c) { return tbl(dfgpu_exc_http, DFGPU_EXC_HTTP_LEN, c); }
static std::string_view GetDubboExceptionDesc(uint16_t c) { return tbl(dfgpu_exc_dubbo, DFGPU_EXC_DUBBO_LEN, c); }
static std::string_view GetDNSExceptionDesc(uint16_t c) {
  if (c < DFGPU_EXC_DNS_LEN) return dfgpu_exc_dns[c];
  if (c >= 22 && c <= 3840) return "available for assignment";
  if (c >= 3841 && c <= 4095) return "Private Use";
  if (c >= 4096) return "available for assignment";
  return "";
}
static std::string_view GetKafkaExceptionDesc(int16_t c) {
  if (c == -1) return DFGPU_EXC_KAFKA_UNKNOWN_SERVER_ERROR;
  if (c > 0 && c < DFGPU_EXC_KAFKA_LEN) return dfgpu_exc_kafka[c];
  return "";
}
static std::string_view GetMQTTV3ExceptionDesc(uint16_t c) {
  if (c > 0 && c < DFGPU_EXC_MQTT_V3_LEN) return dfgpu_exc_mqtt_v3[c];
  return "";
}
static std::string_view GetMQTTV5ExceptionDesc(uint16_t c) {
  if (c > 0 && c < DFGPU_EXC_MQTT_V5_LEN) return dfgpu_exc_mqtt_v5[c];
  return "";
}
static bool IsKafkaSupportedCommand(std::string_view cmd) {  // l7_flow_log.go:454-461
  return cmd == "Fetch" || cmd == "Produce" || cmd == "JoinGroup" || cmd == "LeaveGroup" || cmd == "SyncGroup";
}

void L7Blocks::init() {
  Block& b = l7;
  auto U8 = [&](const char* n) { b.add(n, DFGPU_CH_UINT8, 1); };
  auto U16 = [&](const char* n) { b.add(n, DFGPU_CH_UINT16, 2); };
  auto U32 = [&](const char* n) { b.add(n, DFGPU_CH_UINT32, 4); };
  auto U64 = [&](const char* n) { b.add(n, DFGPU_CH_UINT64, 8); };
  auto STR = [&](const char* n) { b.add(n, DFGPU_CH_STR); };
  auto LCR = [&](const char* n) { b.add(n, DFGPU_CH_LC_STR_ROWS); };
  KnowledgeGraph::add_columns(b);  // l7_column_block.go:115
  // L7BaseBlock.ToInput :114-158
  b.add("time", DFGPU_CH_DATETIME, 4); b.add("ip4_0", DFGPU_CH_IPV4, 4); b.add("ip4_1", DFGPU_CH_IPV4, 4);
  b.add("ip6_0", DFGPU_CH_IPV6, 16); b.add("ip6_1", DFGPU_CH_IPV6, 16); U8("is_ipv4"); U8("protocol");
  U16("client_port"); U16("server_port"); U64("flow_id"); U8("capture_network_type_id");
  U8("nat_source"); U8("capture_nic_type"); U16("signal_source"); U8("tunnel_type"); U32("capture_nic");
  b.add("observation_point", DFGPU_CH_LC_STR, 4); U16("agent_id"); U32("req_tcp_seq");
  U32("resp_tcp_seq"); b.add("start_time", DFGPU_CH_DATETIME64, 8); b.add("end_time", DFGPU_CH_DATETIME64, 8);
  U32("gprocess_id_0"); U32("gprocess_id_1"); U8("biz_type"); STR("biz_code"); STR("biz_scenario");
  STR("biz_response_code"); b.add("process_id_0", DFGPU_CH_INT32, 4); b.add("process_id_1", DFGPU_CH_INT32, 4);
  STR("process_kname_0"); STR("process_kname_1"); U64("syscall_trace_id_request");
  U64("syscall_trace_id_response"); U32("syscall_thread_0"); U32("syscall_thread_1");
  U64("syscall_coroutine_0"); U64("syscall_coroutine_1"); U32("syscall_cap_seq_0"); U32("syscall_cap_seq_1");
  // L7FlowLogBlock.ToInput :304-350
  U64("_id"); U8("l7_protocol"); LCR("biz_protocol"); LCR("version"); U8("type"); U8("is_tls");
  U8("is_async"); U8("is_reversed"); LCR("request_type"); STR("request_domain"); STR("request_resource");
  STR("endpoint"); b.add("request_id", DFGPU_CH_NULLABLE_UINT64, 8); U8("response_status");
  b.add("response_code", DFGPU_CH_NULLABLE_INT32, 4); STR("response_exception"); STR("response_result");
  STR("http_proxy_client"); STR("x_request_id_0"); STR("x_request_id_1"); STR("trace_id");
  STR("_trace_id_2"); U64("trace_id_index"); STR("span_id"); STR("parent_span_id");
  b.add("span_kind", DFGPU_CH_NULLABLE_UINT8, 1); LCR("app_service"); LCR("app_instance");
  U64("response_duration"); b.add("request_length", DFGPU_CH_NULLABLE_INT64, 8);
  b.add("response_length", DFGPU_CH_NULLABLE_INT64, 8); b.add("sql_affected_rows", DFGPU_CH_NULLABLE_UINT64, 8);
  U8("direction_score"); U32("captured_request_byte"); U32("captured_response_byte");
  b.add("attribute_names", DFGPU_CH_ARR_LC_STR); b.add("attribute_values", DFGPU_CH_ARR_STR);
  b.add("metrics_names", DFGPU_CH_ARR_LC_STR); b.add("metrics_values", DFGPU_CH_ARR_FLOAT64, 8);
  STR("events");

  for (Block* t : {&field, &field_value}) {  // flow_tag_column_block.go:50-66
    t->add("time", DFGPU_CH_DATETIME, 4);
    t->add("table", DFGPU_CH_LC_STR, 4);
    t->add("vpc_id", DFGPU_CH_INT32, 4);
    t->add("pod_ns_id", DFGPU_CH_UINT16, 2);
    t->add("field_type", DFGPU_CH_LC_STR, 4);
    t->add("field_name", DFGPU_CH_LC_STR_ROWS);
    t->add("field_value_type", DFGPU_CH_LC_STR, 4);
    t->add("team_id", DFGPU_CH_UINT16, 2);
  }
  field_value.add("field_value", DFGPU_CH_STR);
  field_value.add("count", DFGPU_CH_UINT64, 8);

  app_service.add("time", DFGPU_CH_DATETIME, 4);  // app_service_tag_column_block.go:39-47
  app_service.add("table", DFGPU_CH_LC_STR, 4);
  app_service.add("app_service", DFGPU_CH_LC_STR_ROWS);
  app_service.add("app_instance", DFGPU_CH_LC_STR_ROWS);
  app_service.add("team_id", DFGPU_CH_UINT16, 2);
}

static const char* field_type_str(uint8_t t) {  // flow_tag.go:53-64
  return t == 0 ? "tag" : t == 1 ? "metrics" : t == 2 ? "custom_tag" : "invalid field type";
}

// FlowTag.AppendToColumnBlock flow_tag_column_block.go:79-101
static void append_flow_tag(Block& b, bool is_value, uint32_t ts, const FlowTagInfo& k) {
  // FieldValueType stays Auto on this path -> string, or float when the value is
  // empty and the field type is not tag (:82-88)
  const char* fvt = "string";
  if (k.FieldValue.empty() && k.FieldType != 0) fvt = "float";
  size_t c = 0;
  b[c++].append<uint32_t>(ts);
  b[c++].append_lc(k.Table);
  b[c++].append<int32_t>(k.VpcId);
  b[c++].append<uint16_t>(k.PodNsId);
  b[c++].append_lc(field_type_str(k.FieldType));
  b[c++].append_str(k.FieldName);
  b[c++].append_lc(fvt);
  b[c++].append<uint16_t>(k.TeamID);
  if (is_value) {
    b[c++].append_str(k.FieldValue);
    b[c++].append<uint64_t>(1);
  }
  b.nrows++;
}

void L7Blocks::append_record(const Msg& m, uint16_t orgId, uint16_t teamId,
                             const PlatformInfoTable* pd, L7State& st, uint32_t* id_counter,
                             uint64_t* lastTraceIdIndex, L7Counters& cnt) {
  using namespace fl;
  const Msg& l = m.sub(AppLog::base);
  const Msg& head = l.sub(Base::head);
  Block& b = l7;

  // ---- ProtoLogToL7FlowLog :638-644
  uint64_t _id = genID((uint32_t)(l.u64(Base::end_time) / 1000000000ull), id_counter, pd->QueryAnalyzerID());

  // ---- L7Base.Fill :514-597
  bool v6 = l.u32(Base::is_ipv6) == 1;
  uint32_t IP40 = 0, IP41 = 0;
  std::string_view IP60, IP61;
  if (v6) {
    IP60 = l.str(Base::ip6_src);
    IP61 = l.str(Base::ip6_dst);
  } else {
    IP40 = l.u32(Base::ip_src);
    IP41 = l.u32(Base::ip_dst);
  }
  uint64_t tapPort = l.u64(Base::tap_port);
  uint32_t TapPort = (uint32_t)tapPort;
  uint8_t TapPortType = (uint8_t)(tapPort >> 60);
  uint8_t NatSource = (uint8_t)((tapPort >> 36) & 0xf);
  uint8_t TunnelType = (uint8_t)((tapPort >> 32) & 0xf);
  uint16_t SignalSource = 0;               // SIGNAL_SOURCE_PACKET
  if (TapPortType == 8) SignalSource = 4;  // TAPPORT_FROM_OTEL -> SIGNAL_SOURCE_OTEL
  else if (TapPortType == 7) SignalSource = 3;  // TAPPORT_FROM_EBPF -> SIGNAL_SOURCE_EBPF
  bool tsok;
  std::string_view TapSide = TAPSideString(l.u32(Base::tap_side) & 0xff, tsok);
  int64_t StartTime = (int64_t)l.u64(Base::start_time) / 1000;
  int64_t EndTime = (int64_t)l.u64(Base::end_time) / 1000;
  uint32_t Time = (uint32_t)(l.u64(Base::end_time) / 1000000000ull);
  uint32_t GPID0 = l.u32(Base::gpid_0), GPID1 = l.u32(Base::gpid_1);
  uint8_t Protocol = (uint8_t)l.u32(Base::protocol);

  KnowledgeGraph kg;
  kg.OrgId = orgId;
  kg.TeamID = teamId;
  kg.fill(pd, v6, l.u32(Base::is_vip_interface_src) == 1, l.u32(Base::is_vip_interface_dst) == 1,
          l.i32(Base::l3_epc_id_src), l.i32(Base::l3_epc_id_dst), l.u32(Base::ip_src), l.u32(Base::ip_dst),
          l.str(Base::ip6_src), l.str(Base::ip6_dst), l.u64(Base::mac_src), l.u64(Base::mac_dst),
          l.u32(Base::gpid_0), l.u32(Base::gpid_1), (uint16_t)l.u32(Base::vtap_id), l.u32(Base::pod_id_0),
          l.u32(Base::pod_id_1), (uint16_t)l.u32(Base::port_dst), l.u32(Base::tap_side), Protocol);
  if (l.u32(Base::process_id_0) != 0 && l.u32(Base::gpid_0) == 0) {  // :589-592
    GPID0 = pd->QueryProcessInfo((uint16_t)l.u32(Base::vtap_id), l.u32(Base::process_id_0));
    kg.TagSource0 |= TS_ProcessId;
  }
  if (l.u32(Base::process_id_1) != 0 && l.u32(Base::gpid_1) == 0) {
    GPID1 = pd->QueryProcessInfo((uint16_t)l.u32(Base::vtap_id), l.u32(Base::process_id_1));
    kg.TagSource1 |= TS_ProcessId;
  }

  // ---- L7FlowLog.Fill :312-342
  uint8_t Type = (uint8_t)head.u32(Head::msg_type);
  uint32_t flags = m.u32(AppLog::flags);
  uint8_t IsTLS = (flags & 1) ? 1 : 0, IsAsync = (flags & 2) ? 1 : 0, IsReversed = (flags & 4) ? 1 : 0;
  uint8_t L7Protocol = (uint8_t)head.u32(Head::proto);
  bool hasReq = m.present(AppLog::req), hasResp = m.present(AppLog::resp);
  bool hasExt = m.present(AppLog::ext_info), hasTrace = m.present(AppLog::trace_info);
  const Msg& req = m.sub(AppLog::req);
  const Msg& resp = m.sub(AppLog::resp);
  const Msg& ext = m.sub(AppLog::ext_info);
  const Msg& trace = m.sub(AppLog::trace_info);
  std::string_view BizProtocol;
  if (hasExt && !ext.str(Ext::protocol_str).empty()) BizProtocol = ext.str(Ext::protocol_str);
  else BizProtocol = L7ProtocolString(L7Protocol, IsTLS == 1);
  uint8_t ResponseStatus = 5;  // STATUS_UNKNOWN
  uint64_t ResponseDuration = head.u64(Head::rrt) / 1000;

  // ---- fillL7FlowLog :363-452
  std::string_view Version = m.str(AppLog::version);
  int64_t requestLength = (int64_t)m.i32(AppLog::req_len);
  int64_t responseLength = (int64_t)m.i32(AppLog::resp_len);
  uint64_t sqlAffectedRows = (uint64_t)m.u32(AppLog::row_effect);
  const uint64_t* SqlAffectedRows = sqlAffectedRows != 0 ? &sqlAffectedRows : nullptr;
  const int64_t* RequestLength = nullptr;
  const int64_t* ResponseLength = nullptr;
  int32_t responseCode = 0;
  const int32_t* ResponseCode = nullptr;
  uint64_t requestId = 0;
  const uint64_t* RequestId = nullptr;
  std::string_view RequestDomain, RequestResource, RequestType, Endpoint, ResponseResult,
      ResponseException, AppService, XRequestId0, XRequestId1, HttpProxyClient, SpanId, ParentSpanId,
      TraceId, TraceId2;
  std::vector<std::string_view> AttributeNames, AttributeValues, MetricsNames;
  std::vector<double> MetricsValues;
  if (hasReq) {
    RequestDomain = req.str(Req::domain);
    RequestResource = req.str(Req::resource);
    RequestType = req.str(Req::req_type);
    if (requestLength != -1 && Type != 1 /*MSG_T_RESPONSE*/) RequestLength = &requestLength;
    Endpoint = req.str(Req::endpoint);
  }
  if (hasResp) {
    ResponseStatus = (uint8_t)resp.u32(Resp::status);
    if (Type != 0 /*MSG_T_REQUEST*/) {
      ResponseResult = resp.str(Resp::result);
      responseCode = resp.i32(Resp::code);
      ResponseException = resp.str(Resp::exception);
      if (ResponseException.empty()) {  // fillExceptionDesc :463-488
        if (ResponseStatus == 3 || ResponseStatus == 4) {
          int32_t code = resp.i32(Resp::code);
          switch (L7Protocol) {
            case 20: case 21: ResponseException = GetHTTPExceptionDesc((uint16_t)code); break;
            case 120: ResponseException = GetDNSExceptionDesc((uint16_t)code); break;
            case 40: ResponseException = GetDubboExceptionDesc((uint16_t)code); break;
            case 100: ResponseException = GetKafkaExceptionDesc((int16_t)code); break;
            case 101:
              ResponseException = (m.str(AppLog::version) != "5") ? GetMQTTV3ExceptionDesc((uint16_t)code)
                                                                   : GetMQTTV5ExceptionDesc((uint16_t)code);
              break;
            default: ResponseException = resp.str(Resp::exception); break;
          }
        }
      }
      if (responseCode != -32768) ResponseCode = &responseCode;
      if (responseLength != -1) ResponseLength = &responseLength;
    }
  }
  if (hasExt) {
    requestId = (uint64_t)ext.u32(Ext::request_id);
    if (requestId != 0) RequestId = &requestId;
    AppService = ext.str(Ext::service_name);
    XRequestId0 = ext.str(Ext::x_request_id_0);
    XRequestId1 = ext.str(Ext::x_request_id_1);
    HttpProxyClient = ext.str(Ext::client_ip);
    if (!ext.str(Ext::http_user_agent).empty()) {
      AttributeNames.push_back("http_user_agent");
      AttributeValues.push_back(ext.str(Ext::http_user_agent));
    }
    if (!ext.str(Ext::http_referer).empty()) {
      AttributeNames.push_back("http_referer");
      AttributeValues.push_back(ext.str(Ext::http_referer));
    }
    if (!ext.str(Ext::rpc_service).empty()) {
      AttributeNames.push_back("rpc_service");
      AttributeValues.push_back(ext.str(Ext::rpc_service));
    }
    for (auto& s : ext.rs[Ext::attribute_names]) AttributeNames.push_back(s);
    for (auto& s : ext.rs[Ext::attribute_values]) AttributeValues.push_back(s);
    for (auto& s : ext.rs[Ext::metrics_names]) MetricsNames.push_back(s);
    for (auto v : ext.rd[Ext::metrics_values]) MetricsValues.push_back(v);
  }
  if (hasTrace) {
    SpanId = trace.str(Trace::span_id);
    const auto& ids = trace.rs[Trace::trace_ids];  // fillTraceIds :344-360
    if (ids.size() > 0) TraceId = ids[0];
    if (ids.size() > 1) TraceId2 = ids[1];
    if (TraceId.empty()) TraceId = trace.str(Trace::trace_id);
    ParentSpanId = trace.str(Trace::parent_span_id);
  }
  uint64_t TraceIdIndex;  // ParseTraceIdIndex :296-310 (hash index type)
  if (st.cfg.trace_id_index_disabled) {
    TraceIdIndex = 0;
  } else if (TraceId.empty()) {
    TraceIdIndex = *lastTraceIdIndex + 1;
  } else {
    TraceIdIndex = DJBHash(17, TraceId);
    *lastTraceIdIndex = TraceIdIndex;
  }
  switch (L7Protocol) {  // :438-451
    case 100:
      if (hasReq) {
        if (responseCode == 0 && !IsKafkaSupportedCommand(req.str(Req::req_type))) {
          ResponseStatus = 2;  // STATUS_TIMEOUT
          ResponseCode = nullptr;
        }
        RequestId = &requestId;
      }
      break;
    case 43: RequestId = &requestId; break;
  }

  // ---- L7FlowLog.AppendToColumnBlock l7_column_block.go:376-422
  size_t c = 0;
  kg.append(b, c);
  c += 39;
  b[c++].append<uint32_t>(Time);
  b[c++].append<uint32_t>(IP40);
  b[c++].append<uint32_t>(IP41);
  b[c++].append_ip6((const uint8_t*)IP60.data(), IP60.size());
  b[c++].append_ip6((const uint8_t*)IP61.data(), IP61.size());
  b[c++].append<uint8_t>(v6 ? 0 : 1);
  b[c++].append<uint8_t>(Protocol);
  b[c++].append<uint16_t>((uint16_t)l.u32(Base::port_src));
  b[c++].append<uint16_t>((uint16_t)l.u32(Base::port_dst));
  b[c++].append<uint64_t>(l.u64(Base::flow_id));
  b[c++].append<uint8_t>((uint8_t)l.u32(Base::tap_type));
  b[c++].append<uint8_t>(NatSource);
  b[c++].append<uint8_t>(TapPortType);
  b[c++].append<uint16_t>(SignalSource);
  b[c++].append<uint8_t>(TunnelType);
  b[c++].append<uint32_t>(TapPort);
  b[c++].append_lc(TapSide);
  b[c++].append<uint16_t>((uint16_t)l.u32(Base::vtap_id));
  b[c++].append<uint32_t>(l.u32(Base::req_tcp_seq));
  b[c++].append<uint32_t>(l.u32(Base::resp_tcp_seq));
  b[c++].append<int64_t>(StartTime);
  b[c++].append<int64_t>(EndTime);
  b[c++].append<uint32_t>(GPID0);
  b[c++].append<uint32_t>(GPID1);
  b[c++].append<uint8_t>((uint8_t)l.u32(Base::biz_type));
  b[c++].append_str(m.str(AppLog::biz_code));
  b[c++].append_str(m.str(AppLog::biz_scenario));
  b[c++].append_str(m.str(AppLog::biz_response_code));
  b[c++].append<int32_t>((int32_t)l.u32(Base::process_id_0));
  b[c++].append<int32_t>((int32_t)l.u32(Base::process_id_1));
  b[c++].append_str(l.str(Base::process_kname_0));
  b[c++].append_str(l.str(Base::process_kname_1));
  b[c++].append<uint64_t>(l.u64(Base::syscall_trace_id_request));
  b[c++].append<uint64_t>(l.u64(Base::syscall_trace_id_response));
  b[c++].append<uint32_t>(l.u32(Base::syscall_trace_id_thread_0));
  b[c++].append<uint32_t>(l.u32(Base::syscall_trace_id_thread_1));
  b[c++].append<uint64_t>(l.u64(Base::syscall_coroutine_0));
  b[c++].append<uint64_t>(l.u64(Base::syscall_coroutine_1));
  b[c++].append<uint32_t>(l.u32(Base::syscall_cap_seq_0));
  b[c++].append<uint32_t>(l.u32(Base::syscall_cap_seq_1));
  b[c++].append<uint64_t>(_id);
  b[c++].append<uint8_t>(L7Protocol);
  b[c++].append_str(BizProtocol);
  b[c++].append_str(Version);
  b[c++].append<uint8_t>(Type);
  b[c++].append<uint8_t>(IsTLS);
  b[c++].append<uint8_t>(IsAsync);
  b[c++].append<uint8_t>(IsReversed);
  b[c++].append_str(RequestType);
  b[c++].append_str(RequestDomain);
  b[c++].append_str(RequestResource);
  b[c++].append_str(Endpoint);
  b[c++].append_nullable<uint64_t>(RequestId);
  b[c++].append<uint8_t>(ResponseStatus);
  b[c++].append_nullable<int32_t>(ResponseCode);
  b[c++].append_str(ResponseException);
  b[c++].append_str(ResponseResult);
  b[c++].append_str(HttpProxyClient);
  b[c++].append_str(XRequestId0);
  b[c++].append_str(XRequestId1);
  b[c++].append_str(TraceId);
  b[c++].append_str(TraceId2);
  b[c++].append<uint64_t>(TraceIdIndex);
  b[c++].append_str(SpanId);
  b[c++].append_str(ParentSpanId);
  b[c++].append_nullable<uint8_t>(nullptr);  // spanKind: only the OTel import sets it
  b[c++].append_str(AppService);
  b[c++].append_str("");                     // AppInstance: only the OTel import sets it
  b[c++].append<uint64_t>(ResponseDuration);
  b[c++].append_nullable<int64_t>(RequestLength);
  b[c++].append_nullable<int64_t>(ResponseLength);
  b[c++].append_nullable<uint64_t>(SqlAffectedRows);
  b[c++].append<uint8_t>((uint8_t)m.u32(AppLog::direction_score));
  b[c++].append<uint32_t>(m.u32(AppLog::captured_request_byte));
  b[c++].append<uint32_t>(m.u32(AppLog::captured_response_byte));
  b[c++].append_arr_str(AttributeNames);
  b[c++].append_arr_str(AttributeValues);
  b[c++].append_arr_str(MetricsNames);
  b[c++].append_arr<double>(MetricsValues);
  b[c++].append_str("");  // Events
  b.nrows++;

  // ---- GenerateNewFlowTags :648-745 (throttle disabled => every record is "sent")
  {
    int n_sides = 2;
    int32_t L3EpcIDs[2] = {kg.L3EpcID0, kg.L3EpcID1};
    uint16_t PodNSIDs[2] = {kg.PodNSID0, kg.PodNSID1};
    if (kg.L3EpcID0 == kg.L3EpcID1 && kg.PodNSID0 == kg.PodNSID1) n_sides = 1;
    uint32_t time = (uint32_t)(EndTime / 1000000);  // US_TO_S_DEVISOR
    std::vector<std::string_view> attributeNames = AttributeNames, attributeValues = AttributeValues;
    attributeNames.push_back("app_service");
    attributeNames.push_back("endpoint");
    attributeNames.push_back("app_instance");
    attributeValues.push_back(AppService);
    attributeValues.push_back(Endpoint);
    attributeValues.push_back("");
    size_t minNamesLen = attributeNames.size();
    if (attributeNames.size() != attributeValues.size() && attributeNames.size() > attributeValues.size())
      minNamesLen = attributeValues.size();
    uint32_t timeout = st.cfg.flow_tag_cache_flush_timeout;
    for (int idx = 0; idx < n_sides; idx++) {
      FlowTagInfo info;
      info.Table = "l7_flow_log";
      info.VpcId = L3EpcIDs[idx];
      info.PodNsId = PodNSIDs[idx];
      info.OrgId = orgId;
      info.TeamID = teamId;
      for (size_t i = 0; i < minNamesLen; i++) {
        info.FieldName = std::string(attributeNames[i]);
        info.FieldValue = std::string(attributeValues[i]);
        auto r = st.FieldValueCache.AddOrGet(info, time);
        if (r.second) {
          if ((uint32_t)(r.first + timeout) >= time) continue;
          st.FieldValueCache.Add(info, time);
        }
        append_flow_tag(field_value, true, time, info);
        cnt.new_field_value++;
        if (i >= AttributeNames.size()) continue;
        info.FieldValue.clear();
        auto r2 = st.FieldCache.AddOrGet(info, time);
        if (r2.second) {
          if ((uint32_t)(r2.first + timeout) >= time) continue;
          st.FieldCache.Add(info, time);
        }
        append_flow_tag(field, false, time, info);
        cnt.new_field++;
      }
      info.FieldType = 1;  // FieldMetrics
      info.FieldValue.clear();
      for (auto& name : MetricsNames) {
        info.FieldName = std::string(name);
        auto r = st.FieldCache.AddOrGet(info, time);
        if (r.second) {
          if ((uint32_t)(r.first + timeout) >= time) continue;
          st.FieldCache.Add(info, time);
        }
        append_flow_tag(field, false, time, info);
        cnt.new_field++;
      }
    }
  }
  // ---- appServiceTagWrite decoder.go:464-472 + AppServiceTagWriter.Write :92-113
  if (!AppService.empty()) {  // AppInstance is always "" on this path
    AppServiceTagKey k;
    k.Table = "l7_flow_log";
    k.AppService = std::string(AppService);
    k.OrgId = orgId;
    k.TeamID = teamId;
    bool emit = true;
    auto r = st.AppServiceCache.AddOrGet(k, Time);
    if (r.second) {
      if ((uint32_t)(r.first + st.cfg.flow_tag_cache_flush_timeout) >= Time) emit = false;
      else st.AppServiceCache.Add(k, Time);
    }
    if (emit) {
      size_t a = 0;
      app_service[a++].append<uint32_t>(Time);
      app_service[a++].append_lc("l7_flow_log");
      app_service[a++].append_str(AppService);
      app_service[a++].append_str("");
      app_service[a++].append<uint16_t>(teamId);
      app_service.nrows++;
    }
  }
  // ---- updateCounter decoder.go:498-525
  switch (L7Protocol) {
    case 20: case 21: cnt.http++; break;
    case 120: cnt.dns++; break;
    case 60: case 61: cnt.sql++; break;
    case 80: cnt.nosql++; break;
    case 40: cnt.rpc++; break;
    case 101: cnt.mq++; break;
  }
}

}  // namespace oracle
