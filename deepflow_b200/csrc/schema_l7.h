/*
 * schema_l7.h — column set of flow_log.l7_flow_log in ToInput order: the 39
 * KnowledgeGraph columns, the L7Base block, then the L7FlowLog block
 * (server/ingester/flow_log/log_data/l7_column_block.go:114-158,167-169,304-350).
 * The per-org NativeTagsBlock (libs/nativetag) is nil unless an operator configured
 * native tags and is out of scope.
 * X(enum id, "clickhouse name", ch type, fixed bytes per row).
 */
#pragma once
#include "schema_l4.h"

#define L7_COLUMNS(X)                                                                          \
  KG_COLUMNS(X)                                                                                \
  X(TIME, "time", DFGPU_CH_DATETIME, 4)                                                        \
  X(IP4_0, "ip4_0", DFGPU_CH_IPV4, 4) X(IP4_1, "ip4_1", DFGPU_CH_IPV4, 4)                     \
  X(IP6_0, "ip6_0", DFGPU_CH_IPV6, 16) X(IP6_1, "ip6_1", DFGPU_CH_IPV6, 16)                   \
  X(IS_IPV4, "is_ipv4", DFGPU_CH_UINT8, 1) X(PROTOCOL, "protocol", DFGPU_CH_UINT8, 1)         \
  X(CLIENT_PORT, "client_port", DFGPU_CH_UINT16, 2) X(SERVER_PORT, "server_port", DFGPU_CH_UINT16, 2) \
  X(FLOW_ID, "flow_id", DFGPU_CH_UINT64, 8)                                                    \
  X(CAPTURE_NETWORK_TYPE_ID, "capture_network_type_id", DFGPU_CH_UINT8, 1)                     \
  X(NAT_SOURCE, "nat_source", DFGPU_CH_UINT8, 1) X(CAPTURE_NIC_TYPE, "capture_nic_type", DFGPU_CH_UINT8, 1) \
  X(SIGNAL_SOURCE, "signal_source", DFGPU_CH_UINT16, 2) X(TUNNEL_TYPE, "tunnel_type", DFGPU_CH_UINT8, 1) \
  X(CAPTURE_NIC, "capture_nic", DFGPU_CH_UINT32, 4)                                            \
  X(OBSERVATION_POINT, "observation_point", DFGPU_CH_LC_STR, 1)                                \
  X(AGENT_ID, "agent_id", DFGPU_CH_UINT16, 2)                                                  \
  X(REQ_TCP_SEQ, "req_tcp_seq", DFGPU_CH_UINT32, 4) X(RESP_TCP_SEQ, "resp_tcp_seq", DFGPU_CH_UINT32, 4) \
  X(START_TIME, "start_time", DFGPU_CH_DATETIME64, 8) X(END_TIME, "end_time", DFGPU_CH_DATETIME64, 8) \
  X(GPROCESS_ID_0, "gprocess_id_0", DFGPU_CH_UINT32, 4) X(GPROCESS_ID_1, "gprocess_id_1", DFGPU_CH_UINT32, 4) \
  X(BIZ_TYPE, "biz_type", DFGPU_CH_UINT8, 1)                                                   \
  X(BIZ_CODE, "biz_code", DFGPU_CH_STR, 0) X(BIZ_SCENARIO, "biz_scenario", DFGPU_CH_STR, 0)   \
  X(BIZ_RESPONSE_CODE, "biz_response_code", DFGPU_CH_STR, 0)                                   \
  X(PROCESS_ID_0, "process_id_0", DFGPU_CH_INT32, 4) X(PROCESS_ID_1, "process_id_1", DFGPU_CH_INT32, 4) \
  X(PROCESS_KNAME_0, "process_kname_0", DFGPU_CH_STR, 0)                                       \
  X(PROCESS_KNAME_1, "process_kname_1", DFGPU_CH_STR, 0)                                       \
  X(SYSCALL_TRACE_ID_REQUEST, "syscall_trace_id_request", DFGPU_CH_UINT64, 8)                  \
  X(SYSCALL_TRACE_ID_RESPONSE, "syscall_trace_id_response", DFGPU_CH_UINT64, 8)                \
  X(SYSCALL_THREAD_0, "syscall_thread_0", DFGPU_CH_UINT32, 4)                                  \
  X(SYSCALL_THREAD_1, "syscall_thread_1", DFGPU_CH_UINT32, 4)                                  \
  X(SYSCALL_COROUTINE_0, "syscall_coroutine_0", DFGPU_CH_UINT64, 8)                            \
  X(SYSCALL_COROUTINE_1, "syscall_coroutine_1", DFGPU_CH_UINT64, 8)                            \
  X(SYSCALL_CAP_SEQ_0, "syscall_cap_seq_0", DFGPU_CH_UINT32, 4)                                \
  X(SYSCALL_CAP_SEQ_1, "syscall_cap_seq_1", DFGPU_CH_UINT32, 4)                                \
  X(_ID, "_id", DFGPU_CH_UINT64, 8) X(L7_PROTOCOL, "l7_protocol", DFGPU_CH_UINT8, 1)           \
  X(BIZ_PROTOCOL, "biz_protocol", DFGPU_CH_LC_STR_ROWS, 0)                                     \
  X(VERSION, "version", DFGPU_CH_LC_STR_ROWS, 0)                                               \
  X(TYPE, "type", DFGPU_CH_UINT8, 1) X(IS_TLS, "is_tls", DFGPU_CH_UINT8, 1)                    \
  X(IS_ASYNC, "is_async", DFGPU_CH_UINT8, 1) X(IS_REVERSED, "is_reversed", DFGPU_CH_UINT8, 1)  \
  X(REQUEST_TYPE, "request_type", DFGPU_CH_LC_STR_ROWS, 0)                                     \
  X(REQUEST_DOMAIN, "request_domain", DFGPU_CH_STR, 0)                                         \
  X(REQUEST_RESOURCE, "request_resource", DFGPU_CH_STR, 0)                                     \
  X(ENDPOINT, "endpoint", DFGPU_CH_STR, 0)                                                     \
  X(REQUEST_ID, "request_id", DFGPU_CH_NULLABLE_UINT64, 8)                                     \
  X(RESPONSE_STATUS, "response_status", DFGPU_CH_UINT8, 1)                                     \
  X(RESPONSE_CODE, "response_code", DFGPU_CH_NULLABLE_INT32, 4)                                \
  X(RESPONSE_EXCEPTION, "response_exception", DFGPU_CH_STR, 0)                                 \
  X(RESPONSE_RESULT, "response_result", DFGPU_CH_STR, 0)                                       \
  X(HTTP_PROXY_CLIENT, "http_proxy_client", DFGPU_CH_STR, 0)                                   \
  X(X_REQUEST_ID_0, "x_request_id_0", DFGPU_CH_STR, 0) X(X_REQUEST_ID_1, "x_request_id_1", DFGPU_CH_STR, 0) \
  X(TRACE_ID, "trace_id", DFGPU_CH_STR, 0) X(TRACE_ID_2, "_trace_id_2", DFGPU_CH_STR, 0)       \
  X(TRACE_ID_INDEX, "trace_id_index", DFGPU_CH_UINT64, 8)                                      \
  X(SPAN_ID, "span_id", DFGPU_CH_STR, 0) X(PARENT_SPAN_ID, "parent_span_id", DFGPU_CH_STR, 0)  \
  X(SPAN_KIND, "span_kind", DFGPU_CH_NULLABLE_UINT8, 1)                                        \
  X(APP_SERVICE, "app_service", DFGPU_CH_LC_STR_ROWS, 0)                                       \
  X(APP_INSTANCE, "app_instance", DFGPU_CH_LC_STR_ROWS, 0)                                     \
  X(RESPONSE_DURATION, "response_duration", DFGPU_CH_UINT64, 8)                                \
  X(REQUEST_LENGTH, "request_length", DFGPU_CH_NULLABLE_INT64, 8)                              \
  X(RESPONSE_LENGTH, "response_length", DFGPU_CH_NULLABLE_INT64, 8)                            \
  X(SQL_AFFECTED_ROWS, "sql_affected_rows", DFGPU_CH_NULLABLE_UINT64, 8)                       \
  X(DIRECTION_SCORE, "direction_score", DFGPU_CH_UINT8, 1)                                     \
  X(CAPTURED_REQUEST_BYTE, "captured_request_byte", DFGPU_CH_UINT32, 4)                        \
  X(CAPTURED_RESPONSE_BYTE, "captured_response_byte", DFGPU_CH_UINT32, 4)                      \
  X(ATTRIBUTE_NAMES, "attribute_names", DFGPU_CH_ARR_LC_STR, 0)                                \
  X(ATTRIBUTE_VALUES, "attribute_values", DFGPU_CH_ARR_STR, 0)                                 \
  X(METRICS_NAMES, "metrics_names", DFGPU_CH_ARR_LC_STR, 0)                                    \
  X(METRICS_VALUES, "metrics_values", DFGPU_CH_ARR_FLOAT64, 8)                                 \
  X(EVENTS, "events", DFGPU_CH_STR, 0)

namespace dfgpu {
#define X(id, name, t, es) L7C_##id,
enum L7Col : int { L7_COLUMNS(X) L7C_COUNT };
#undef X

// variable-length streams of the L7 table, in column order (= stream index)
enum L7Stream : int {
  L7V_BIZ_CODE = 0, L7V_BIZ_SCENARIO, L7V_BIZ_RESPONSE_CODE, L7V_KNAME_0, L7V_KNAME_1, L7V_BIZ_PROTOCOL,
  L7V_VERSION, L7V_REQUEST_TYPE, L7V_REQUEST_DOMAIN, L7V_REQUEST_RESOURCE, L7V_ENDPOINT,
  L7V_RESPONSE_EXCEPTION, L7V_RESPONSE_RESULT, L7V_HTTP_PROXY_CLIENT, L7V_X_REQUEST_ID_0,
  L7V_X_REQUEST_ID_1, L7V_TRACE_ID, L7V_TRACE_ID_2, L7V_SPAN_ID, L7V_PARENT_SPAN_ID, L7V_APP_SERVICE,
  L7V_APP_INSTANCE, L7V_ATTR_NAMES, L7V_ATTR_VALUES, L7V_METRICS_NAMES, L7V_METRICS_VALUES, L7V_EVENTS,
  L7V_COUNT
};
// scan slots: 0 rows, 1..22 the byte counts of streams 0..21, then the array streams
enum {
  L7S_ROWS = 0, L7S_STR0 = 1, L7S_AN_ELEMS = 23, L7S_AN_BYTES = 24, L7S_AV_ELEMS = 25, L7S_AV_BYTES = 26,
  L7S_MN_ELEMS = 27, L7S_MN_BYTES = 28, L7S_MV_ELEMS = 29, L7S_COUNT = 30,
  // trace-id-index carry (not scanned as sums): per-warp flag + hash of the last
  // valid record with a non-empty trace id
  L7S_TRACE_HAS = 30, L7S_TRACE_LO = 31, L7S_TRACE_HI = 32
};
enum { LC7_TAPSIDE = 0 };
}  // namespace dfgpu
