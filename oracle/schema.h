/*
 * ORACLE — test infrastructure only.
 * schema.h: message descriptors transcribed from message/flow_log.proto:14-310
 * (TaggedFlow, AppProtoLogsData) and message/metric.proto (Document), plus named
 * field numbers.  Field numbers are the wire contract.
 */
#pragma once
#include "pbwire.h"

namespace oracle {
namespace fl {  // flow_log.proto field numbers

namespace TaggedFlow { enum { flow = 1 }; }
namespace Flow {
enum {
  flow_key = 1, metrics_peer_src = 2, metrics_peer_dst = 3, tunnel = 4, flow_id = 5,
  start_time = 6, end_time = 7, duration = 8, vlan = 10, eth_type = 11,
  has_perf_stats = 12, perf_stats = 13, close_type = 14, signal_source = 15,
  is_active_service = 16, queue_hash = 17, is_new_flow = 18, tap_side = 19,
  syn_seq = 20, synack_seq = 21, last_keepalive_seq = 22, last_keepalive_ack = 23,
  acl_gids = 24, direction_score = 25, request_domain = 26, aggregated_flow_ids = 27,
  init_ipid = 28
};
}
namespace FlowKey {
enum {
  vtap_id = 1, tap_type = 2, tap_port = 3, mac_src = 4, mac_dst = 5, ip_src = 6,
  ip_dst = 7, ip6_src = 8, ip6_dst = 9, port_src = 10, port_dst = 11, proto = 12
};
}
namespace Peer {  // FlowMetricsPeer
enum {
  byte_count = 1, l3_byte_count = 2, l4_byte_count = 3, packet_count = 4,
  total_byte_count = 5, total_packet_count = 6, first = 7, last = 8, tcp_flags = 9,
  l3_epc_id = 10, is_l2_end = 11, is_l3_end = 12, is_active_host = 13, is_device = 14,
  is_vip_interface = 15, is_vip = 16, real_ip = 20, real_port = 21, gpid = 22
};
}
namespace Tunnel {
enum {
  tx_ip0 = 1, tx_ip1 = 2, rx_ip0 = 3, rx_ip1 = 4, tx_mac0 = 5, tx_mac1 = 6, rx_mac0 = 7,
  rx_mac1 = 8, tx_id = 9, rx_id = 10, tunnel_type = 11, tier = 12, is_ipv6 = 13
};
}
namespace PerfStats { enum { tcp = 1, l7 = 2, l4_protocol = 3, l7_protocol = 4, l7_failed_count = 5 }; }
namespace Tcp {
enum {
  rtt_client_max = 1, rtt_server_max = 2, srt_max = 3, art_max = 4, rtt = 5, srt_sum = 8,
  art_sum = 9, srt_count = 12, art_count = 13, counts_peer_tx = 14, counts_peer_rx = 15,
  total_retrans_count = 16, syn_count = 17, synack_count = 18, cit_max = 19, cit_sum = 20,
  cit_count = 21, fin_count = 22
};
}
namespace Counts { enum { retrans_count = 1, zero_win_count = 2, ooo_count = 3 }; }
namespace L7Perf {
enum {
  request_count = 1, response_count = 2, err_client_count = 3, err_server_count = 4,
  err_timeout = 5, rrt_count = 6, rrt_sum = 7, rrt_max = 8, tls_rtt = 9
};
}

namespace AppLog {  // AppProtoLogsData
enum {
  base = 1, req_len = 9, resp_len = 10, req = 11, resp = 12, version = 13, trace_info = 14,
  ext_info = 15, row_effect = 16, direction_score = 17, flags = 18,
  captured_request_byte = 19, captured_response_byte = 20, biz_code = 21,
  biz_scenario = 22, biz_response_code = 23
};
}
namespace Base {  // AppProtoLogsBaseInfo
enum {
  start_time = 1, end_time = 2, flow_id = 3, tap_port = 4, vtap_id = 5, tap_type = 6,
  is_ipv6 = 7, tap_side = 8, head = 9, mac_src = 10, mac_dst = 11, ip_src = 12, ip_dst = 13,
  ip6_src = 14, ip6_dst = 15, l3_epc_id_src = 16, l3_epc_id_dst = 17, port_src = 18,
  port_dst = 19, protocol = 20, is_vip_interface_src = 21, is_vip_interface_dst = 22,
  req_tcp_seq = 23, resp_tcp_seq = 24, process_id_0 = 25, process_id_1 = 26,
  process_kname_0 = 27, process_kname_1 = 28, syscall_trace_id_request = 29,
  syscall_trace_id_response = 30, syscall_trace_id_thread_0 = 31,
  syscall_trace_id_thread_1 = 32, syscall_cap_seq_0 = 33, syscall_cap_seq_1 = 34,
  gpid_0 = 35, gpid_1 = 36, syscall_coroutine_0 = 39, syscall_coroutine_1 = 40,
  pod_id_0 = 41, pod_id_1 = 42, biz_type = 43
};
}
namespace Head { enum { proto = 1, msg_type = 2, rrt = 5 }; }
namespace Req { enum { req_type = 1, domain = 2, resource = 3, endpoint = 4 }; }
namespace Resp { enum { status = 1, code = 2, exception = 3, result = 4 }; }
namespace Trace { enum { trace_id = 1, span_id = 2, parent_span_id = 3, trace_ids = 4 }; }
namespace Ext {
enum {
  service_name = 1, client_ip = 2, request_id = 3, x_request_id_0 = 4, http_user_agent = 6,
  http_referer = 7, rpc_service = 8, protocol_str = 9, x_request_id_1 = 10,
  attribute_names = 16, attribute_values = 17, metrics_names = 18, metrics_values = 19
};
}
}  // namespace fl

struct Schema {
  MsgDesc flow_key{"FlowKey",
                   {{1, K_U32}, {2, K_U32}, {3, K_U64}, {4, K_U64}, {5, K_U64}, {6, K_U32},
                    {7, K_U32}, {8, K_STR}, {9, K_STR}, {10, K_U32}, {11, K_U32}, {12, K_U32}}};
  MsgDesc peer{"FlowMetricsPeer",
               {{1, K_U64}, {2, K_U64}, {3, K_U64}, {4, K_U64}, {5, K_U64}, {6, K_U64},
                {7, K_U64}, {8, K_U64}, {9, K_U32}, {10, K_I32}, {11, K_U32}, {12, K_U32},
                {13, K_U32}, {14, K_U32}, {15, K_U32}, {16, K_U32}, {20, K_U32}, {21, K_U32},
                {22, K_U32}}};
  MsgDesc tunnel{"TunnelField",
                 {{1, K_U32}, {2, K_U32}, {3, K_U32}, {4, K_U32}, {5, K_U32}, {6, K_U32},
                  {7, K_U32}, {8, K_U32}, {9, K_U32}, {10, K_U32}, {11, K_U32}, {12, K_U32},
                  {13, K_U32}}};
  MsgDesc counts{"TcpPerfCountsPeer", {{1, K_U32}, {2, K_U32}, {3, K_U32}}};
  MsgDesc tcp{"TCPPerfStats",
              {{1, K_U32}, {2, K_U32}, {3, K_U32}, {4, K_U32}, {5, K_U32}, {8, K_U32},
               {9, K_U32}, {12, K_U32}, {13, K_U32}, {14, K_MSG, &counts}, {15, K_MSG, &counts},
               {16, K_U32}, {17, K_U32}, {18, K_U32}, {19, K_U32}, {20, K_U32}, {21, K_U32},
               {22, K_U32}}};
  MsgDesc l7perf{"L7PerfStats",
                 {{1, K_U32}, {2, K_U32}, {3, K_U32}, {4, K_U32}, {5, K_U32}, {6, K_U32},
                  {7, K_U64}, {8, K_U32}, {9, K_U32}}};
  MsgDesc perf{"FlowPerfStats",
               {{1, K_MSG, &tcp}, {2, K_MSG, &l7perf}, {3, K_U32}, {4, K_U32}, {5, K_U32}}};
  MsgDesc flow{"Flow",
               {{1, K_MSG, &flow_key}, {2, K_MSG, &peer}, {3, K_MSG, &peer}, {4, K_MSG, &tunnel},
                {5, K_U64}, {6, K_U64}, {7, K_U64}, {8, K_U64}, {10, K_U32}, {11, K_U32},
                {12, K_U32}, {13, K_MSG, &perf}, {14, K_U32}, {15, K_U32}, {16, K_U32},
                {17, K_U32}, {18, K_U32}, {19, K_U32}, {20, K_U32}, {21, K_U32}, {22, K_U32},
                {23, K_U32}, {24, K_RU32}, {25, K_U32}, {26, K_STR}, {27, K_RU64}, {28, K_U32}}};
  MsgDesc tagged_flow{"TaggedFlow", {{1, K_MSG, &flow}}};

  MsgDesc head{"AppProtoHead", {{1, K_U32}, {2, K_U32}, {5, K_U64}}};
  MsgDesc base{"AppProtoLogsBaseInfo",
               {{1, K_U64}, {2, K_U64}, {3, K_U64}, {4, K_U64}, {5, K_U32}, {6, K_U32},
                {7, K_U32}, {8, K_U32}, {9, K_MSG, &head}, {10, K_U64}, {11, K_U64},
                {12, K_U32}, {13, K_U32}, {14, K_STR}, {15, K_STR}, {16, K_I32}, {17, K_I32},
                {18, K_U32}, {19, K_U32}, {20, K_U32}, {21, K_U32}, {22, K_U32}, {23, K_U32},
                {24, K_U32}, {25, K_U32}, {26, K_U32}, {27, K_STR}, {28, K_STR}, {29, K_U64},
                {30, K_U64}, {31, K_U32}, {32, K_U32}, {33, K_U32}, {34, K_U32}, {35, K_U32},
                {36, K_U32}, {39, K_U64}, {40, K_U64}, {41, K_U32}, {42, K_U32}, {43, K_U32}}};
  MsgDesc req{"L7Request", {{1, K_STR}, {2, K_STR}, {3, K_STR}, {4, K_STR}}};
  MsgDesc resp{"L7Response", {{1, K_U32}, {2, K_I32}, {3, K_STR}, {4, K_STR}}};
  MsgDesc trace{"TraceInfo", {{1, K_STR}, {2, K_STR}, {3, K_STR}, {4, K_RSTR}}};
  MsgDesc ext{"ExtendedInfo",
              {{1, K_STR}, {2, K_STR}, {3, K_U32}, {4, K_STR}, {6, K_STR}, {7, K_STR},
               {8, K_STR}, {9, K_STR}, {10, K_STR}, {16, K_RSTR}, {17, K_RSTR}, {18, K_RSTR},
               {19, K_RDBL}}};
  MsgDesc app_log{"AppProtoLogsData",
                  {{1, K_MSG, &base}, {9, K_I32}, {10, K_I32}, {11, K_MSG, &req},
                   {12, K_MSG, &resp}, {13, K_STR}, {14, K_MSG, &trace}, {15, K_MSG, &ext},
                   {16, K_U32}, {17, K_U32}, {18, K_U32}, {19, K_U32}, {20, K_U32}, {21, K_STR},
                   {22, K_STR}, {23, K_STR}}};
};

inline const Schema& schema() {
  static Schema s;
  return s;
}

}  // namespace oracle
