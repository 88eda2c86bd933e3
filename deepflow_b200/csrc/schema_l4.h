/*
 * schema_l4.h — column set of flow_log.l4_flow_log in ToInput order
 * (server/ingester/flow_log/log_data/l4_column_block.go:822-833 and the
 * sub-block ToInput functions :40-50,113-140,196-208,233-236,258-262,362-403,
 * 519-553,693-742).  X(enum id, "clickhouse name", ch type, elem bytes).
 */
#pragma once

#define KG_COLUMNS(X)                                                                          \
  X(REGION_ID_0, "region_id_0", DFGPU_CH_UINT16, 2) X(REGION_ID_1, "region_id_1", DFGPU_CH_UINT16, 2) \
  X(AZ_ID_0, "az_id_0", DFGPU_CH_UINT16, 2) X(AZ_ID_1, "az_id_1", DFGPU_CH_UINT16, 2)         \
  X(HOST_ID_0, "host_id_0", DFGPU_CH_UINT16, 2) X(HOST_ID_1, "host_id_1", DFGPU_CH_UINT16, 2) \
  X(L3_DEVICE_TYPE_0, "l3_device_type_0", DFGPU_CH_UINT8, 1)                                   \
  X(L3_DEVICE_TYPE_1, "l3_device_type_1", DFGPU_CH_UINT8, 1)                                   \
  X(L3_DEVICE_ID_0, "l3_device_id_0", DFGPU_CH_UINT32, 4)                                      \
  X(L3_DEVICE_ID_1, "l3_device_id_1", DFGPU_CH_UINT32, 4)                                      \
  X(POD_NODE_ID_0, "pod_node_id_0", DFGPU_CH_UINT32, 4) X(POD_NODE_ID_1, "pod_node_id_1", DFGPU_CH_UINT32, 4) \
  X(POD_NS_ID_0, "pod_ns_id_0", DFGPU_CH_UINT16, 2) X(POD_NS_ID_1, "pod_ns_id_1", DFGPU_CH_UINT16, 2) \
  X(POD_GROUP_ID_0, "pod_group_id_0", DFGPU_CH_UINT32, 4)                                      \
  X(POD_GROUP_ID_1, "pod_group_id_1", DFGPU_CH_UINT32, 4)                                      \
  X(POD_ID_0, "pod_id_0", DFGPU_CH_UINT32, 4) X(POD_ID_1, "pod_id_1", DFGPU_CH_UINT32, 4)     \
  X(POD_CLUSTER_ID_0, "pod_cluster_id_0", DFGPU_CH_UINT16, 2)                                  \
  X(POD_CLUSTER_ID_1, "pod_cluster_id_1", DFGPU_CH_UINT16, 2)                                  \
  X(L3_EPC_ID_0, "l3_epc_id_0", DFGPU_CH_INT32, 4) X(L3_EPC_ID_1, "l3_epc_id_1", DFGPU_CH_INT32, 4) \
  X(EPC_ID_0, "epc_id_0", DFGPU_CH_INT32, 4) X(EPC_ID_1, "epc_id_1", DFGPU_CH_INT32, 4)       \
  X(SUBNET_ID_0, "subnet_id_0", DFGPU_CH_UINT16, 2) X(SUBNET_ID_1, "subnet_id_1", DFGPU_CH_UINT16, 2) \
  X(SERVICE_ID_0, "service_id_0", DFGPU_CH_UINT32, 4) X(SERVICE_ID_1, "service_id_1", DFGPU_CH_UINT32, 4) \
  X(AUTO_INSTANCE_ID_0, "auto_instance_id_0", DFGPU_CH_UINT32, 4)                              \
  X(AUTO_INSTANCE_TYPE_0, "auto_instance_type_0", DFGPU_CH_UINT8, 1)                           \
  X(AUTO_SERVICE_ID_0, "auto_service_id_0", DFGPU_CH_UINT32, 4)                                \
  X(AUTO_SERVICE_TYPE_0, "auto_service_type_0", DFGPU_CH_UINT8, 1)                             \
  X(AUTO_INSTANCE_ID_1, "auto_instance_id_1", DFGPU_CH_UINT32, 4)                              \
  X(AUTO_INSTANCE_TYPE_1, "auto_instance_type_1", DFGPU_CH_UINT8, 1)                           \
  X(AUTO_SERVICE_ID_1, "auto_service_id_1", DFGPU_CH_UINT32, 4)                                \
  X(AUTO_SERVICE_TYPE_1, "auto_service_type_1", DFGPU_CH_UINT8, 1)                             \
  X(TAG_SOURCE_0, "tag_source_0", DFGPU_CH_UINT8, 1) X(TAG_SOURCE_1, "tag_source_1", DFGPU_CH_UINT8, 1) \
  X(TEAM_ID, "team_id", DFGPU_CH_UINT16, 2)

#define L4_COLUMNS(X)                                                                          \
  X(_ID, "_id", DFGPU_CH_UINT64, 8)                                                            \
  X(MAC_0, "mac_0", DFGPU_CH_UINT64, 8) X(MAC_1, "mac_1", DFGPU_CH_UINT64, 8)                 \
  X(ETH_TYPE, "eth_type", DFGPU_CH_UINT16, 2) X(VLAN, "vlan", DFGPU_CH_UINT16, 2)             \
  X(IP4_0, "ip4_0", DFGPU_CH_IPV4, 4) X(IP4_1, "ip4_1", DFGPU_CH_IPV4, 4)                     \
  X(IP6_0, "ip6_0", DFGPU_CH_IPV6, 16) X(IP6_1, "ip6_1", DFGPU_CH_IPV6, 16)                   \
  X(IS_IPV4, "is_ipv4", DFGPU_CH_UINT8, 1) X(PROTOCOL, "protocol", DFGPU_CH_UINT8, 1)         \
  X(TUNNEL_TIER, "tunnel_tier", DFGPU_CH_UINT8, 1) X(TUNNEL_TYPE, "tunnel_type", DFGPU_CH_UINT16, 2) \
  X(TUNNEL_TX_ID, "tunnel_tx_id", DFGPU_CH_UINT32, 4) X(TUNNEL_RX_ID, "tunnel_rx_id", DFGPU_CH_UINT32, 4) \
  X(TUNNEL_TX_IP4_0, "tunnel_tx_ip4_0", DFGPU_CH_IPV4, 4) X(TUNNEL_TX_IP4_1, "tunnel_tx_ip4_1", DFGPU_CH_IPV4, 4) \
  X(TUNNEL_RX_IP4_0, "tunnel_rx_ip4_0", DFGPU_CH_IPV4, 4) X(TUNNEL_RX_IP4_1, "tunnel_rx_ip4_1", DFGPU_CH_IPV4, 4) \
  X(TUNNEL_TX_IP6_0, "tunnel_tx_ip6_0", DFGPU_CH_IPV6, 16) X(TUNNEL_TX_IP6_1, "tunnel_tx_ip6_1", DFGPU_CH_IPV6, 16) \
  X(TUNNEL_RX_IP6_0, "tunnel_rx_ip6_0", DFGPU_CH_IPV6, 16) X(TUNNEL_RX_IP6_1, "tunnel_rx_ip6_1", DFGPU_CH_IPV6, 16) \
  X(TUNNEL_IS_IPV4, "tunnel_is_ipv4", DFGPU_CH_UINT8, 1)                                       \
  X(TUNNEL_TX_MAC_0, "tunnel_tx_mac_0", DFGPU_CH_UINT32, 4) X(TUNNEL_TX_MAC_1, "tunnel_tx_mac_1", DFGPU_CH_UINT32, 4) \
  X(TUNNEL_RX_MAC_0, "tunnel_rx_mac_0", DFGPU_CH_UINT32, 4) X(TUNNEL_RX_MAC_1, "tunnel_rx_mac_1", DFGPU_CH_UINT32, 4) \
  X(CLIENT_PORT, "client_port", DFGPU_CH_UINT16, 2) X(SERVER_PORT, "server_port", DFGPU_CH_UINT16, 2) \
  X(TCP_FLAGS_BIT_0, "tcp_flags_bit_0", DFGPU_CH_UINT16, 2) X(TCP_FLAGS_BIT_1, "tcp_flags_bit_1", DFGPU_CH_UINT16, 2) \
  X(SYN_SEQ, "syn_seq", DFGPU_CH_UINT32, 4) X(SYN_ACK_SEQ, "syn_ack_seq", DFGPU_CH_UINT32, 4) \
  X(LAST_KEEPALIVE_SEQ, "last_keepalive_seq", DFGPU_CH_UINT32, 4)                              \
  X(LAST_KEEPALIVE_ACK, "last_keepalive_ack", DFGPU_CH_UINT32, 4)                              \
  X(L7_PROTOCOL, "l7_protocol", DFGPU_CH_UINT8, 1)                                             \
  X(PROVINCE_0, "province_0", DFGPU_CH_LC_STR, 1) X(PROVINCE_1, "province_1", DFGPU_CH_LC_STR, 1) \
  KG_COLUMNS(X)                                                                                \
  X(TIME, "time", DFGPU_CH_DATETIME, 4) X(CLOSE_TYPE, "close_type", DFGPU_CH_UINT16, 2)       \
  X(SIGNAL_SOURCE, "signal_source", DFGPU_CH_UINT16, 2) X(FLOW_ID, "flow_id", DFGPU_CH_UINT64, 8) \
  X(AGGREGATED_FLOW_IDS, "aggregated_flow_ids", DFGPU_CH_STR, 0)                               \
  X(INIT_IPID, "init_ipid", DFGPU_CH_UINT32, 4)                                                \
  X(CAPTURE_NETWORK_TYPE_ID, "capture_network_type_id", DFGPU_CH_UINT8, 1)                     \
  X(NAT_SOURCE, "nat_source", DFGPU_CH_UINT8, 1) X(CAPTURE_NIC_TYPE, "capture_nic_type", DFGPU_CH_UINT8, 1) \
  X(CAPTURE_NIC, "capture_nic", DFGPU_CH_UINT32, 4)                                            \
  X(OBSERVATION_POINT, "observation_point", DFGPU_CH_LC_STR, 1)                                \
  X(AGENT_ID, "agent_id", DFGPU_CH_UINT16, 2)                                                  \
  X(L2_END_0, "l2_end_0", DFGPU_CH_UINT8, 1) X(L2_END_1, "l2_end_1", DFGPU_CH_UINT8, 1)       \
  X(L3_END_0, "l3_end_0", DFGPU_CH_UINT8, 1) X(L3_END_1, "l3_end_1", DFGPU_CH_UINT8, 1)       \
  X(START_TIME, "start_time", DFGPU_CH_DATETIME64, 8) X(END_TIME, "end_time", DFGPU_CH_DATETIME64, 8) \
  X(DURATION, "duration", DFGPU_CH_UINT64, 8) X(IS_NEW_FLOW, "is_new_flow", DFGPU_CH_UINT8, 1) \
  X(STATUS, "status", DFGPU_CH_UINT8, 1) X(ACL_GIDS, "acl_gids", DFGPU_CH_ARR_UINT16, 2)      \
  X(GPROCESS_ID_0, "gprocess_id_0", DFGPU_CH_UINT32, 4) X(GPROCESS_ID_1, "gprocess_id_1", DFGPU_CH_UINT32, 4) \
  X(NAT_REAL_IP4_0, "nat_real_ip4_0", DFGPU_CH_IPV4, 4) X(NAT_REAL_IP4_1, "nat_real_ip4_1", DFGPU_CH_IPV4, 4) \
  X(NAT_REAL_PORT_0, "nat_real_port_0", DFGPU_CH_UINT16, 2) X(NAT_REAL_PORT_1, "nat_real_port_1", DFGPU_CH_UINT16, 2) \
  X(DIRECTION_SCORE, "direction_score", DFGPU_CH_UINT8, 1)                                     \
  X(REQUEST_DOMAIN, "request_domain", DFGPU_CH_STR, 0)                                         \
  X(PACKET_TX, "packet_tx", DFGPU_CH_UINT64, 8) X(PACKET_RX, "packet_rx", DFGPU_CH_UINT64, 8) \
  X(BYTE_TX, "byte_tx", DFGPU_CH_UINT64, 8) X(BYTE_RX, "byte_rx", DFGPU_CH_UINT64, 8)         \
  X(L3_BYTE_TX, "l3_byte_tx", DFGPU_CH_UINT64, 8) X(L3_BYTE_RX, "l3_byte_rx", DFGPU_CH_UINT64, 8) \
  X(L4_BYTE_TX, "l4_byte_tx", DFGPU_CH_UINT64, 8) X(L4_BYTE_RX, "l4_byte_rx", DFGPU_CH_UINT64, 8) \
  X(TOTAL_PACKET_TX, "total_packet_tx", DFGPU_CH_UINT64, 8) X(TOTAL_PACKET_RX, "total_packet_rx", DFGPU_CH_UINT64, 8) \
  X(TOTAL_BYTE_TX, "total_byte_tx", DFGPU_CH_UINT64, 8) X(TOTAL_BYTE_RX, "total_byte_rx", DFGPU_CH_UINT64, 8) \
  X(L7_REQUEST, "l7_request", DFGPU_CH_UINT32, 4) X(L7_RESPONSE, "l7_response", DFGPU_CH_UINT32, 4) \
  X(L7_PARSE_FAILED, "l7_parse_failed", DFGPU_CH_UINT32, 4)                                    \
  X(RTT, "rtt", DFGPU_CH_FLOAT64, 8) X(RTT_CLIENT, "rtt_client", DFGPU_CH_FLOAT64, 8)         \
  X(RTT_SERVER, "rtt_server", DFGPU_CH_FLOAT64, 8) X(TLS_RTT, "tls_rtt", DFGPU_CH_FLOAT64, 8) \
  X(SRT_SUM, "srt_sum", DFGPU_CH_FLOAT64, 8) X(ART_SUM, "art_sum", DFGPU_CH_FLOAT64, 8)       \
  X(RRT_SUM, "rrt_sum", DFGPU_CH_FLOAT64, 8) X(CIT_SUM, "cit_sum", DFGPU_CH_FLOAT64, 8)       \
  X(SRT_COUNT, "srt_count", DFGPU_CH_UINT64, 8) X(ART_COUNT, "art_count", DFGPU_CH_UINT64, 8) \
  X(RRT_COUNT, "rrt_count", DFGPU_CH_UINT64, 8) X(CIT_COUNT, "cit_count", DFGPU_CH_UINT64, 8) \
  X(SRT_MAX, "srt_max", DFGPU_CH_UINT32, 4) X(ART_MAX, "art_max", DFGPU_CH_UINT32, 4)         \
  X(RRT_MAX, "rrt_max", DFGPU_CH_UINT32, 4) X(CIT_MAX, "cit_max", DFGPU_CH_UINT32, 4)         \
  X(RETRANS_TX, "retrans_tx", DFGPU_CH_UINT32, 4) X(RETRANS_RX, "retrans_rx", DFGPU_CH_UINT32, 4) \
  X(ZERO_WIN_TX, "zero_win_tx", DFGPU_CH_UINT32, 4) X(ZERO_WIN_RX, "zero_win_rx", DFGPU_CH_UINT32, 4) \
  X(SYN_COUNT, "syn_count", DFGPU_CH_UINT32, 4) X(SYNACK_COUNT, "synack_count", DFGPU_CH_UINT32, 4) \
  X(RETRANS_SYN, "retrans_syn", DFGPU_CH_UINT32, 4) X(RETRANS_SYNACK, "retrans_synack", DFGPU_CH_UINT32, 4) \
  X(L7_CLIENT_ERROR, "l7_client_error", DFGPU_CH_UINT32, 4) X(L7_SERVER_ERROR, "l7_server_error", DFGPU_CH_UINT32, 4) \
  X(L7_SERVER_TIMEOUT, "l7_server_timeout", DFGPU_CH_UINT32, 4) X(L7_ERROR, "l7_error", DFGPU_CH_UINT32, 4) \
  X(OOO_RX, "ooo_rx", DFGPU_CH_UINT32, 4) X(OOO_TX, "ooo_tx", DFGPU_CH_UINT32, 4)             \
  X(FIN_COUNT, "fin_count", DFGPU_CH_UINT32, 4)

namespace dfgpu {
#define X(id, name, t, es) L4C_##id,
enum L4Col : int { L4_COLUMNS(X) L4C_COUNT };
#undef X
// scan slots of the L4 pipeline
enum { L4S_ROWS = 0, L4S_AGG_BYTES = 1, L4S_ACL_ELEMS = 2, L4S_DOM_BYTES = 3, L4S_COUNT = 4 };
// LowCardinality code spaces (first-appearance tracking)
enum { LC_PROVINCE_0 = 0, LC_PROVINCE_1 = 1, LC_TAPSIDE = 2 };
}  // namespace dfgpu
