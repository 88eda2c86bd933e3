/*
 * dfgpu_metrics_schema.h — DATA ONLY: the flow_metrics tag Code bits, the 9 table
 * codes and the output column lists (name, container, order, source field) of the
 * metrics tables, as X-macros.  Shared by the CUDA library and by the test oracle the
 * same way the wire schema is; the names / order / container types are checked against
 * lists machine-extracted from the reference's ToInput functions
 * (tests/golden/metrics_columns.json, tools/make_metrics_golden.py).
 *
 * Reference:
 *   Code bits            server/libs/flow-metrics/tag.go:36-98
 *   table codes          tag.go:495-506, table ids :445-461
 *   TagBlock.ToInput     server/libs/flow-metrics/tag_column_block.go:290-510
 *   Tag.AppendTo...      tag_column_block.go:521-721
 *   FlowMeter blocks     server/libs/flow-metrics/basic_meter_column_block.go:53-422
 *   AppMeter blocks      app_meter_column_block.go:24-120
 *   UsageMeter block     usage_meter_column_block.go:24-82
 *
 * Expressions name members of a tag row `t` and a meter `m`; each side declares its own
 * structs with these member names (oracle/metrics.h, csrc/metrics_dev.h).
 */
#ifndef DFGPU_METRICS_SCHEMA_H
#define DFGPU_METRICS_SCHEMA_H

#include <stdint.h>

/* ---- Code bits (tag.go:38-94) ------------------------------------------------------- */
#define MC_IP               (1ull << 0)
#define MC_L3EPCID          (1ull << 1)
#define MC_L3DEVICE         (1ull << 2)
#define MC_SUBNETID         (1ull << 3)
#define MC_REGIONID         (1ull << 4)
#define MC_PODNODEID        (1ull << 5)
#define MC_HOSTID           (1ull << 6)
#define MC_AZID             (1ull << 7)
#define MC_PODGROUPID       (1ull << 8)
#define MC_PODNSID          (1ull << 9)
#define MC_PODID            (1ull << 10)
#define MC_MAC              (1ull << 11)
#define MC_PODCLUSTERID     (1ull << 12)
#define MC_SERVICEID        (1ull << 13)
#define MC_RESOURCE         (1ull << 14)
#define MC_GPID             (1ull << 15)
#define MC_PATH_SHIFT       20 /* X##Path = X << 20 (tag.go:59-79) */
#define MC_IPPATH           (MC_IP << MC_PATH_SHIFT)
#define MC_L3EPCIDPATH      (MC_L3EPCID << MC_PATH_SHIFT)
#define MC_L3DEVICEPATH     (MC_L3DEVICE << MC_PATH_SHIFT)
#define MC_SUBNETIDPATH     (MC_SUBNETID << MC_PATH_SHIFT)
#define MC_REGIONIDPATH     (MC_REGIONID << MC_PATH_SHIFT)
#define MC_PODNODEIDPATH    (MC_PODNODEID << MC_PATH_SHIFT)
#define MC_HOSTIDPATH       (MC_HOSTID << MC_PATH_SHIFT)
#define MC_AZIDPATH         (MC_AZID << MC_PATH_SHIFT)
#define MC_PODGROUPIDPATH   (MC_PODGROUPID << MC_PATH_SHIFT)
#define MC_PODNSIDPATH      (MC_PODNSID << MC_PATH_SHIFT)
#define MC_PODIDPATH        (MC_PODID << MC_PATH_SHIFT)
#define MC_MACPATH          (MC_MAC << MC_PATH_SHIFT)
#define MC_PODCLUSTERIDPATH (MC_PODCLUSTERID << MC_PATH_SHIFT)
#define MC_SERVICEIDPATH    (MC_SERVICEID << MC_PATH_SHIFT)
#define MC_RESOURCEPATH     (MC_RESOURCE << MC_PATH_SHIFT)
#define MC_GPIDPATH         (MC_GPID << MC_PATH_SHIFT)
#define MC_DIRECTION        (1ull << 40)
#define MC_ACLGID           (1ull << 41)
#define MC_PROTOCOL         (1ull << 42)
#define MC_SERVERPORT       (1ull << 43)
#define MC_TAPTYPE          (1ull << 45)
#define MC_VTAPID           (1ull << 47)
#define MC_TAPSIDE          (1ull << 48)
#define MC_TAPPORT          (1ull << 49)
#define MC_ISKEYSERVICE     (1ull << 50)
#define MC_L7PROTOCOL       (1ull << 51)
#define MC_SIGNALSOURCE     (1ull << 52)
#define MC_TUNNELIPID       (1ull << 62)
#define MC_EDGE_FIELDS      0xfffff00000ull /* Code.HasEdgeTagField tag.go:100-102 */

/* tag.go:495-506 */
#define MC_BASECODE                                                                                          \
  (MC_AZID | MC_HOSTID | MC_IP | MC_L3DEVICE | MC_L3EPCID | MC_PODCLUSTERID | MC_PODGROUPID | MC_PODID |     \
   MC_PODNODEID | MC_PODNSID | MC_REGIONID | MC_SUBNETID | MC_TAPTYPE | MC_VTAPID | MC_SERVICEID |           \
   MC_RESOURCE | MC_GPID | MC_SIGNALSOURCE)
#define MC_BASEPATHCODE                                                                                      \
  (MC_AZIDPATH | MC_HOSTIDPATH | MC_IPPATH | MC_L3DEVICEPATH | MC_L3EPCIDPATH | MC_PODCLUSTERIDPATH |        \
   MC_PODGROUPIDPATH | MC_PODIDPATH | MC_PODNODEIDPATH | MC_PODNSIDPATH | MC_REGIONIDPATH |                  \
   MC_SUBNETIDPATH | MC_TAPSIDE | MC_TAPTYPE | MC_VTAPID | MC_SERVICEIDPATH | MC_RESOURCEPATH |              \
   MC_GPIDPATH | MC_SIGNALSOURCE)
#define MC_BASEPORTCODE    (MC_PROTOCOL | MC_SERVERPORT | MC_ISKEYSERVICE)
#define MC_NETWORK         (MC_BASECODE | MC_BASEPORTCODE | MC_DIRECTION)
#define MC_NETWORK_MAP     (MC_BASEPATHCODE | MC_BASEPORTCODE | MC_TAPPORT)
#define MC_APPLICATION     (MC_BASECODE | MC_BASEPORTCODE | MC_DIRECTION | MC_L7PROTOCOL)
#define MC_APPLICATION_MAP (MC_BASEPATHCODE | MC_BASEPORTCODE | MC_TAPPORT | MC_L7PROTOCOL)
#define MC_TRAFFIC_POLICY  (MC_ACLGID | MC_TUNNELIPID | MC_VTAPID)

/* unmarshaller/handle_document.go:32-38 */
#define MC_EDGECODE (MC_IPPATH | MC_L3EPCIDPATH)
#define MC_MAINADDCODE                                                                                        \
  (MC_REGIONID | MC_HOSTID | MC_L3DEVICE | MC_SUBNETID | MC_PODNODEID | MC_AZID | MC_PODGROUPID | MC_PODNSID | \
   MC_PODID | MC_PODCLUSTERID | MC_SERVICEID | MC_RESOURCE)
#define MC_EDGEADDCODE (MC_MAINADDCODE << MC_PATH_SHIFT)
#define MC_PORTADDCODE MC_ISKEYSERVICE

/* MetricsTableID (tag.go:445-461); DFGPU_TABLE_METRICS_BASE + id */
enum {
  MT_NETWORK_1M = 0, MT_NETWORK_MAP_1M, MT_APPLICATION_1M, MT_APPLICATION_MAP_1M, MT_TRAFFIC_POLICY_1M,
  MT_NETWORK_1S, MT_NETWORK_MAP_1S, MT_APPLICATION_1S, MT_APPLICATION_MAP_1S, MT_TABLE_COUNT
};
/* meter ids (flow-metrics/const.go:27-36) and the meter each table stores */
enum { MT_METER_FLOW = 1, MT_METER_USAGE = 4, MT_METER_APP = 5 };

/*
 * Tag columns, TagBlock.ToInput order.
 *   F(bit, name, CH, ctype, expr)   fixed-width value
 *   V6(bit, name, side)             ColIPv6 from t.ip6[side] (nil when IPv4 -> zeros, ckdb.go:46-56)
 *   OBS(bit, name)                  observation_point, LowCardinality of TAPSide.String()
 *   LCS(bit, name, which)           LowCardinality string column fed per row (0 app_service, 1 app_instance)
 *   STR(bit, name, which)           ColStr (2 endpoint)
 * bit = 0: always present.
 */
#define MT_TAG_COLUMNS(F, V6, OBS, LCS, STR)                                                  \
  F(0, "time", DFGPU_CH_DATETIME, uint32_t, t.time)                                           \
  F(0, "_tid", DFGPU_CH_UINT8, uint8_t, t.gtid)                                               \
  F(MC_ACLGID, "acl_gid", DFGPU_CH_UINT16, uint16_t, t.acl_gid)                               \
  F(MC_AZID, "az_id", DFGPU_CH_UINT16, uint16_t, t.az_id[0])                                  \
  F(MC_AZIDPATH, "az_id_0", DFGPU_CH_UINT16, uint16_t, t.az_id[0])                            \
  F(MC_AZIDPATH, "az_id_1", DFGPU_CH_UINT16, uint16_t, t.az_id[1])                            \
  F(MC_DIRECTION, "role", DFGPU_CH_UINT8, uint8_t, t.role)                                    \
  F(MC_GPID, "gprocess_id", DFGPU_CH_UINT32, uint32_t, t.gpid[0])                             \
  F(MC_GPIDPATH, "gprocess_id_0", DFGPU_CH_UINT32, uint32_t, t.gpid[0])                       \
  F(MC_GPIDPATH, "gprocess_id_1", DFGPU_CH_UINT32, uint32_t, t.gpid[1])                       \
  F(MC_HOSTID, "host_id", DFGPU_CH_UINT16, uint16_t, t.host_id[0])                            \
  F(MC_HOSTIDPATH, "host_id_0", DFGPU_CH_UINT16, uint16_t, t.host_id[0])                      \
  F(MC_HOSTIDPATH, "host_id_1", DFGPU_CH_UINT16, uint16_t, t.host_id[1])                      \
  F(MC_IP, "ip4", DFGPU_CH_IPV4, uint32_t, t.ip4[0])                                          \
  V6(MC_IP, "ip6", 0)                                                                         \
  F(MC_IP, "is_ipv4", DFGPU_CH_UINT8, uint8_t, t.is_ipv4)                                     \
  F(MC_IP, "tag_source", DFGPU_CH_UINT8, uint8_t, t.tag_source[0])                            \
  F(MC_IPPATH, "ip4_0", DFGPU_CH_IPV4, uint32_t, t.ip4[0])                                    \
  F(MC_IPPATH, "ip4_1", DFGPU_CH_IPV4, uint32_t, t.ip4[1])                                    \
  V6(MC_IPPATH, "ip6_0", 0)                                                                   \
  V6(MC_IPPATH, "ip6_1", 1)                                                                   \
  F(MC_IPPATH, "is_ipv4", DFGPU_CH_UINT8, uint8_t, t.is_ipv4)                                 \
  F(MC_IPPATH, "tag_source_0", DFGPU_CH_UINT8, uint8_t, t.tag_source[0])                      \
  F(MC_IPPATH, "tag_source_1", DFGPU_CH_UINT8, uint8_t, t.tag_source[1])                      \
  F(MC_ISKEYSERVICE, "is_key_service", DFGPU_CH_UINT8, uint8_t, t.is_key_service)             \
  F(MC_L3DEVICE, "l3_device_id", DFGPU_CH_UINT32, uint32_t, t.l3_device_id[0])                \
  F(MC_L3DEVICE, "l3_device_type", DFGPU_CH_UINT8, uint8_t, t.l3_device_type[0])              \
  F(MC_L3DEVICEPATH, "l3_device_id_0", DFGPU_CH_UINT32, uint32_t, t.l3_device_id[0])          \
  F(MC_L3DEVICEPATH, "l3_device_id_1", DFGPU_CH_UINT32, uint32_t, t.l3_device_id[1])          \
  F(MC_L3DEVICEPATH, "l3_device_type_0", DFGPU_CH_UINT8, uint8_t, t.l3_device_type[0])        \
  F(MC_L3DEVICEPATH, "l3_device_type_1", DFGPU_CH_UINT8, uint8_t, t.l3_device_type[1])        \
  F(MC_L3EPCID, "l3_epc_id", DFGPU_CH_INT32, int32_t, t.l3_epc[0])                            \
  F(MC_L3EPCIDPATH, "l3_epc_id_0", DFGPU_CH_INT32, int32_t, t.l3_epc[0])                      \
  F(MC_L3EPCIDPATH, "l3_epc_id_1", DFGPU_CH_INT32, int32_t, t.l3_epc[1])                      \
  F(MC_L7PROTOCOL, "l7_protocol", DFGPU_CH_UINT8, uint8_t, t.l7_protocol)                     \
  LCS(MC_L7PROTOCOL, "app_service", 0)                                                        \
  LCS(MC_L7PROTOCOL, "app_instance", 1)                                                       \
  STR(MC_L7PROTOCOL, "endpoint", 2)                                                           \
  F(MC_L7PROTOCOL, "biz_type", DFGPU_CH_UINT8, uint8_t, t.biz_type)                           \
  F(MC_PODCLUSTERID, "pod_cluster_id", DFGPU_CH_UINT16, uint16_t, t.pod_cluster_id[0])        \
  F(MC_PODCLUSTERIDPATH, "pod_cluster_id_0", DFGPU_CH_UINT16, uint16_t, t.pod_cluster_id[0])  \
  F(MC_PODCLUSTERIDPATH, "pod_cluster_id_1", DFGPU_CH_UINT16, uint16_t, t.pod_cluster_id[1])  \
  F(MC_PODGROUPID, "pod_group_id", DFGPU_CH_UINT32, uint32_t, t.pod_group_id[0])              \
  F(MC_PODGROUPIDPATH, "pod_group_id_0", DFGPU_CH_UINT32, uint32_t, t.pod_group_id[0])        \
  F(MC_PODGROUPIDPATH, "pod_group_id_1", DFGPU_CH_UINT32, uint32_t, t.pod_group_id[1])        \
  F(MC_PODID, "pod_id", DFGPU_CH_UINT32, uint32_t, t.pod_id[0])                               \
  F(MC_PODIDPATH, "pod_id_0", DFGPU_CH_UINT32, uint32_t, t.pod_id[0])                         \
  F(MC_PODIDPATH, "pod_id_1", DFGPU_CH_UINT32, uint32_t, t.pod_id[1])                         \
  F(MC_PODNODEID, "pod_node_id", DFGPU_CH_UINT32, uint32_t, t.pod_node_id[0])                 \
  F(MC_PODNODEIDPATH, "pod_node_id_0", DFGPU_CH_UINT32, uint32_t, t.pod_node_id[0])           \
  F(MC_PODNODEIDPATH, "pod_node_id_1", DFGPU_CH_UINT32, uint32_t, t.pod_node_id[1])           \
  F(MC_PODNSID, "pod_ns_id", DFGPU_CH_UINT16, uint16_t, t.pod_ns_id[0])                       \
  F(MC_PODNSIDPATH, "pod_ns_id_0", DFGPU_CH_UINT16, uint16_t, t.pod_ns_id[0])                 \
  F(MC_PODNSIDPATH, "pod_ns_id_1", DFGPU_CH_UINT16, uint16_t, t.pod_ns_id[1])                 \
  F(MC_PROTOCOL, "protocol", DFGPU_CH_UINT8, uint8_t, t.protocol)                             \
  F(MC_REGIONID, "region_id", DFGPU_CH_UINT16, uint16_t, t.region_id[0])                      \
  F(MC_REGIONIDPATH, "region_id_0", DFGPU_CH_UINT16, uint16_t, t.region_id[0])                \
  F(MC_REGIONIDPATH, "region_id_1", DFGPU_CH_UINT16, uint16_t, t.region_id[1])                \
  F(MC_RESOURCE, "auto_instance_id", DFGPU_CH_UINT32, uint32_t, t.auto_instance_id[0])        \
  F(MC_RESOURCE, "auto_instance_type", DFGPU_CH_UINT8, uint8_t, t.auto_instance_type[0])      \
  F(MC_RESOURCE, "auto_service_id", DFGPU_CH_UINT32, uint32_t, t.auto_service_id[0])          \
  F(MC_RESOURCE, "auto_service_type", DFGPU_CH_UINT8, uint8_t, t.auto_service_type[0])        \
  F(MC_RESOURCEPATH, "auto_instance_id_0", DFGPU_CH_UINT32, uint32_t, t.auto_instance_id[0])  \
  F(MC_RESOURCEPATH, "auto_instance_type_0", DFGPU_CH_UINT8, uint8_t, t.auto_instance_type[0]) \
  F(MC_RESOURCEPATH, "auto_service_id_0", DFGPU_CH_UINT32, uint32_t, t.auto_service_id[0])    \
  F(MC_RESOURCEPATH, "auto_service_type_0", DFGPU_CH_UINT8, uint8_t, t.auto_service_type[0])  \
  F(MC_RESOURCEPATH, "auto_instance_id_1", DFGPU_CH_UINT32, uint32_t, t.auto_instance_id[1])  \
  F(MC_RESOURCEPATH, "auto_instance_type_1", DFGPU_CH_UINT8, uint8_t, t.auto_instance_type[1]) \
  F(MC_RESOURCEPATH, "auto_service_id_1", DFGPU_CH_UINT32, uint32_t, t.auto_service_id[1])    \
  F(MC_RESOURCEPATH, "auto_service_type_1", DFGPU_CH_UINT8, uint8_t, t.auto_service_type[1])  \
  F(MC_SIGNALSOURCE, "signal_source", DFGPU_CH_UINT16, uint16_t, t.signal_source)             \
  F(MC_SERVICEID, "service_id", DFGPU_CH_UINT32, uint32_t, t.service_id[0])                   \
  F(MC_SERVICEIDPATH, "service_id_0", DFGPU_CH_UINT32, uint32_t, t.service_id[0])             \
  F(MC_SERVICEIDPATH, "service_id_1", DFGPU_CH_UINT32, uint32_t, t.service_id[1])             \
  F(MC_SERVERPORT, "server_port", DFGPU_CH_UINT16, uint16_t, t.server_port)                   \
  F(MC_SUBNETID, "subnet_id", DFGPU_CH_UINT16, uint16_t, t.subnet_id[0])                      \
  F(MC_SUBNETIDPATH, "subnet_id_0", DFGPU_CH_UINT16, uint16_t, t.subnet_id[0])                \
  F(MC_SUBNETIDPATH, "subnet_id_1", DFGPU_CH_UINT16, uint16_t, t.subnet_id[1])                \
  F(MC_TUNNELIPID, "tunnel_ip_id", DFGPU_CH_UINT16, uint16_t, t.tunnel_ip_id)                 \
  F(MC_TAPPORT, "capture_nic_type", DFGPU_CH_UINT8, uint8_t, t.tap_port_type)                 \
  F(MC_TAPPORT, "tunnel_type", DFGPU_CH_UINT8, uint8_t, t.tunnel_type)                        \
  F(MC_TAPPORT, "capture_nic", DFGPU_CH_UINT32, uint32_t, t.tap_port)                         \
  F(MC_TAPPORT, "nat_source", DFGPU_CH_UINT8, uint8_t, t.nat_source)                          \
  OBS(MC_TAPSIDE, "observation_point")                                                        \
  F(MC_TAPTYPE, "capture_network_type_id", DFGPU_CH_UINT8, uint8_t, t.tap_type)               \
  F(MC_VTAPID, "agent_id", DFGPU_CH_UINT16, uint16_t, t.vtap_id)                              \
  F(MC_VTAPID, "team_id", DFGPU_CH_UINT16, uint16_t, t.team_id)

/* FlowMeter columns: Traffic, Latency, Performance, Anomaly, FlowLoad (ToInput order).
 * M(name, CH, ctype, expr) */
#define MT_FLOW_METER_COLUMNS(M)                                                               \
  M("packet_tx", DFGPU_CH_UINT64, uint64_t, m.packet_tx)                                       \
  M("packet_rx", DFGPU_CH_UINT64, uint64_t, m.packet_rx)                                       \
  M("packet", DFGPU_CH_UINT64, uint64_t, m.packet_tx + m.packet_rx)                            \
  M("byte_tx", DFGPU_CH_UINT64, uint64_t, m.byte_tx)                                           \
  M("byte_rx", DFGPU_CH_UINT64, uint64_t, m.byte_rx)                                           \
  M("byte", DFGPU_CH_UINT64, uint64_t, m.byte_tx + m.byte_rx)                                  \
  M("l3_byte_tx", DFGPU_CH_UINT64, uint64_t, m.l3_byte_tx)                                     \
  M("l3_byte_rx", DFGPU_CH_UINT64, uint64_t, m.l3_byte_rx)                                     \
  M("l4_byte_tx", DFGPU_CH_UINT64, uint64_t, m.l4_byte_tx)                                     \
  M("l4_byte_rx", DFGPU_CH_UINT64, uint64_t, m.l4_byte_rx)                                     \
  M("new_flow", DFGPU_CH_UINT64, uint64_t, m.new_flow)                                         \
  M("closed_flow", DFGPU_CH_UINT64, uint64_t, m.closed_flow)                                   \
  M("l7_request", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.l7_request)                           \
  M("l7_response", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.l7_response)                         \
  M("syn_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.syn_count)                             \
  M("synack_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.synack_count)                       \
  M("direction_score", DFGPU_CH_UINT8, uint8_t, m.direction_score)                             \
  M("rtt_sum", DFGPU_CH_FLOAT64, double, (double)m.rtt_sum)                                    \
  M("rtt_client_sum", DFGPU_CH_FLOAT64, double, (double)m.rtt_client_sum)                      \
  M("rtt_server_sum", DFGPU_CH_FLOAT64, double, (double)m.rtt_server_sum)                      \
  M("srt_sum", DFGPU_CH_FLOAT64, double, (double)m.srt_sum)                                    \
  M("art_sum", DFGPU_CH_FLOAT64, double, (double)m.art_sum)                                    \
  M("rrt_sum", DFGPU_CH_FLOAT64, double, (double)m.rrt_sum)                                    \
  M("cit_sum", DFGPU_CH_FLOAT64, double, (double)m.cit_sum)                                    \
  M("rtt_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.rtt_count)                             \
  M("rtt_client_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.rtt_client_count)               \
  M("rtt_server_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.rtt_server_count)               \
  M("srt_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.srt_count)                             \
  M("art_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.art_count)                             \
  M("rrt_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.rrt_count)                             \
  M("cit_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.cit_count)                             \
  M("rtt_max", DFGPU_CH_UINT32, uint32_t, m.rtt_max)                                           \
  M("rtt_client_max", DFGPU_CH_UINT32, uint32_t, m.rtt_client_max)                             \
  M("rtt_server_max", DFGPU_CH_UINT32, uint32_t, m.rtt_server_max)                             \
  M("srt_max", DFGPU_CH_UINT32, uint32_t, m.srt_max)                                           \
  M("art_max", DFGPU_CH_UINT32, uint32_t, m.art_max)                                           \
  M("rrt_max", DFGPU_CH_UINT32, uint32_t, m.rrt_max)                                           \
  M("cit_max", DFGPU_CH_UINT32, uint32_t, m.cit_max)                                           \
  M("retrans_tx", DFGPU_CH_UINT64, uint64_t, m.retrans_tx)                                     \
  M("retrans_rx", DFGPU_CH_UINT64, uint64_t, m.retrans_rx)                                     \
  M("retrans", DFGPU_CH_UINT64, uint64_t, m.retrans_tx + m.retrans_rx)                         \
  M("zero_win_tx", DFGPU_CH_UINT64, uint64_t, m.zero_win_tx)                                   \
  M("zero_win_rx", DFGPU_CH_UINT64, uint64_t, m.zero_win_rx)                                   \
  M("zero_win", DFGPU_CH_UINT64, uint64_t, m.zero_win_tx + m.zero_win_rx)                      \
  M("retrans_syn", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.retrans_syn)                         \
  M("retrans_synack", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.retrans_synack)                   \
  M("client_rst_flow", DFGPU_CH_UINT64, uint64_t, m.client_rst_flow)                           \
  M("server_rst_flow", DFGPU_CH_UINT64, uint64_t, m.server_rst_flow)                           \
  M("server_syn_miss", DFGPU_CH_UINT64, uint64_t, m.server_syn_miss)                           \
  M("client_ack_miss", DFGPU_CH_UINT64, uint64_t, m.client_ack_miss)                           \
  M("client_half_close_flow", DFGPU_CH_UINT64, uint64_t, m.client_half_close_flow)             \
  M("server_half_close_flow", DFGPU_CH_UINT64, uint64_t, m.server_half_close_flow)             \
  M("client_source_port_reuse", DFGPU_CH_UINT64, uint64_t, m.client_source_port_reuse)         \
  M("server_reset", DFGPU_CH_UINT64, uint64_t, m.server_reset)                                 \
  M("server_queue_lack", DFGPU_CH_UINT64, uint64_t, m.server_queue_lack)                       \
  M("client_establish_other_rst", DFGPU_CH_UINT64, uint64_t, m.client_establish_reset)         \
  M("server_establish_other_rst", DFGPU_CH_UINT64, uint64_t, m.server_establish_reset)         \
  M("tcp_timeout", DFGPU_CH_UINT64, uint64_t, m.tcp_timeout)                                   \
  M("client_establish_fail", DFGPU_CH_UINT64, uint64_t, m.client_establish_fail)               \
  M("server_establish_fail", DFGPU_CH_UINT64, uint64_t, m.server_establish_fail)               \
  M("tcp_establish_fail", DFGPU_CH_UINT64, uint64_t, m.tcp_establish_fail)                     \
  M("tcp_transfer_fail", DFGPU_CH_UINT64, uint64_t, m.tcp_transfer_fail)                       \
  M("tcp_rst_fail", DFGPU_CH_UINT64, uint64_t, m.tcp_rst_fail)                                 \
  M("l7_client_error", DFGPU_CH_UINT32, uint32_t, m.l7_client_error)                           \
  M("l7_server_error", DFGPU_CH_UINT32, uint32_t, m.l7_server_error)                           \
  M("l7_timeout", DFGPU_CH_UINT32, uint32_t, m.l7_timeout)                                     \
  M("l7_error", DFGPU_CH_UINT32, uint32_t, (uint32_t)(m.l7_client_error + m.l7_server_error))  \
  M("ooo_tx", DFGPU_CH_UINT64, uint64_t, m.ooo_tx)                                             \
  M("ooo_rx", DFGPU_CH_UINT64, uint64_t, m.ooo_rx)                                             \
  M("flow_load", DFGPU_CH_UINT64, uint64_t, m.flow_load)

/* AppMeter columns: AppTraffic, AppLatency, AppAnomaly */
#define MT_APP_METER_COLUMNS(M)                                                                \
  M("request", DFGPU_CH_UINT32, uint32_t, m.request)                                           \
  M("response", DFGPU_CH_UINT32, uint32_t, m.response)                                         \
  M("direction_score", DFGPU_CH_UINT8, uint8_t, m.direction_score)                             \
  M("rrt_max", DFGPU_CH_UINT32, uint32_t, m.rrt_max)                                           \
  M("rrt_sum", DFGPU_CH_FLOAT64, double, (double)m.rrt_sum)                                    \
  M("rrt_count", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.rrt_count)                             \
  M("client_error", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.client_error)                       \
  M("server_error", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.server_error)                       \
  M("timeout", DFGPU_CH_UINT64, uint64_t, (uint64_t)m.timeout)                                 \
  M("error", DFGPU_CH_UINT64, uint64_t, (uint64_t)(uint32_t)(m.client_error + m.server_error))

/* UsageMeter columns */
#define MT_USAGE_METER_COLUMNS(M)                                                              \
  M("packet_tx", DFGPU_CH_UINT64, uint64_t, m.packet_tx)                                       \
  M("packet_rx", DFGPU_CH_UINT64, uint64_t, m.packet_rx)                                       \
  M("packet", DFGPU_CH_UINT64, uint64_t, m.packet_tx + m.packet_rx)                            \
  M("byte_tx", DFGPU_CH_UINT64, uint64_t, m.byte_tx)                                           \
  M("byte_rx", DFGPU_CH_UINT64, uint64_t, m.byte_rx)                                           \
  M("byte", DFGPU_CH_UINT64, uint64_t, m.byte_tx + m.byte_rx)                                  \
  M("l3_byte_tx", DFGPU_CH_UINT64, uint64_t, m.l3_byte_tx)                                     \
  M("l3_byte_rx", DFGPU_CH_UINT64, uint64_t, m.l3_byte_rx)                                     \
  M("l4_byte_tx", DFGPU_CH_UINT64, uint64_t, m.l4_byte_tx)                                     \
  M("l4_byte_rx", DFGPU_CH_UINT64, uint64_t, m.l4_byte_rx)

#endif /* DFGPU_METRICS_SCHEMA_H */
