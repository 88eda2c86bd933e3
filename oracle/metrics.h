/*
 * ORACLE — test infrastructure only.
 * metrics.h: CPU restatement of the flow_metrics unmarshaller path (SURVEY A10, A5
 * twin, A8 metrics blocks, A9):
 *   app.DecodePB                 server/libs/app/codec.go:28-65
 *   pb.Document prealloc/reset   server/libs/flow-metrics/pb/doc.go:19-106
 *   Tag.ReadFromPB               server/libs/flow-metrics/tag.go:1176-1240
 *   *Meter.ReadFromPB            flow-metrics/basic_meter.go, app_meter.go, usage_meter.go
 *   DocumentExpand               server/ingester/flow_metrics/unmarshaller/handle_document.go:41-306
 *   Tag.TableID                  tag.go:920-931 (+ app/document.go:81-83)
 *   QueueProcess control flow    unmarshaller/unmarshaller.go:215-275, appServiceTagWrite :309-327
 *   column blocks                tag_column_block.go, basic_meter_column_block.go,
 *                                app_meter_column_block.go, usage_meter_column_block.go,
 *                                app/document_column_block.go
 *   ConcurrentMerge              basic_meter.go:94-113,277-315,417-424,556-577,707-709; app_meter.go; usage_meter.go
 */
#pragma once
#include <unordered_map>
#include <string>

#include "../include/dfgpu_metrics_schema.h"
#include "l7.h"

namespace oracle {

struct MetricSchema {  // message/metric.proto:14-195
  MsgDesc mini_field{"MiniField",
                     {{1, K_STR, nullptr},  {2, K_STR, nullptr},  {3, K_U32, nullptr},  {4, K_U32, nullptr},
                      {5, K_I32, nullptr},  {6, K_I32, nullptr},  {7, K_U64, nullptr},  {8, K_U64, nullptr},
                      {9, K_U32, nullptr},  {10, K_U32, nullptr}, {11, K_U32, nullptr}, {12, K_U32, nullptr},
                      {13, K_U32, nullptr}, {14, K_U32, nullptr}, {15, K_U64, nullptr}, {16, K_U32, nullptr},
                      {17, K_U32, nullptr}, {20, K_U32, nullptr}, {21, K_U32, nullptr}, {22, K_U32, nullptr},
                      {23, K_STR, nullptr}, {24, K_STR, nullptr}, {25, K_STR, nullptr}, {27, K_U32, nullptr},
                      {28, K_U32, nullptr}}};
  MsgDesc mini_tag{"MiniTag", {{1, K_MSG, &mini_field}, {2, K_U64, nullptr}}};
  MsgDesc traffic{"Traffic",
                  {{1, K_U64, nullptr},  {2, K_U64, nullptr},  {3, K_U64, nullptr},  {4, K_U64, nullptr},
                   {5, K_U64, nullptr},  {6, K_U64, nullptr},  {7, K_U64, nullptr},  {8, K_U64, nullptr},
                   {9, K_U64, nullptr},  {10, K_U64, nullptr}, {11, K_U32, nullptr}, {12, K_U32, nullptr},
                   {13, K_U32, nullptr}, {14, K_U32, nullptr}, {15, K_U32, nullptr}}};
  MsgDesc latency{"Latency",
                  {{1, K_U32, nullptr},  {2, K_U32, nullptr},  {3, K_U32, nullptr},  {4, K_U32, nullptr},
                   {5, K_U32, nullptr},  {6, K_U32, nullptr},  {19, K_U32, nullptr}, {7, K_U64, nullptr},
                   {8, K_U64, nullptr},  {9, K_U64, nullptr},  {10, K_U64, nullptr}, {11, K_U64, nullptr},
                   {12, K_U64, nullptr}, {20, K_U64, nullptr}, {13, K_U32, nullptr}, {14, K_U32, nullptr},
                   {15, K_U32, nullptr}, {16, K_U32, nullptr}, {17, K_U32, nullptr}, {18, K_U32, nullptr},
                   {21, K_U32, nullptr}}};
  MsgDesc performance{"Performance",
                      {{1, K_U64, nullptr}, {2, K_U64, nullptr}, {3, K_U64, nullptr}, {4, K_U64, nullptr},
                       {5, K_U32, nullptr}, {6, K_U32, nullptr}}};
  MsgDesc anomaly{"Anomaly",
                  {{1, K_U64, nullptr},  {2, K_U64, nullptr},  {3, K_U64, nullptr},  {4, K_U64, nullptr},
                   {5, K_U64, nullptr},  {6, K_U64, nullptr},  {7, K_U64, nullptr},  {8, K_U64, nullptr},
                   {9, K_U64, nullptr},  {10, K_U64, nullptr}, {11, K_U64, nullptr}, {12, K_U64, nullptr},
                   {13, K_U32, nullptr}, {14, K_U32, nullptr}, {15, K_U32, nullptr}, {20, K_U64, nullptr},
                   {21, K_U64, nullptr}}};
  MsgDesc flow_load{"FlowLoad", {{1, K_U64, nullptr}}};
  MsgDesc flow_meter{"FlowMeter",
                     {{1, K_MSG, &traffic}, {2, K_MSG, &latency}, {3, K_MSG, &performance}, {4, K_MSG, &anomaly},
                      {5, K_MSG, &flow_load}}};
  MsgDesc usage_meter{"UsageMeter",
                      {{1, K_U64, nullptr}, {2, K_U64, nullptr}, {3, K_U64, nullptr}, {4, K_U64, nullptr},
                       {5, K_U64, nullptr}, {6, K_U64, nullptr}, {7, K_U64, nullptr}, {8, K_U64, nullptr}}};
  MsgDesc app_traffic{"AppTraffic", {{1, K_U32, nullptr}, {2, K_U32, nullptr}, {3, K_U32, nullptr}}};
  MsgDesc app_latency{"AppLatency", {{1, K_U32, nullptr}, {2, K_U64, nullptr}, {3, K_U32, nullptr}}};
  MsgDesc app_anomaly{"AppAnomaly", {{1, K_U32, nullptr}, {2, K_U32, nullptr}, {3, K_U32, nullptr}}};
  MsgDesc app_meter{"AppMeter", {{1, K_MSG, &app_traffic}, {2, K_MSG, &app_latency}, {3, K_MSG, &app_anomaly}}};
  MsgDesc meter{"Meter", {{1, K_U32, nullptr}, {2, K_MSG, &flow_meter}, {3, K_MSG, &usage_meter}, {4, K_MSG, &app_meter}}};
  MsgDesc document{"Document", {{1, K_U32, nullptr}, {2, K_MSG, &mini_tag}, {3, K_MSG, &meter}, {4, K_U32, nullptr}}};
};
const MetricSchema& metric_schema();

// flow_metrics.Tag = Field + Code (tag.go:269-354) with the member names the shared
// column table (dfgpu_metrics_schema.h) refers to.
struct MTag {
  uint64_t code = 0;
  uint32_t time = 0;  // Document.Timestamp
  uint8_t gtid = 0, is_ipv4 = 0;
  uint32_t ip4[2] = {0, 0};
  uint8_t ip6[2][16] = {{0}, {0}};
  uint64_t mac[2] = {0, 0};
  int32_t l3_epc[2] = {0, 0};
  uint32_t l3_device_id[2] = {0, 0};
  uint8_t l3_device_type[2] = {0, 0};
  uint16_t region_id[2] = {0, 0}, subnet_id[2] = {0, 0}, host_id[2] = {0, 0};
  uint32_t pod_node_id[2] = {0, 0};
  uint16_t az_id[2] = {0, 0};
  uint32_t pod_group_id[2] = {0, 0};
  uint16_t pod_ns_id[2] = {0, 0};
  uint32_t pod_id[2] = {0, 0};
  uint16_t pod_cluster_id[2] = {0, 0};
  uint32_t service_id[2] = {0, 0};
  uint32_t auto_instance_id[2] = {0, 0};
  uint8_t auto_instance_type[2] = {0, 0};
  uint32_t auto_service_id[2] = {0, 0};
  uint8_t auto_service_type[2] = {0, 0};
  uint32_t gpid[2] = {0, 0};
  uint16_t acl_gid = 0;
  uint8_t role = 0, protocol = 0;
  uint16_t server_port = 0, vtap_id = 0, org_id = 0, team_id = 0;
  uint32_t tap_port = 0;
  uint8_t tap_port_type = 0, nat_source = 0, tunnel_type = 0;
  uint8_t tap_side = 0, tap_type = 0, is_key_service = 0, l7_protocol = 0, biz_type = 0;
  uint16_t signal_source = 0;
  uint8_t tag_source[2] = {0, 0};
  uint16_t tunnel_ip_id = 0;
  std::string app_service, app_instance, endpoint;
};

struct MFlowMeter {  // FlowMeter = Traffic + Latency + Performance + Anomaly + FlowLoad
  uint64_t packet_tx = 0, packet_rx = 0, byte_tx = 0, byte_rx = 0, l3_byte_tx = 0, l3_byte_rx = 0, l4_byte_tx = 0,
           l4_byte_rx = 0, new_flow = 0, closed_flow = 0;
  uint32_t l7_request = 0, l7_response = 0, syn_count = 0, synack_count = 0;
  uint8_t direction_score = 0;
  uint32_t rtt_max = 0, rtt_client_max = 0, rtt_server_max = 0, srt_max = 0, art_max = 0, rrt_max = 0, cit_max = 0;
  uint64_t rtt_sum = 0, rtt_client_sum = 0, rtt_server_sum = 0, srt_sum = 0, art_sum = 0, rrt_sum = 0, cit_sum = 0;
  uint32_t rtt_count = 0, rtt_client_count = 0, rtt_server_count = 0, srt_count = 0, art_count = 0, rrt_count = 0,
           cit_count = 0;
  uint64_t retrans_tx = 0, retrans_rx = 0, zero_win_tx = 0, zero_win_rx = 0;
  uint32_t retrans_syn = 0, retrans_synack = 0;
  uint64_t client_rst_flow = 0, server_rst_flow = 0, server_syn_miss = 0, client_ack_miss = 0,
           client_half_close_flow = 0, server_half_close_flow = 0, client_source_port_reuse = 0,
           client_establish_reset = 0, server_reset = 0, server_queue_lack = 0, server_establish_reset = 0,
           tcp_timeout = 0, client_establish_fail = 0, server_establish_fail = 0, tcp_establish_fail = 0,
           tcp_transfer_fail = 0, tcp_rst_fail = 0;
  uint32_t l7_client_error = 0, l7_server_error = 0, l7_timeout = 0;
  uint64_t ooo_tx = 0, ooo_rx = 0;
  uint64_t flow_load = 0;
  void ReadFromPB(const Msg& flow);           // flow_meter.go:83-89
  void ConcurrentMerge(const MFlowMeter& o);  // flow_meter.go:91-99
};
struct MAppMeter {
  uint32_t request = 0, response = 0;
  uint8_t direction_score = 0;
  uint32_t rrt_max = 0;
  uint64_t rrt_sum = 0;
  uint32_t rrt_count = 0;
  uint32_t client_error = 0, server_error = 0, timeout = 0;
  void ReadFromPB(const Msg& app);           // app_meter.go:72-76
  void ConcurrentMerge(const MAppMeter& o);  // app_meter.go:78-84
};
struct MUsageMeter {
  uint64_t packet_tx = 0, packet_rx = 0, byte_tx = 0, byte_rx = 0, l3_byte_tx = 0, l3_byte_rx = 0, l4_byte_tx = 0,
           l4_byte_rx = 0;
  void ReadFromPB(const Msg& usage);           // usage_meter.go
  void ConcurrentMerge(const MUsageMeter& o);  // usage_meter.go:150-154 -> Merge
};

struct MDoc {
  MTag t;
  uint32_t flags = 0;
  uint8_t meter_id = 0;
  MFlowMeter flow;
  MAppMeter app;
  MUsageMeter usage;
};

void TagReadFromPB(MTag& t, const Msg& miniTag);                 // tag.go:1176-1240
bool DocumentExpand(MDoc& d, const PlatformInfoTable* pd);       // handle_document.go:150-306; false = dropped
int TagTableID(uint64_t code, bool isSecond);                    // tag.go:920-931; -1 = no table
uint64_t metrics_table_code(int table);                          // metricsTableCodes tag.go:508-523
int metrics_table_meter(int table);                              // which meter the table's documents carry

struct MetricsBlocks {
  Block tables[MT_TABLE_COUNT];
  Block app_service;  // AppServiceTagWriter rows (unmarshaller.go:309-327)
  Block reduced;
  std::vector<MDoc> rows[MT_TABLE_COUNT];  // kept for reduce()
  void init();
  void clear();
  void append_record(const Msg& doc, uint16_t orgId, uint16_t teamId, const PlatformInfoTable* pd,
                     dfgpu_counters& cnt, L7State& st);
  int reduce(int table, uint32_t bucket_seconds);
  // rollup accumulator across submits (dfgpu_rollup_add / dfgpu_rollup_flush)
  std::unordered_map<std::string, size_t> acc_idx[MT_TABLE_COUNT];
  std::vector<MDoc> acc_out[MT_TABLE_COUNT];
  int rollup_add(int table, uint32_t bucket_seconds);
  int rollup_flush(int table);
  static void add_columns(Block& b, int table);
  static void append_row(Block& b, int table, const MDoc& d);
};

}  // namespace oracle
