/*
 * metrics_dev.h — device-side declarations of the flow_metrics path (metrics.cu):
 * interpreter slots of metric.Document (message/metric.proto:14-195), the decoded tag /
 * meter rows with the member names the shared column table refers to
 * (include/dfgpu_metrics_schema.h), and the per-batch output layout.
 */
#pragma once
#include "ctx.h"
#include "pb_interp.cuh"
#include "pb_tok.cuh"
#include "tagset.cuh"

#include "../../include/dfgpu_metrics_schema.h"

namespace dfgpu {

enum MMsg : uint32_t {
  MM_DOC = 0, MM_TAG, MM_FIELD, MM_METER, MM_FLOW, MM_TRAFFIC, MM_LATENCY, MM_PERF, MM_ANOMALY, MM_FLOWLOAD,
  MM_USAGE, MM_APP, MM_APPTRAFFIC, MM_APPLATENCY, MM_APPANOMALY, MM_COUNT
};
constexpr uint32_t MM_MAXF = 32;  // field numbers < 32 (MiniField goes up to 28)

enum MSlot : int {
  // Document
  MS_TS, MS_FLAGS, MS_TAG_CNT, MS_TAG_RANGE, MS_METER_CNT, MS_METER_RANGE,
  // MiniTag
  MS_FIELD_CNT, MS_FIELD_RANGE, MS_CODE,
  // MiniField
  MS_IP, MS_IP1, MS_GTID, MS_IS_IPV6, MS_EPC, MS_EPC1, MS_MAC, MS_MAC1, MS_DIRECTION, MS_TAP_SIDE, MS_PROTOCOL,
  MS_ACL_GID, MS_SERVER_PORT, MS_VTAP_ID, MS_TAP_PORT, MS_TAP_TYPE, MS_L7_PROTOCOL, MS_GPID, MS_GPID1,
  MS_SIGNAL_SOURCE, MS_APP_SERVICE, MS_APP_INSTANCE, MS_ENDPOINT, MS_POD_ID, MS_BIZ_TYPE,
  // Meter
  MS_METER_ID, MS_FLOW_CNT, MS_FLOW_RANGE, MS_USAGE_CNT, MS_USAGE_RANGE, MS_APP_CNT, MS_APP_RANGE,
  // FlowMeter sub-message occurrence slots (count, range) x 5
  MS_FLOW_SUB,
  // Traffic fields 1..15 -> MS_TRAFFIC + (f-1)
  MS_TRAFFIC = MS_FLOW_SUB + 10,
  // Latency: maxes 1..6,19 ; sums 7..12,20 ; counts 13..18,21 -> MS_LATENCY + index below
  MS_LATENCY = MS_TRAFFIC + 15,
  MS_PERF = MS_LATENCY + 21,      // fields 1..6
  MS_ANOMALY = MS_PERF + 6,       // fields 1..15 -> +0..14, 20 -> +15, 21 -> +16
  MS_FLOWLOAD = MS_ANOMALY + 17,  // field 1
  MS_USAGE = MS_FLOWLOAD + 1,     // fields 1..8
  MS_APP_SUB = MS_USAGE + 8,      // (count, range) x 3
  MS_APPTRAFFIC = MS_APP_SUB + 6, // fields 1..3
  MS_APPLATENCY = MS_APPTRAFFIC + 3,
  MS_APPANOMALY = MS_APPLATENCY + 3,
  MS_COUNT = MS_APPANOMALY + 3
};
static_assert(MS_COUNT < 256, "slot index must fit the interpreter table");

// Latency field number -> slot offset (maxes 0..6, sums 7..13, counts 14..20)
inline int latency_slot(uint32_t f) {
  if (f >= 1 && f <= 6) return (int)f - 1;
  if (f == 19) return 6;
  if (f >= 7 && f <= 12) return (int)f;
  if (f == 20) return 13;
  if (f >= 13 && f <= 18) return (int)f + 1;
  if (f == 21) return 20;
  return -1;
}

inline std::vector<uint32_t> metrics_pb_table() {
  std::vector<uint32_t> t(MM_COUNT * MM_MAXF, 0u);
  auto set = [&](uint32_t m, uint32_t f, uint32_t kind, int slot, uint32_t sub = 0) {
    t[m * MM_MAXF + f] = PB_ENT(kind, slot, sub);
  };
  set(MM_DOC, 1, PK_U32, MS_TS);
  set(MM_DOC, 2, PK_MSG, MS_TAG_CNT, MM_TAG);
  set(MM_DOC, 3, PK_MSG, MS_METER_CNT, MM_METER);
  set(MM_DOC, 4, PK_U32, MS_FLAGS);
  set(MM_TAG, 1, PK_MSG, MS_FIELD_CNT, MM_FIELD);
  set(MM_TAG, 2, PK_U64, MS_CODE);
  const struct { uint32_t f, kind; int slot; } fld[] = {
      {1, PK_STR, MS_IP}, {2, PK_STR, MS_IP1}, {3, PK_U32, MS_GTID}, {4, PK_U32, MS_IS_IPV6}, {5, PK_U32, MS_EPC},
      {6, PK_U32, MS_EPC1}, {7, PK_U64, MS_MAC}, {8, PK_U64, MS_MAC1}, {9, PK_U32, MS_DIRECTION},
      {10, PK_U32, MS_TAP_SIDE}, {11, PK_U32, MS_PROTOCOL}, {12, PK_U32, MS_ACL_GID}, {13, PK_U32, MS_SERVER_PORT},
      {14, PK_U32, MS_VTAP_ID}, {15, PK_U64, MS_TAP_PORT}, {16, PK_U32, MS_TAP_TYPE}, {17, PK_U32, MS_L7_PROTOCOL},
      {20, PK_U32, MS_GPID}, {21, PK_U32, MS_GPID1}, {22, PK_U32, MS_SIGNAL_SOURCE}, {23, PK_STR, MS_APP_SERVICE},
      {24, PK_STR, MS_APP_INSTANCE}, {25, PK_STR, MS_ENDPOINT}, {27, PK_U32, MS_POD_ID}, {28, PK_U32, MS_BIZ_TYPE}};
  for (auto& e : fld) set(MM_FIELD, e.f, e.kind, e.slot);
  set(MM_METER, 1, PK_U32, MS_METER_ID);
  set(MM_METER, 2, PK_MSG, MS_FLOW_CNT, MM_FLOW);
  set(MM_METER, 3, PK_MSG, MS_USAGE_CNT, MM_USAGE);
  set(MM_METER, 4, PK_MSG, MS_APP_CNT, MM_APP);
  const uint32_t flow_sub[5] = {MM_TRAFFIC, MM_LATENCY, MM_PERF, MM_ANOMALY, MM_FLOWLOAD};
  for (uint32_t i = 0; i < 5; i++) set(MM_FLOW, i + 1, PK_MSG, MS_FLOW_SUB + 2 * (int)i, flow_sub[i]);
  for (uint32_t f = 1; f <= 15; f++) set(MM_TRAFFIC, f, f <= 10 ? PK_U64 : PK_U32, MS_TRAFFIC + (int)f - 1);
  for (uint32_t f = 1; f <= 21; f++) {
    const bool sum = (f >= 7 && f <= 12) || f == 20;
    set(MM_LATENCY, f, sum ? PK_U64 : PK_U32, MS_LATENCY + latency_slot(f));
  }
  for (uint32_t f = 1; f <= 6; f++) set(MM_PERF, f, f <= 4 ? PK_U64 : PK_U32, MS_PERF + (int)f - 1);
  for (uint32_t f = 1; f <= 15; f++) set(MM_ANOMALY, f, f <= 12 ? PK_U64 : PK_U32, MS_ANOMALY + (int)f - 1);
  set(MM_ANOMALY, 20, PK_U64, MS_ANOMALY + 15);
  set(MM_ANOMALY, 21, PK_U64, MS_ANOMALY + 16);
  set(MM_FLOWLOAD, 1, PK_U64, MS_FLOWLOAD);
  for (uint32_t f = 1; f <= 8; f++) set(MM_USAGE, f, PK_U64, MS_USAGE + (int)f - 1);
  const uint32_t app_sub[3] = {MM_APPTRAFFIC, MM_APPLATENCY, MM_APPANOMALY};
  for (uint32_t i = 0; i < 3; i++) set(MM_APP, i + 1, PK_MSG, MS_APP_SUB + 2 * (int)i, app_sub[i]);
  for (uint32_t f = 1; f <= 3; f++) set(MM_APPTRAFFIC, f, PK_U32, MS_APPTRAFFIC + (int)f - 1);
  set(MM_APPLATENCY, 1, PK_U32, MS_APPLATENCY);
  set(MM_APPLATENCY, 2, PK_U64, MS_APPLATENCY + 1);
  set(MM_APPLATENCY, 3, PK_U32, MS_APPLATENCY + 2);
  for (uint32_t f = 1; f <= 3; f++) set(MM_APPANOMALY, f, PK_U32, MS_APPANOMALY + (int)f - 1);
  return t;
}

// ---- tokenizer words (pb_tok.cuh) ----------------------------------------------------------------------
// The decode kernel keeps a record as 4-byte words in shared memory: one word for a uint32 slot, two (lo, hi)
// for a uint64 slot, (offset, length) for a bytes / string slot; the occurrence counters and ranges of the
// sub-messages, which nothing reads, share word 0.  mw() maps a slot to its first word with plain arithmetic so
// that it folds to a constant wherever the slot is one.
enum : int {
  MWB_TAG = 1,                    // 38 words, see mw()
  MWB_TRAFFIC = MWB_TAG + 38,     // 10 x u64, 5 x u32
  MWB_LATENCY = MWB_TRAFFIC + 25, // 7 x u32 max, 7 x u64 sum, 7 x u32 count
  MWB_PERF = MWB_LATENCY + 28,    // 4 x u64, 2 x u32
  MWB_ANOMALY = MWB_PERF + 10,    // 12 x u64, 3 x u32, 2 x u64
  MWB_FLOWLOAD = MWB_ANOMALY + 31,
  MWB_USAGE = MWB_FLOWLOAD + 2,   // 8 x u64
  MWB_APP = MWB_USAGE + 16,       // traffic 3, latency 1 + 2 + 1, anomaly 3
  MW_COUNT = MWB_APP + 10
};
DF_HD constexpr bool mw64(int s) {
  return s == MS_CODE || s == MS_IP || s == MS_IP1 || s == MS_MAC || s == MS_MAC1 || s == MS_TAP_PORT ||
         s == MS_APP_SERVICE || s == MS_APP_INSTANCE || s == MS_ENDPOINT || (s >= MS_TRAFFIC && s < MS_TRAFFIC + 10) ||
         (s >= MS_LATENCY + 7 && s < MS_LATENCY + 14) || (s >= MS_PERF && s < MS_PERF + 4) ||
         (s >= MS_ANOMALY && s < MS_ANOMALY + 12) || s == MS_ANOMALY + 15 || s == MS_ANOMALY + 16 || s == MS_FLOWLOAD ||
         (s >= MS_USAGE && s < MS_USAGE + 8) || s == MS_APPLATENCY + 1;
}
DF_HD constexpr int mw(int s) {
  if (s >= MS_TRAFFIC && s < MS_LATENCY) return MWB_TRAFFIC + (s - MS_TRAFFIC < 10 ? 2 * (s - MS_TRAFFIC) : 20 + (s - MS_TRAFFIC - 10));
  if (s >= MS_LATENCY && s < MS_PERF) {
    const int i = s - MS_LATENCY;
    return MWB_LATENCY + (i < 7 ? i : i < 14 ? 7 + 2 * (i - 7) : 21 + (i - 14));
  }
  if (s >= MS_PERF && s < MS_ANOMALY) return MWB_PERF + (s - MS_PERF < 4 ? 2 * (s - MS_PERF) : 8 + (s - MS_PERF - 4));
  if (s >= MS_ANOMALY && s < MS_FLOWLOAD) {
    const int i = s - MS_ANOMALY;
    return MWB_ANOMALY + (i < 12 ? 2 * i : i < 15 ? 24 + (i - 12) : 27 + 2 * (i - 15));
  }
  if (s == MS_FLOWLOAD) return MWB_FLOWLOAD;
  if (s >= MS_USAGE && s < MS_APP_SUB) return MWB_USAGE + 2 * (s - MS_USAGE);
  if (s >= MS_APPTRAFFIC && s < MS_APPLATENCY) return MWB_APP + (s - MS_APPTRAFFIC);
  if (s >= MS_APPLATENCY && s < MS_APPANOMALY) return MWB_APP + 3 + (s == MS_APPLATENCY ? 0 : s == MS_APPLATENCY + 1 ? 1 : 3);
  if (s >= MS_APPANOMALY && s < MS_COUNT) return MWB_APP + 7 + (s - MS_APPANOMALY);
  switch (s) {
    case MS_TS: return MWB_TAG + 0;
    case MS_FLAGS: return MWB_TAG + 1;
    case MS_CODE: return MWB_TAG + 2;
    case MS_IP: return MWB_TAG + 4;
    case MS_IP1: return MWB_TAG + 6;
    case MS_GTID: return MWB_TAG + 8;
    case MS_IS_IPV6: return MWB_TAG + 9;
    case MS_EPC: return MWB_TAG + 10;
    case MS_EPC1: return MWB_TAG + 11;
    case MS_MAC: return MWB_TAG + 12;
    case MS_MAC1: return MWB_TAG + 14;
    case MS_DIRECTION: return MWB_TAG + 16;
    case MS_TAP_SIDE: return MWB_TAG + 17;
    case MS_PROTOCOL: return MWB_TAG + 18;
    case MS_ACL_GID: return MWB_TAG + 19;
    case MS_SERVER_PORT: return MWB_TAG + 20;
    case MS_VTAP_ID: return MWB_TAG + 21;
    case MS_TAP_PORT: return MWB_TAG + 22;
    case MS_TAP_TYPE: return MWB_TAG + 24;
    case MS_L7_PROTOCOL: return MWB_TAG + 25;
    case MS_GPID: return MWB_TAG + 26;
    case MS_GPID1: return MWB_TAG + 27;
    case MS_SIGNAL_SOURCE: return MWB_TAG + 28;
    case MS_APP_SERVICE: return MWB_TAG + 29;
    case MS_APP_INSTANCE: return MWB_TAG + 31;
    case MS_ENDPOINT: return MWB_TAG + 33;
    case MS_POD_ID: return MWB_TAG + 35;
    case MS_BIZ_TYPE: return MWB_TAG + 36;
    case MS_METER_ID: return MWB_TAG + 37;
  }
  return 0;  // occurrence counters / ranges of the sub-messages
}

// pb_tok table of metric.Document from the interpreter table above: same messages, fields and acceptance
// rules, slots replaced by words (string / bytes fields keep offset + length in two words: TK_STR2W)
inline std::vector<uint32_t> metrics_tok_table() {
  const std::vector<uint32_t> src = metrics_pb_table();
  std::vector<uint32_t> t(MM_COUNT * MM_MAXF, 0u);
  for (uint32_t i = 0; i < src.size(); i++) {
    const uint32_t e = src[i];
    if (!e) continue;
    const uint32_t kind = e & 15u, slot = (e >> 8) & 0xffu, sub = (e >> 16) & 0xffu;
    const uint32_t w = (uint32_t)mw((int)slot);
    switch (kind) {
      case PK_U32: t[i] = TK_ENT(TK_U32, w, 0, 0); break;
      case PK_U64: t[i] = TK_ENT(TK_U64, w, 0, 0); break;
      case PK_STR: t[i] = TK_ENT(TK_STR2W, w, 0, 0); break;
      case PK_MSG: t[i] = TK_ENT(TK_MSG, 0, sub, 0); break;
      default: t[i] = 0; break;
    }
  }
  return t;
}

struct MStr {
  const uint8_t* p;
  uint32_t n;
};

// flow_metrics.Tag (tag.go:269-354); member names = dfgpu_metrics_schema.h
struct MTag {
  uint64_t code;
  uint32_t time;
  uint8_t gtid, is_ipv4;
  uint32_t ip4[2];
  uint8_t ip6[2][16];
  uint64_t mac[2];
  int32_t l3_epc[2];
  uint32_t l3_device_id[2];
  uint8_t l3_device_type[2];
  uint16_t region_id[2], subnet_id[2], host_id[2];
  uint32_t pod_node_id[2];
  uint16_t az_id[2];
  uint32_t pod_group_id[2];
  uint16_t pod_ns_id[2];
  uint32_t pod_id[2];
  uint16_t pod_cluster_id[2];
  uint32_t service_id[2];
  uint32_t auto_instance_id[2];
  uint8_t auto_instance_type[2];
  uint32_t auto_service_id[2];
  uint8_t auto_service_type[2];
  uint32_t gpid[2];
  uint16_t acl_gid;
  uint8_t role, protocol;
  uint16_t server_port, vtap_id, org_id, team_id;
  uint32_t tap_port;
  uint8_t tap_port_type, nat_source, tunnel_type;
  uint8_t tap_side, tap_type, is_key_service, l7_protocol, biz_type;
  uint16_t signal_source;
  uint8_t tag_source[2];
  uint16_t tunnel_ip_id;
  MStr str[3];  // app_service, app_instance, endpoint
};

struct MFlowMeter {
  uint64_t packet_tx, packet_rx, byte_tx, byte_rx, l3_byte_tx, l3_byte_rx, l4_byte_tx, l4_byte_rx, new_flow, closed_flow;
  uint32_t l7_request, l7_response, syn_count, synack_count;
  uint8_t direction_score;
  uint32_t rtt_max, rtt_client_max, rtt_server_max, srt_max, art_max, rrt_max, cit_max;
  uint64_t rtt_sum, rtt_client_sum, rtt_server_sum, srt_sum, art_sum, rrt_sum, cit_sum;
  uint32_t rtt_count, rtt_client_count, rtt_server_count, srt_count, art_count, rrt_count, cit_count;
  uint64_t retrans_tx, retrans_rx, zero_win_tx, zero_win_rx;
  uint32_t retrans_syn, retrans_synack;
  uint64_t client_rst_flow, server_rst_flow, server_syn_miss, client_ack_miss, client_half_close_flow,
      server_half_close_flow, client_source_port_reuse, client_establish_reset, server_reset, server_queue_lack,
      server_establish_reset, tcp_timeout, client_establish_fail, server_establish_fail, tcp_establish_fail,
      tcp_transfer_fail, tcp_rst_fail;
  uint32_t l7_client_error, l7_server_error, l7_timeout;
  uint64_t ooo_tx, ooo_rx, flow_load;
};
struct MAppMeter {
  uint32_t request, response;
  uint8_t direction_score;
  uint32_t rrt_max;
  uint64_t rrt_sum;
  uint32_t rrt_count, client_error, server_error, timeout;
};
struct MUsageMeter {
  uint64_t packet_tx, packet_rx, byte_tx, byte_rx, l3_byte_tx, l3_byte_rx, l4_byte_tx, l4_byte_rx;
};

// ---- meter fields in struct order with their ConcurrentMerge rule ------------------
//   S64/S32: wrapping sum; X32/X8: max; FST: kept from the first row of the group (the derived
//   Anomaly *_fail fields are not merged, basic_meter.go:556-577); LST: assigned from the last
//   merged row (Anomaly.OooTx/OooRx :575-576)
enum : uint8_t { MOP_S64 = 0, MOP_S32, MOP_X32, MOP_X8, MOP_FST, MOP_LST };
#define M_FLOW_FIELDS(X)                                                                                          \
  X(packet_tx, MOP_S64) X(packet_rx, MOP_S64) X(byte_tx, MOP_S64) X(byte_rx, MOP_S64) X(l3_byte_tx, MOP_S64)      \
  X(l3_byte_rx, MOP_S64) X(l4_byte_tx, MOP_S64) X(l4_byte_rx, MOP_S64) X(new_flow, MOP_S64)                        \
  X(closed_flow, MOP_S64) X(l7_request, MOP_S32) X(l7_response, MOP_S32) X(syn_count, MOP_S32)                    \
  X(synack_count, MOP_S32) X(direction_score, MOP_X8) X(rtt_max, MOP_X32) X(rtt_client_max, MOP_X32)              \
  X(rtt_server_max, MOP_X32) X(srt_max, MOP_X32) X(art_max, MOP_X32) X(rrt_max, MOP_X32) X(cit_max, MOP_X32)      \
  X(rtt_sum, MOP_S64) X(rtt_client_sum, MOP_S64) X(rtt_server_sum, MOP_S64) X(srt_sum, MOP_S64)                   \
  X(art_sum, MOP_S64) X(rrt_sum, MOP_S64) X(cit_sum, MOP_S64) X(rtt_count, MOP_S32) X(rtt_client_count, MOP_S32)  \
  X(rtt_server_count, MOP_S32) X(srt_count, MOP_S32) X(art_count, MOP_S32) X(rrt_count, MOP_S32)                  \
  X(cit_count, MOP_S32) X(retrans_tx, MOP_S64) X(retrans_rx, MOP_S64) X(zero_win_tx, MOP_S64)                     \
  X(zero_win_rx, MOP_S64) X(retrans_syn, MOP_S32) X(retrans_synack, MOP_S32) X(client_rst_flow, MOP_S64)          \
  X(server_rst_flow, MOP_S64) X(server_syn_miss, MOP_S64) X(client_ack_miss, MOP_S64)                             \
  X(client_half_close_flow, MOP_S64) X(server_half_close_flow, MOP_S64) X(client_source_port_reuse, MOP_S64)      \
  X(client_establish_reset, MOP_S64) X(server_reset, MOP_S64) X(server_queue_lack, MOP_S64)                       \
  X(server_establish_reset, MOP_S64) X(tcp_timeout, MOP_S64) X(client_establish_fail, MOP_FST)                    \
  X(server_establish_fail, MOP_FST) X(tcp_establish_fail, MOP_FST) X(tcp_transfer_fail, MOP_FST)                  \
  X(tcp_rst_fail, MOP_FST) X(l7_client_error, MOP_S32) X(l7_server_error, MOP_S32) X(l7_timeout, MOP_S32)         \
  X(ooo_tx, MOP_LST) X(ooo_rx, MOP_LST) X(flow_load, MOP_S64)
#define M_APP_FIELDS(X)                                                                                           \
  X(request, MOP_S32) X(response, MOP_S32) X(direction_score, MOP_X8) X(rrt_max, MOP_X32) X(rrt_sum, MOP_S64)     \
  X(rrt_count, MOP_S32) X(client_error, MOP_S32) X(server_error, MOP_S32) X(timeout, MOP_S32)
#define M_USAGE_FIELDS(X)                                                                                         \
  X(packet_tx, MOP_S64) X(packet_rx, MOP_S64) X(byte_tx, MOP_S64) X(byte_rx, MOP_S64) X(l3_byte_tx, MOP_S64)      \
  X(l3_byte_rx, MOP_S64) X(l4_byte_tx, MOP_S64) X(l4_byte_rx, MOP_S64)
constexpr int M_MAX_FIELDS = 65;

// ---- output layout ---------------------------------------------------------------------------
enum : uint8_t { MK_FIX = 0, MK_IP6 = 1, MK_OBS = 2, MK_LCS = 3, MK_STR = 4 };
struct MColDesc {
  uint8_t kind, elem, which, pad;
};
constexpr int M_MAX_DESC = 1024;
// scan slots: rows per table, then bytes of (app table a in 0..3, string 0..2)
constexpr int M_SLOT_ROWS = 0;
constexpr int M_SLOT_STR = MT_TABLE_COUNT;  // + a * 3 + which
constexpr int M_SLOTS = MT_TABLE_COUNT + 12;
// app tables: APPLICATION_1M, APPLICATION_MAP_1M, APPLICATION_1S, APPLICATION_MAP_1S -> 0..3 (-1 otherwise)
DF_HD int m_app_index(int table) {
  return table == MT_APPLICATION_1M ? 0 : table == MT_APPLICATION_MAP_1M ? 1 : table == MT_APPLICATION_1S ? 2
       : table == MT_APPLICATION_MAP_1S ? 3 : -1;
}
// observation_point LowCardinality spaces: the four *_MAP tables
DF_HD int m_obs_space(int table) {
  return table == MT_NETWORK_MAP_1M ? 0 : table == MT_APPLICATION_MAP_1M ? 1 : table == MT_NETWORK_MAP_1S ? 2
       : table == MT_APPLICATION_MAP_1S ? 3 : -1;
}

struct MLayout {  // written by k_m_layout, read by k_m_emit and by fetch
  uint64_t col_off[M_MAX_DESC];  // values / keys / string bytes
  uint64_t aux_off[M_MAX_DESC];  // dfgpu_pos array of string columns
  uint64_t str_bytes[12];
  uint32_t nrows[MT_TABLE_COUNT];
  uint32_t row_base[MT_TABLE_COUNT];  // exclusive prefix of nrows: index into MDev::row_rec
  uint32_t overflow;
  uint64_t end;                  // first free byte behind the 9 tables
  // app_service tag table (unmarshaller.go:309-327)
  uint64_t as_off[8];            // time, table keys, svc pos, inst pos, team, svc bytes, inst bytes
  uint64_t as_rows, as_bytes[2], as_end;
  uint32_t as_pending;           // candidates left unresolved after the fixed rounds
  uint32_t as_count, as_unrecorded, as_evicted;  // TagSetState of the app_service set, for the host
  uint32_t as_lc_first[2];
  uint32_t count;                // DocCount: documents that decoded
  uint32_t drop_count;           // DropDocCount
  uint64_t raw_off[MT_TABLE_COUNT];  // raw meters [row][field] of the rollup-enabled tables (device only)
  uint64_t raw_end;
};

struct MCand {  // per record: what the app_service tag stage needs
  uint32_t svc_off, svc_len, inst_off, inst_len;  // offsets into the input arena
  uint32_t time;
  uint32_t slot;   // tag-set slot after the prepare kernel
  uint16_t team;
  uint8_t table;   // 0 application, 1 application_map
  uint8_t state;   // 0 none, 1 pending, 2 emit, 3 resolved (no emit)
};

struct MDev {
  uint32_t tok_tbl[MM_COUNT * MM_MAXF];  // pb_tok table (metrics_tok_table)
  MColDesc descs[M_MAX_DESC];
  uint32_t desc_base[MT_TABLE_COUNT + 1];
  uint32_t ntag[MT_TABLE_COUNT];  // tag columns of the table (the meter columns follow)
  uint64_t table_code[MT_TABLE_COUNT];
  uint8_t table_meter[MT_TABLE_COUNT];
  MLayout* lay;
  MCand* cand;
  uint32_t* row_rec;  // [row_base[table] + row] -> record index (the reduce re-reads raw meters)
  // decode kernel -> emit kernel, all indexed by record (plane-major: [plane][rec_cap])
  uint64_t* tagv;     // values of the tag columns of the record's table, in descriptor order
  uint64_t* rawv;     // raw meter fields (M_*_FIELDS order)
  uint4* strv;        // lengths of app_service / app_instance / endpoint (application tables)
  uint32_t* meta;     // table | meter id << 8 | tap-side code << 16
  uint32_t* tagw;     // decoded tag words (MWB_TAG ...), k_m_decode -> k_m_tag
  uint32_t rec_cap;
  FtEntry* as_set;
  TagSetState* as_state;
  uint32_t as_mask, as_timeout, epoch;
  uint16_t org_id;
  uint32_t raw_tables;            // dfgpu_cfg.rollup_tables: keep raw meters of these tables for the rollup
  uint8_t raw_nf[MT_TABLE_COUNT];  // meter fields per row (M_*_FIELDS)
};

}  // namespace dfgpu
