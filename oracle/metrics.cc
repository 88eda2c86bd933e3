/*
 * ORACLE — test infrastructure only.
 * metrics.cc: see metrics.h for the reference map.
 * Pinned: column names / order / container types and Code / table-code constants (machine-extracted,
 * tests/golden/metrics_columns.json), MiniTag literals + direction tables (tag_test.go:259-392), meter merge KATs
 * (basic_meter_test.go:112-135, usage_meter_test.go:57-79).  Parity unpinned by the reference: DocumentExpand, TableID
 * routing, the meters' AppendToColumnBlock, the reduce — cross-checked by tests/test_kg_crosscheck_cpu.py.
 */
#include "metrics.h"

#include <unordered_map>

#include "kg.h"
#include "l7.h"

namespace oracle {

const MetricSchema& metric_schema() {
  static MetricSchema s;
  return s;
}

// ---- meters ------------------------------------------------------------------------------
void MFlowMeter::ReadFromPB(const Msg& f) {
  const Msg& t = f.sub(1);  // Traffic.ReadFromPB basic_meter.go:75-92
  packet_tx = t.u64(1); packet_rx = t.u64(2); byte_tx = t.u64(3); byte_rx = t.u64(4);
  l3_byte_tx = t.u64(5); l3_byte_rx = t.u64(6); l4_byte_tx = t.u64(7); l4_byte_rx = t.u64(8);
  new_flow = t.u64(9); closed_flow = t.u64(10);
  l7_request = t.u32(11); l7_response = t.u32(12); syn_count = t.u32(13); synack_count = t.u32(14);
  direction_score = (uint8_t)t.u32(15);
  const Msg& l = f.sub(2);  // Latency.ReadFromPB :251-275
  rtt_max = l.u32(1); rtt_client_max = l.u32(2); rtt_server_max = l.u32(3); srt_max = l.u32(4);
  art_max = l.u32(5); rrt_max = l.u32(6); cit_max = l.u32(19);
  rtt_sum = l.u64(7); rtt_client_sum = l.u64(8); rtt_server_sum = l.u64(9); srt_sum = l.u64(10);
  art_sum = l.u64(11); rrt_sum = l.u64(12); cit_sum = l.u64(20);
  rtt_count = l.u32(13); rtt_client_count = l.u32(14); rtt_server_count = l.u32(15); srt_count = l.u32(16);
  art_count = l.u32(17); rrt_count = l.u32(18); cit_count = l.u32(21);
  const Msg& p = f.sub(3);  // Performance.ReadFromPB :408-415
  retrans_tx = p.u64(1); retrans_rx = p.u64(2); zero_win_tx = p.u64(3); zero_win_rx = p.u64(4);
  retrans_syn = p.u32(5); retrans_synack = p.u32(6);
  const Msg& a = f.sub(4);  // Anomaly.ReadFromPB :525-554
  client_rst_flow = a.u64(1); server_rst_flow = a.u64(2); server_syn_miss = a.u64(3); client_ack_miss = a.u64(4);
  client_half_close_flow = a.u64(5); server_half_close_flow = a.u64(6);
  client_source_port_reuse = a.u64(7); client_establish_reset = a.u64(8); server_reset = a.u64(9);
  server_queue_lack = a.u64(10); server_establish_reset = a.u64(11); tcp_timeout = a.u64(12);
  client_establish_fail = client_ack_miss + client_source_port_reuse + client_establish_reset;
  server_establish_fail = server_syn_miss + server_reset + server_queue_lack + server_establish_reset;
  tcp_establish_fail = client_establish_fail + server_establish_fail;
  tcp_transfer_fail = client_rst_flow + server_rst_flow + server_queue_lack + tcp_timeout;
  tcp_rst_fail = client_establish_reset + server_establish_reset + server_reset + client_rst_flow + server_rst_flow;
  l7_client_error = a.u32(13); l7_server_error = a.u32(14); l7_timeout = a.u32(15);
  ooo_tx = a.u64(20); ooo_rx = a.u64(21);
  flow_load = f.sub(5).u64(1);  // FlowLoad.ReadFromPB :703-705
}
#define SUM(x) x += o.x
#define MAX(x) if (x < o.x) x = o.x
void MFlowMeter::ConcurrentMerge(const MFlowMeter& o) {
  // Traffic :94-113
  SUM(packet_tx); SUM(packet_rx); SUM(byte_tx); SUM(byte_rx); SUM(l3_byte_tx); SUM(l3_byte_rx); SUM(l4_byte_tx);
  SUM(l4_byte_rx); SUM(new_flow); SUM(closed_flow); SUM(l7_request); SUM(l7_response); SUM(syn_count);
  SUM(synack_count); MAX(direction_score);
  // Latency :277-315
  MAX(rtt_max); MAX(rtt_client_max); MAX(rtt_server_max); MAX(srt_max); MAX(art_max); MAX(rrt_max); MAX(cit_max);
  SUM(rtt_sum); SUM(rtt_client_sum); SUM(rtt_server_sum); SUM(srt_sum); SUM(art_sum); SUM(rrt_sum); SUM(cit_sum);
  SUM(rtt_count); SUM(rtt_client_count); SUM(rtt_server_count); SUM(srt_count); SUM(art_count); SUM(rrt_count);
  SUM(cit_count);
  // Performance :417-424
  SUM(retrans_tx); SUM(retrans_rx); SUM(zero_win_tx); SUM(zero_win_rx); SUM(retrans_syn); SUM(retrans_synack);
  // Anomaly :556-577 — the derived *_fail fields are NOT merged; Ooo* are assigned
  SUM(client_rst_flow); SUM(server_rst_flow); SUM(server_syn_miss); SUM(client_ack_miss);
  SUM(client_half_close_flow); SUM(server_half_close_flow); SUM(client_source_port_reuse);
  SUM(client_establish_reset); SUM(server_reset); SUM(server_queue_lack); SUM(server_establish_reset);
  SUM(tcp_timeout); SUM(l7_client_error); SUM(l7_server_error); SUM(l7_timeout);
  ooo_tx = o.ooo_tx;
  ooo_rx = o.ooo_rx;
  SUM(flow_load);  // FlowLoad :707-709
}
void MAppMeter::ReadFromPB(const Msg& a) {
  request = a.sub(1).u32(1); response = a.sub(1).u32(2); direction_score = (uint8_t)a.sub(1).u32(3);  // :147-151
  rrt_max = a.sub(2).u32(1); rrt_sum = a.sub(2).u64(2); rrt_count = a.sub(2).u32(3);                  // :201-205
  client_error = a.sub(3).u32(1); server_error = a.sub(3).u32(2); timeout = a.sub(3).u32(3);          // :255-259
}
void MAppMeter::ConcurrentMerge(const MAppMeter& o) {
  SUM(request); SUM(response); MAX(direction_score);  // :153-159
  MAX(rrt_max); SUM(rrt_sum); SUM(rrt_count);          // :207-213
  SUM(client_error); SUM(server_error); SUM(timeout);  // :261-265
}
void MUsageMeter::ReadFromPB(const Msg& u) {
  packet_tx = u.u64(1); packet_rx = u.u64(2); byte_tx = u.u64(3); byte_rx = u.u64(4);
  l3_byte_tx = u.u64(5); l3_byte_rx = u.u64(6); l4_byte_tx = u.u64(7); l4_byte_rx = u.u64(8);
}
void MUsageMeter::ConcurrentMerge(const MUsageMeter& o) {
  SUM(packet_tx); SUM(packet_rx); SUM(byte_tx); SUM(byte_rx); SUM(l3_byte_tx); SUM(l3_byte_rx); SUM(l4_byte_tx);
  SUM(l4_byte_rx);
}
#undef SUM
#undef MAX

// ---- Tag.ReadFromPB tag.go:1176-1240 ----------------------------------------------------------
static uint8_t ToRole(uint32_t direction) {  // DirectionEnum.ToRole tag.go:174-185 (DirectionEnum is a uint8)
  switch ((uint8_t)direction & 0x7) {
    case 1: return 0;  // ROLE_CLIENT
    case 2: return 1;  // ROLE_SERVER
    case 4: return 2;  // ROLE_LOCAL
  }
  return 3;  // ROLE_REST
}
// p.Field.Ip[:n]: agents always send 4 or 16 bytes; a shorter slice would panic (or read
// the pooled buffer's stale capacity) in Go — restated as zero extension, see DESIGN.md.
static void ip_bytes(std::string_view s, uint8_t* out, size_t n) {
  memset(out, 0, n);
  memcpy(out, s.data(), s.size() < n ? s.size() : n);
}
void TagReadFromPB(MTag& t, const Msg& p) {
  const Msg& f = p.sub(1);
  t.code = p.u64(2);
  t.gtid = (uint8_t)f.u32(3);
  t.is_ipv4 = (uint8_t)(1 - (uint8_t)f.u32(4));
  if (t.is_ipv4 == 0) {
    ip_bytes(f.str(1), t.ip6[0], 16);
    if (t.code & MC_IPPATH) ip_bytes(f.str(2), t.ip6[1], 16);
  } else {
    uint8_t b[4];
    ip_bytes(f.str(1), b, 4);
    t.ip4[0] = (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3];
    if (t.code & MC_IPPATH) {
      ip_bytes(f.str(2), b, 4);
      t.ip4[1] = (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3];
    }
  }
  t.mac[0] = f.u64(7);
  t.mac[1] = f.u64(8);
  t.l3_epc[0] = MarshalInt32WithSpecialID(f.i32(5));
  t.l3_epc[1] = MarshalInt32WithSpecialID(f.i32(6));
  t.role = ToRole(f.u32(9));
  t.tap_side = (uint8_t)f.u32(10);
  t.protocol = (uint8_t)f.u32(11);
  t.acl_gid = (uint16_t)f.u32(12);
  t.server_port = (uint16_t)f.u32(13);
  t.vtap_id = (uint16_t)f.u32(14);
  uint64_t tp = f.u64(15);  // SplitToPortTypeTunnel datatype/tap_port.go:113-115
  t.tap_port = (uint32_t)tp;
  t.tap_port_type = (uint8_t)(tp >> 60);
  t.nat_source = (uint8_t)((tp >> 36) & 0xf);
  t.tunnel_type = (uint8_t)((tp >> 32) & 0xf);
  t.tap_type = (uint8_t)f.u32(16);
  t.l7_protocol = (uint8_t)f.u32(17);
  t.app_service = std::string(f.str(23));
  t.app_instance = std::string(f.str(24));
  t.endpoint = std::string(f.str(25));
  t.biz_type = (uint8_t)f.u32(28);
  if (t.code & MC_IPPATH) t.code |= MC_GPIDPATH | MC_SIGNALSOURCE;  // :1218-1225
  else if (t.code != MC_TRAFFIC_POLICY) t.code |= MC_GPID | MC_SIGNALSOURCE;
  t.gpid[0] = f.u32(20);
  t.gpid[1] = f.u32(21);
  if (f.u32(27) != 0) {  // :1229-1235
    if ((t.code & MC_IPPATH) && t.role == 1) t.pod_id[1] = f.u32(27);
    else t.pod_id[0] = f.u32(27);
  }
  t.signal_source = (uint16_t)f.u32(22);
  t.tunnel_ip_id = (uint16_t)f.u32(13);
}

// ---- getPlatformInfos handle_document.go:41-148 -------------------------------------------------
static bool side_info(MTag& t, int s, const PlatformInfoTable* pd, Info& info) {
  bool has = false;
  const int o = 1 - s;
  const bool v6 = t.is_ipv4 == 0;
  if (IsLoopback(v6, t.ip4[s], t.ip6[s], 16)) {
    if (t.pod_id[s] == 0 && t.pod_id[o] != 0) {
      t.tag_source[s] |= TS_Peer;
      t.pod_id[s] = t.pod_id[o];
    }
  }
  if (t.gpid[s] != 0 && t.pod_id[s] == 0) {
    uint16_t vtap = 0;
    uint32_t pod = 0;
    pd->QueryGprocessInfo(t.gpid[s], vtap, pod);
    if (pod != 0 && vtap == t.vtap_id) {
      t.pod_id[s] = pod;
      t.tag_source[s] |= TS_GpId;
    }
  }
  if (t.pod_id[s] != 0) {
    has = pd->QueryPodIdInfo(t.pod_id[s], info);
    t.tag_source[s] |= TS_PodId;
  }
  if (!has) {
    if (t.mac[s] != 0) {
      t.tag_source[s] |= TS_Mac;
      has = pd->QueryMacInfo(t.mac[s] | (uint64_t)(int64_t)t.l3_epc[s] << 48, info);
    }
    if (!has) {
      bool lookupByAgent = false;
      if (t.tap_side == TAP_Local || t.tap_side == TAP_ClientProcess || t.tap_side == TAP_ServerProcess)
        if (IsLocalIP(v6, t.ip4[s], t.ip6[s], 16)) lookupByAgent = true;
      if (lookupByAgent) {
        t.tag_source[s] |= TS_Agent;
        VtapInfo vi;
        if (pd->QueryVtapInfo(t.vtap_id, vi))
          has = RegetInfoFromIP(pd, !vi.IsIPv4, vi.IP6v.data(), vi.has_ip6 ? 16 : 0, vi.IP4, vi.EpcId, info);
      } else {
        t.tag_source[s] |= TS_EpcIP;
        has = RegetInfoFromIP(pd, v6, t.ip6[s], 16, t.ip4[s], t.l3_epc[s], info);
      }
    }
  }
  return has;
}

static bool is_multicast(const MTag& t, int s) {  // net.IP.IsMulticast
  if (t.is_ipv4 == 0) return t.ip6[s][0] == 0xff;
  return (t.ip4[s] >> 28) == 0xe;
}

static int DatabaseSuffixID(uint64_t code) {  // tag.go:1283-1295
  int c = 0;
  if (code & MC_ACLGID) c |= 1;
  if (code & MC_EDGE_FIELDS) c |= 2;
  if (code & MC_SERVERPORT) c |= 4;
  return c;
}

bool DocumentExpand(MDoc& d, const PlatformInfoTable* pd) {
  MTag& t = d.t;
  if (d.meter_id == MT_METER_USAGE && DatabaseSuffixID(t.code) == 1) return true;  // :154-157
  const uint16_t myRegionID = (uint16_t)pd->regionID;
  if ((t.code & MC_SERVERPORT) == MC_SERVERPORT) t.code |= MC_PORTADDCODE;
  const bool edge = (t.code & MC_EDGECODE) == MC_EDGECODE;
  const bool v6 = t.is_ipv4 == 0;
  Info info, info1;
  bool has = false, has1 = false;
  if (t.l3_epc[0] != EPC_FROM_INTERNET) has = side_info(t, 0, pd, info);
  if (edge && t.l3_epc[1] != EPC_FROM_INTERNET) has1 = side_info(t, 1, pd, info1);
  if (edge) t.code |= MC_EDGEADDCODE;
  else t.code |= MC_MAINADDCODE;

  uint8_t podGroupType = 0, podGroupType1 = 0;
  const bool defaultOrg = t.org_id <= 1;  // ckdb.IsDefaultOrgID
  if (has1) {  // :176-213
    t.region_id[1] = (uint16_t)info1.RegionID;
    t.host_id[1] = (uint16_t)info1.HostID;
    t.l3_device_id[1] = info1.DeviceID;
    t.l3_device_type[1] = (uint8_t)info1.DeviceType;
    t.subnet_id[1] = (uint16_t)info1.SubnetID;
    t.pod_node_id[1] = info1.PodNodeID;
    t.pod_ns_id[1] = (uint16_t)info1.PodNSID;
    t.az_id[1] = (uint16_t)info1.AZID;
    t.pod_group_id[1] = info1.PodGroupID;
    podGroupType1 = info1.PodGroupType;
    t.pod_id[1] = info1.PodID;
    t.pod_cluster_id[1] = (uint16_t)info1.PodClusterID;
    if (IsPodServiceIP(t.l3_device_type[1], t.pod_id[1], t.pod_node_id[1]))
      t.service_id[1] = pd->services.QueryPodService(pd, t.pod_id[1], t.pod_node_id[1], t.pod_cluster_id[1],
                                                     t.pod_group_id[1], t.l3_epc[1], v6, t.ip4[1], t.ip6[1], 16,
                                                     t.protocol, t.server_port);
    if (!has && is_multicast(t, 0)) {
      t.region_id[0] = t.region_id[1];
      t.subnet_id[0] = t.subnet_id[1];
      t.az_id[0] = t.az_id[1];
      t.tag_source[0] |= TS_Peer;
    }
    if (defaultOrg && myRegionID != 0 && t.region_id[1] != 0)
      if (t.tap_side == TAP_Server && t.region_id[1] != myRegionID) return false;
  }
  GetAutoInstance(t.pod_id[1], t.gpid[1], t.pod_node_id[1], t.l3_device_id[1], t.subnet_id[1], t.l3_device_type[1],
                  t.l3_epc[1], t.auto_instance_id[1], t.auto_instance_type[1]);
  uint32_t custom1 = pd->services.QueryCustomService(t.l3_epc[1], v6, t.ip4[1], t.ip6[1], 16, t.server_port,
                                                     t.service_id[1], t.pod_group_id[1], t.l3_device_id[1],
                                                     t.pod_id[1], t.pod_cluster_id[1], t.l3_device_type[1], t.protocol);
  GetAutoService(custom1, t.service_id[1], t.pod_group_id[1], t.gpid[1], t.pod_cluster_id[1], t.l3_device_id[1],
                 t.subnet_id[1], t.l3_device_type[1], podGroupType1, t.l3_epc[1], t.auto_service_id[1],
                 t.auto_service_type[1]);

  if (has) {  // :218-268
    t.region_id[0] = (uint16_t)info.RegionID;
    t.host_id[0] = (uint16_t)info.HostID;
    t.l3_device_id[0] = info.DeviceID;
    t.l3_device_type[0] = (uint8_t)info.DeviceType;
    t.subnet_id[0] = (uint16_t)info.SubnetID;
    t.pod_node_id[0] = info.PodNodeID;
    t.pod_ns_id[0] = (uint16_t)info.PodNSID;
    t.az_id[0] = (uint16_t)info.AZID;
    t.pod_group_id[0] = info.PodGroupID;
    podGroupType = info.PodGroupType;
    t.pod_id[0] = info.PodID;
    t.pod_cluster_id[0] = (uint16_t)info.PodClusterID;
    if (IsPodServiceIP(t.l3_device_type[0], t.pod_id[0], t.pod_node_id[0])) {
      if (t.server_port > 0 && !edge)
        t.service_id[0] = pd->services.QueryPodService(pd, t.pod_id[0], t.pod_node_id[0], t.pod_cluster_id[0],
                                                       t.pod_group_id[0], t.l3_epc[0], v6, t.ip4[0], t.ip6[0], 16,
                                                       t.protocol, t.server_port);
      else if (IsPodServiceIP(t.l3_device_type[0], t.pod_id[0], 0))
        t.service_id[0] = pd->services.QueryPodService(pd, t.pod_id[0], t.pod_node_id[0], t.pod_cluster_id[0],
                                                       t.pod_group_id[0], t.l3_epc[0], v6, t.ip4[0], t.ip6[0], 16,
                                                       t.protocol, 0);
    }
    if (!has1 && edge && is_multicast(t, 1)) {
      t.region_id[1] = t.region_id[0];
      t.subnet_id[1] = t.subnet_id[0];
      t.az_id[1] = t.az_id[0];
      t.tag_source[1] |= TS_Peer;
    }
    if (defaultOrg && myRegionID != 0 && t.region_id[0] != 0) {
      if (edge) {
        if (t.tap_side == TAP_Client && t.region_id[0] != myRegionID) return false;
      } else if (t.region_id[0] != myRegionID) {
        return false;
      }
    }
  }
  GetAutoInstance(t.pod_id[0], t.gpid[0], t.pod_node_id[0], t.l3_device_id[0], t.subnet_id[0], t.l3_device_type[0],
                  t.l3_epc[0], t.auto_instance_id[0], t.auto_instance_type[0]);
  uint16_t serverPort = edge ? 0 : t.server_port;
  uint32_t custom0 = pd->services.QueryCustomService(t.l3_epc[0], v6, t.ip4[0], t.ip6[0], 16, serverPort,
                                                     t.service_id[0], t.pod_group_id[0], t.l3_device_id[0],
                                                     t.pod_id[0], t.pod_cluster_id[0], t.l3_device_type[0], t.protocol);
  GetAutoService(custom0, t.service_id[0], t.pod_group_id[0], t.gpid[0], t.pod_cluster_id[0], t.l3_device_id[0],
                 t.subnet_id[0], t.l3_device_type[0], podGroupType, t.l3_epc[0], t.auto_service_id[0],
                 t.auto_service_type[0]);

  if (t.signal_source == 4) {  // SIGNAL_SOURCE_OTEL :280-303
    if (t.tap_side == 50 /*ServerApp*/ && t.server_port == 0) t.server_port = 65535;
    if (t.l3_epc[0] == EPC_FROM_INTERNET) t.l3_epc[0] = EPC_UNKNOWN;
    if (t.l3_epc[1] == EPC_FROM_INTERNET) t.l3_epc[1] = EPC_UNKNOWN;
    for (int s = 0; s < 2; s++) {
      if (t.auto_service_type[s] == 0) t.auto_service_type[s] = 255;    // InternetIpType -> IpType
      if (t.auto_instance_type[s] == 0) t.auto_instance_type[s] = 255;
    }
  }
  return true;
}

uint64_t metrics_table_code(int table) {
  switch (table) {
    case MT_NETWORK_1M: case MT_NETWORK_1S: return MC_NETWORK;
    case MT_NETWORK_MAP_1M: case MT_NETWORK_MAP_1S: return MC_NETWORK_MAP;
    case MT_APPLICATION_1M: case MT_APPLICATION_1S: return MC_APPLICATION;
    case MT_APPLICATION_MAP_1M: case MT_APPLICATION_MAP_1S: return MC_APPLICATION_MAP;
    case MT_TRAFFIC_POLICY_1M: return MC_TRAFFIC_POLICY;
  }
  return 0;
}
int metrics_table_meter(int table) {
  uint64_t c = metrics_table_code(table);
  if (c == MC_TRAFFIC_POLICY) return MT_METER_USAGE;
  if (c & MC_L7PROTOCOL) return MT_METER_APP;
  return MT_METER_FLOW;
}
int TagTableID(uint64_t code, bool isSecond) {
  code &= ~(MC_MAC | MC_MACPATH);
  for (int i = 0; i < MT_TABLE_COUNT; i++)  // first match over metricsTableCodes
    if (code == metrics_table_code(i)) {
      // a per-second traffic_policy document would get id 9 == METRICS_TABLE_ID_MAX and
      // index past dbwriter's per-table caches (dbwriter.go:78-95): treated as "no table"
      int id = isSecond ? i + MT_NETWORK_1S : i;
      return id < MT_TABLE_COUNT ? id : -1;
    }
  return -1;
}

// ---- column blocks -------------------------------------------------------------------------------
void MetricsBlocks::add_columns(Block& b, int table) {
  const uint64_t code = metrics_table_code(table);
#define F(bit, name, ch, ctype, expr) if ((bit) == 0 || (code & (bit))) b.add(name, ch, sizeof(ctype));
#define V6(bit, name, side) if (code & (bit)) b.add(name, DFGPU_CH_IPV6, 16);
#define OBS(bit, name) if (code & (bit)) b.add(name, DFGPU_CH_LC_STR, 4);
#define LCS(bit, name, which) if (code & (bit)) b.add(name, DFGPU_CH_LC_STR_ROWS);
#define STR(bit, name, which) if (code & (bit)) b.add(name, DFGPU_CH_STR);
  MT_TAG_COLUMNS(F, V6, OBS, LCS, STR)
#undef F
#undef V6
#undef OBS
#undef LCS
#undef STR
#define M(name, ch, ctype, expr) b.add(name, ch, sizeof(ctype));
  switch (metrics_table_meter(table)) {
    case MT_METER_FLOW: MT_FLOW_METER_COLUMNS(M) break;
    case MT_METER_APP: MT_APP_METER_COLUMNS(M) break;
    case MT_METER_USAGE: MT_USAGE_METER_COLUMNS(M) break;
  }
#undef M
}

void MetricsBlocks::append_row(Block& b, int table, const MDoc& d) {
  const uint64_t code = metrics_table_code(table);
  const MTag& t = d.t;
  size_t c = 0;
  const std::string* strs[3] = {&t.app_service, &t.app_instance, &t.endpoint};
#define F(bit, name, ch, ctype, expr) if ((bit) == 0 || (code & (bit))) b[c++].append<ctype>((ctype)(expr));
#define V6(bit, name, side) if (code & (bit)) b[c++].append_ip6(t.ip6[side], 16);
#define OBS(bit, name)                                 \
  if (code & (bit)) {                                  \
    bool ok;                                           \
    b[c++].append_lc(TAPSideString(t.tap_side, ok));   \
  }
#define LCS(bit, name, which) if (code & (bit)) b[c++].append_str(*strs[which]);
#define STR(bit, name, which) if (code & (bit)) b[c++].append_str(*strs[which]);
  MT_TAG_COLUMNS(F, V6, OBS, LCS, STR)
#undef F
#undef V6
#undef OBS
#undef LCS
#undef STR
#define M(name, ch, ctype, expr) b[c++].append<ctype>((ctype)(expr));
  switch (metrics_table_meter(table)) {
    case MT_METER_FLOW: { const MFlowMeter& m = d.flow; MT_FLOW_METER_COLUMNS(M) break; }
    case MT_METER_APP: { const MAppMeter& m = d.app; MT_APP_METER_COLUMNS(M) break; }
    case MT_METER_USAGE: { const MUsageMeter& m = d.usage; MT_USAGE_METER_COLUMNS(M) break; }
  }
#undef M
  b.nrows++;
}

void MetricsBlocks::init() {
  for (int i = 0; i < MT_TABLE_COUNT; i++) add_columns(tables[i], i);
  app_service.add("time", DFGPU_CH_DATETIME, 4);  // app_service_tag_column_block.go:39-47
  app_service.add("table", DFGPU_CH_LC_STR, 4);
  app_service.add("app_service", DFGPU_CH_LC_STR_ROWS);
  app_service.add("app_instance", DFGPU_CH_LC_STR_ROWS);
  app_service.add("team_id", DFGPU_CH_UINT16, 2);
}
void MetricsBlocks::clear() {
  for (auto& t : tables) t.clear();
  for (auto& r : rows) r.clear();
  app_service.clear();
}

void MetricsBlocks::append_record(const Msg& m, uint16_t orgId, uint16_t teamId, const PlatformInfoTable* pd,
                                  dfgpu_counters& cnt, L7State& st) {
  MDoc d;
  d.meter_id = (uint8_t)m.sub(3).u32(1);  // codec.go:38
  d.t.time = m.u32(1);
  d.flags = m.u32(4);
  TagReadFromPB(d.t, m.sub(2));
  switch (d.meter_id) {
    case MT_METER_FLOW: d.flow.ReadFromPB(m.sub(3).sub(2)); break;
    case MT_METER_USAGE: d.usage.ReadFromPB(m.sub(3).sub(3)); break;
    case MT_METER_APP: d.app.ReadFromPB(m.sub(3).sub(4)); break;
  }
  d.t.team_id = teamId;  // unmarshaller.go:242-243
  d.t.org_id = orgId;
  if (!DocumentExpand(d, pd)) {  // :253-258
    cnt.drop_count++;
    return;
  }
  int table = TagTableID(d.t.code, (d.flags & 1) == 1);  // :260-266
  // A document whose meter is not the one its table stores cannot be packed by the
  // reference either (type assertion in AppendToColumnBlock); counted as dropped here.
  if (table < 0 || metrics_table_meter(table) != d.meter_id) {
    cnt.drop_count++;
    return;
  }
  // appServiceTagWrite :309-327 + AppServiceTagWriter.Write app_service_tag_writer.go:92-113
  if ((table == MT_APPLICATION_1M || table == MT_APPLICATION_MAP_1M) &&
      !(d.t.app_service.empty() && d.t.app_instance.empty())) {
    AppServiceTagKey k;
    k.Table = table == MT_APPLICATION_1M ? "application" : "application_map";
    k.AppService = d.t.app_service;
    k.AppInstance = d.t.app_instance;
    k.OrgId = orgId;
    k.TeamID = teamId;
    bool emit = true;
    auto r = st.AppServiceCache.AddOrGet(k, d.t.time);
    if (r.second) {
      if ((uint32_t)(r.first + st.cfg.flow_tag_cache_flush_timeout) >= d.t.time) emit = false;
      else st.AppServiceCache.Add(k, d.t.time);
    }
    if (emit) {
      size_t a = 0;
      app_service[a++].append<uint32_t>(d.t.time);
      app_service[a++].append_lc(k.Table);
      app_service[a++].append_str(k.AppService);
      app_service[a++].append_str(k.AppInstance);
      app_service[a++].append<uint16_t>(teamId);
      app_service.nrows++;
    }
  }
  append_row(tables[table], table, d);
  rows[table].push_back(std::move(d));
}

// ---- A9: reduce-by-key with ConcurrentMerge ------------------------------------------------------
// Groups the rows of one table by (time bucket, every tag column of the table) in
// first-appearance order and folds the meters of each group in input order.
static std::string group_key(const MDoc& d, uint64_t code, uint32_t bucket) {
  std::string k;
  const MTag& t = d.t;
  const std::string* strs[3] = {&t.app_service, &t.app_instance, &t.endpoint};
  auto put = [&](const void* p, size_t n) { k.append((const char*)p, n); };
  int idx = 0;
#define F(bit, name, ch, ctype, expr)                 \
  if ((bit) == 0 || (code & (bit))) {                 \
    ctype v = (ctype)(expr);                          \
    if (idx == 0) v = (ctype)bucket;                  \
    put(&v, sizeof v);                                \
    idx++;                                            \
  }
#define V6(bit, name, side) if (code & (bit)) put(t.ip6[side], 16);
#define OBS(bit, name) if (code & (bit)) put(&t.tap_side, 1);
#define LCS(bit, name, which)                            \
  if (code & (bit)) {                                    \
    uint32_t n = (uint32_t)strs[which]->size();          \
    put(&n, 4);                                          \
    put(strs[which]->data(), n);                         \
  }
#define STR(bit, name, which) LCS(bit, name, which)
  MT_TAG_COLUMNS(F, V6, OBS, LCS, STR)
#undef F
#undef V6
#undef OBS
#undef LCS
#undef STR
  return k;
}

// The rollup accumulator: the same grouping and merge, with the groups kept across submits
// (a 1s -> 1m rollup folds every submit of the minute) until rollup_flush emits them.
int MetricsBlocks::rollup_add(int table, uint32_t bucket_seconds) {
  if (table < 0 || table >= MT_TABLE_COUNT || bucket_seconds == 0) return DFGPU_EINVAL;
  const uint64_t code = metrics_table_code(table);
  auto& idx = acc_idx[table];
  auto& out = acc_out[table];
  for (const MDoc& d : rows[table]) {
    uint32_t bucket = d.t.time / bucket_seconds * bucket_seconds;
    std::string k = group_key(d, code, bucket);
    auto it = idx.find(k);
    if (it == idx.end()) {
      idx.emplace(std::move(k), out.size());
      out.push_back(d);
      out.back().t.time = bucket;
    } else {
      MDoc& a = out[it->second];
      a.flow.ConcurrentMerge(d.flow);
      a.app.ConcurrentMerge(d.app);
      a.usage.ConcurrentMerge(d.usage);
    }
  }
  return 0;
}
int MetricsBlocks::rollup_flush(int table) {
  if (table < 0 || table >= MT_TABLE_COUNT) return DFGPU_EINVAL;
  reduced = Block();
  add_columns(reduced, table);
  for (const MDoc& d : acc_out[table]) append_row(reduced, table, d);
  acc_out[table].clear();
  acc_idx[table].clear();
  return 0;
}

int MetricsBlocks::reduce(int table, uint32_t bucket_seconds) {
  if (table < 0 || table >= MT_TABLE_COUNT || bucket_seconds == 0) return DFGPU_EINVAL;
  reduced = Block();
  add_columns(reduced, table);
  const uint64_t code = metrics_table_code(table);
  std::unordered_map<std::string, size_t> idx;
  std::vector<MDoc> out;
  for (const MDoc& d : rows[table]) {
    uint32_t bucket = d.t.time / bucket_seconds * bucket_seconds;
    std::string k = group_key(d, code, bucket);
    auto it = idx.find(k);
    if (it == idx.end()) {
      idx.emplace(std::move(k), out.size());
      out.push_back(d);
      out.back().t.time = bucket;
    } else {
      MDoc& a = out[it->second];
      a.flow.ConcurrentMerge(d.flow);
      a.app.ConcurrentMerge(d.app);
      a.usage.ConcurrentMerge(d.usage);
    }
  }
  for (const MDoc& d : out) append_row(reduced, table, d);
  return 0;
}

}  // namespace oracle
