/*
 * metrics.cu — the flow_metrics kernels: metric.Document record -> one row of one of the
 * 9 flow_metrics tables (SURVEY A10 + A5 twin + A8), the unmarshaller's app_service tags,
 * and the reduce-by-key operator (A9).  One thread per record, warp = 32 consecutive
 * records.  Replaces, fused:
 *   decode         app.DecodePB  server/libs/app/codec.go:28-65 (gogo Unmarshal of pb.Document,
 *                  pre-allocated sub-messages: flow-metrics/pb/doc.go:19-106)
 *   tag            Tag.ReadFromPB  flow-metrics/tag.go:1176-1240
 *   meters         FlowMeter/AppMeter/UsageMeter.ReadFromPB (basic_meter.go:75-92,251-275,408-415,
 *                  525-554,703-705; app_meter.go:147-151,201-205,255-259; usage_meter.go)
 *   SmartEncoding  getPlatformInfos + DocumentExpand  unmarshaller/handle_document.go:41-306
 *   routing        Tag.TableID tag.go:920-931, QueueProcess unmarshaller.go:215-275
 *   column pack    Tag.AppendToColumnBlock tag_column_block.go:521-721 + the meter blocks
 *   app_service    appServiceTagWrite unmarshaller.go:309-327 -> AppServiceTagWriter.Write
 *                  flow_tag/app_service_tag_writer.go:92-113
 *   reduce         *.ConcurrentMerge basic_meter.go:94-113,277-315,417-424,556-577,707-709
 */
#include <algorithm>
#include <cstring>

#include "metrics_host.h"

namespace dfgpu {

namespace {

constexpr int AS_ROUNDS = 2;

// ---- Tag.ReadFromPB -------------------------------------------------------------------------
__device__ __forceinline__ uint8_t to_role(uint32_t direction) {  // tag.go:174-185
  switch (direction & 0x7) {
    case 1: return 0;
    case 2: return 1;
    case 4: return 2;
  }
  return 3;
}
// p.Field.Ip[:n] — agents send exactly 4 / 16 bytes; shorter input is zero extended
__device__ __forceinline__ void ip_bytes(const uint8_t* base, uint64_t ref, uint8_t* out, uint32_t n) {
  const uint32_t off = (uint32_t)ref, len = (uint32_t)(ref >> 32);
  for (uint32_t i = 0; i < n; i++) out[i] = i < len ? __ldg(base + off + i) : 0;
}

template <class SL>
__device__ __forceinline__ void m_read_tag(const SL& sl, const uint8_t* base, MTag& t) {
  memset(&t, 0, sizeof t);
  t.code = sl[MS_CODE];
  t.time = sl_u32(sl, MS_TS);
  t.gtid = (uint8_t)sl_u32(sl, MS_GTID);
  t.is_ipv4 = (uint8_t)(1 - (uint8_t)sl_u32(sl, MS_IS_IPV6));
  if (t.is_ipv4 == 0) {
    ip_bytes(base, sl[MS_IP], t.ip6[0], 16);
    if (t.code & MC_IPPATH) ip_bytes(base, sl[MS_IP1], t.ip6[1], 16);
  } else {
    uint8_t b[4];
    ip_bytes(base, sl[MS_IP], b, 4);
    t.ip4[0] = (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3];
    if (t.code & MC_IPPATH) {
      ip_bytes(base, sl[MS_IP1], b, 4);
      t.ip4[1] = (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3];
    }
  }
  t.mac[0] = sl[MS_MAC];
  t.mac[1] = sl[MS_MAC1];
  t.l3_epc[0] = marshal_epc(sl_i32(sl, MS_EPC));
  t.l3_epc[1] = marshal_epc(sl_i32(sl, MS_EPC1));
  t.role = to_role(sl_u32(sl, MS_DIRECTION) & 0xff);
  t.tap_side = (uint8_t)sl_u32(sl, MS_TAP_SIDE);
  t.protocol = (uint8_t)sl_u32(sl, MS_PROTOCOL);
  t.acl_gid = (uint16_t)sl_u32(sl, MS_ACL_GID);
  t.server_port = (uint16_t)sl_u32(sl, MS_SERVER_PORT);
  t.vtap_id = (uint16_t)sl_u32(sl, MS_VTAP_ID);
  const uint64_t tp = sl[MS_TAP_PORT];  // SplitToPortTypeTunnel datatype/tap_port.go:113-115
  t.tap_port = (uint32_t)tp;
  t.tap_port_type = (uint8_t)(tp >> 60);
  t.nat_source = (uint8_t)((tp >> 36) & 0xf);
  t.tunnel_type = (uint8_t)((tp >> 32) & 0xf);
  t.tap_type = (uint8_t)sl_u32(sl, MS_TAP_TYPE);
  t.l7_protocol = (uint8_t)sl_u32(sl, MS_L7_PROTOCOL);
  t.str[0] = MStr{base + sl_off(sl, MS_APP_SERVICE), sl_len(sl, MS_APP_SERVICE)};
  t.str[1] = MStr{base + sl_off(sl, MS_APP_INSTANCE), sl_len(sl, MS_APP_INSTANCE)};
  t.str[2] = MStr{base + sl_off(sl, MS_ENDPOINT), sl_len(sl, MS_ENDPOINT)};
  t.biz_type = (uint8_t)sl_u32(sl, MS_BIZ_TYPE);
  if (t.code & MC_IPPATH) t.code |= MC_GPIDPATH | MC_SIGNALSOURCE;  // :1218-1225
  else if (t.code != MC_TRAFFIC_POLICY) t.code |= MC_GPID | MC_SIGNALSOURCE;
  t.gpid[0] = sl_u32(sl, MS_GPID);
  t.gpid[1] = sl_u32(sl, MS_GPID1);
  const uint32_t pod = sl_u32(sl, MS_POD_ID);
  if (pod != 0) {  // :1229-1235
    if ((t.code & MC_IPPATH) && t.role == 1) t.pod_id[1] = pod;
    else t.pod_id[0] = pod;
  }
  t.signal_source = (uint16_t)sl_u32(sl, MS_SIGNAL_SOURCE);
  t.tunnel_ip_id = (uint16_t)sl_u32(sl, MS_SERVER_PORT);
}

template <class SL>
__device__ __forceinline__ void m_read_flow(const SL& sl, MFlowMeter& m) {
  const auto t = sl.at(MS_TRAFFIC);
  m.packet_tx = t[0]; m.packet_rx = t[1]; m.byte_tx = t[2]; m.byte_rx = t[3]; m.l3_byte_tx = t[4];
  m.l3_byte_rx = t[5]; m.l4_byte_tx = t[6]; m.l4_byte_rx = t[7]; m.new_flow = t[8]; m.closed_flow = t[9];
  m.l7_request = (uint32_t)t[10]; m.l7_response = (uint32_t)t[11]; m.syn_count = (uint32_t)t[12];
  m.synack_count = (uint32_t)t[13]; m.direction_score = (uint8_t)t[14];
  const auto l = sl.at(MS_LATENCY);
  m.rtt_max = (uint32_t)l[0]; m.rtt_client_max = (uint32_t)l[1]; m.rtt_server_max = (uint32_t)l[2];
  m.srt_max = (uint32_t)l[3]; m.art_max = (uint32_t)l[4]; m.rrt_max = (uint32_t)l[5]; m.cit_max = (uint32_t)l[6];
  m.rtt_sum = l[7]; m.rtt_client_sum = l[8]; m.rtt_server_sum = l[9]; m.srt_sum = l[10]; m.art_sum = l[11];
  m.rrt_sum = l[12]; m.cit_sum = l[13];
  m.rtt_count = (uint32_t)l[14]; m.rtt_client_count = (uint32_t)l[15]; m.rtt_server_count = (uint32_t)l[16];
  m.srt_count = (uint32_t)l[17]; m.art_count = (uint32_t)l[18]; m.rrt_count = (uint32_t)l[19];
  m.cit_count = (uint32_t)l[20];
  const auto p = sl.at(MS_PERF);
  m.retrans_tx = p[0]; m.retrans_rx = p[1]; m.zero_win_tx = p[2]; m.zero_win_rx = p[3];
  m.retrans_syn = (uint32_t)p[4]; m.retrans_synack = (uint32_t)p[5];
  const auto a = sl.at(MS_ANOMALY);
  m.client_rst_flow = a[0]; m.server_rst_flow = a[1]; m.server_syn_miss = a[2]; m.client_ack_miss = a[3];
  m.client_half_close_flow = a[4]; m.server_half_close_flow = a[5]; m.client_source_port_reuse = a[6];
  m.client_establish_reset = a[7]; m.server_reset = a[8]; m.server_queue_lack = a[9];
  m.server_establish_reset = a[10]; m.tcp_timeout = a[11];
  m.client_establish_fail = m.client_ack_miss + m.client_source_port_reuse + m.client_establish_reset;  // :540-546
  m.server_establish_fail = m.server_syn_miss + m.server_reset + m.server_queue_lack + m.server_establish_reset;
  m.tcp_establish_fail = m.client_establish_fail + m.server_establish_fail;
  m.tcp_transfer_fail = m.client_rst_flow + m.server_rst_flow + m.server_queue_lack + m.tcp_timeout;
  m.tcp_rst_fail = m.client_establish_reset + m.server_establish_reset + m.server_reset + m.client_rst_flow +
                   m.server_rst_flow;
  m.l7_client_error = (uint32_t)a[12]; m.l7_server_error = (uint32_t)a[13]; m.l7_timeout = (uint32_t)a[14];
  m.ooo_tx = a[15]; m.ooo_rx = a[16];
  m.flow_load = sl[MS_FLOWLOAD];
}
template <class SL>
__device__ __forceinline__ void m_read_app(const SL& sl, MAppMeter& m) {
  m.request = (uint32_t)sl[MS_APPTRAFFIC]; m.response = (uint32_t)sl[MS_APPTRAFFIC + 1];
  m.direction_score = (uint8_t)sl[MS_APPTRAFFIC + 2];
  m.rrt_max = (uint32_t)sl[MS_APPLATENCY]; m.rrt_sum = sl[MS_APPLATENCY + 1]; m.rrt_count = (uint32_t)sl[MS_APPLATENCY + 2];
  m.client_error = (uint32_t)sl[MS_APPANOMALY]; m.server_error = (uint32_t)sl[MS_APPANOMALY + 1];
  m.timeout = (uint32_t)sl[MS_APPANOMALY + 2];
}
template <class SL>
__device__ __forceinline__ void m_read_usage(const SL& sl, MUsageMeter& m) {
  const auto u = sl.at(MS_USAGE);
  m.packet_tx = u[0]; m.packet_rx = u[1]; m.byte_tx = u[2]; m.byte_rx = u[3]; m.l3_byte_tx = u[4];
  m.l3_byte_rx = u[5]; m.l4_byte_tx = u[6]; m.l4_byte_rx = u[7];
}

// ---- getPlatformInfos / DocumentExpand ---------------------------------------------------------
__device__ __forceinline__ void m_addr(const MTag& t, int s, IpAddr& a) {
  a.ip4 = t.ip4[s];
  for (int i = 0; i < 16; i++) a.ip6[i] = t.ip6[s][i];
  a.ip6len = 16;
}

// handle_document.go:44-94 (side 0) / :96-145 (side 1)
__device__ __noinline__ bool m_side_info(const DevDict& d, MTag& t, int s, const IpAddr& a, InfoVal& info) {
  bool has = false;
  const int o = 1 - s;
  const bool v6 = t.is_ipv4 == 0;
  if (is_loopback(v6, a) && t.pod_id[s] == 0 && t.pod_id[o] != 0) {
    t.tag_source[s] |= TS_PEER;
    t.pod_id[s] = t.pod_id[o];
  }
  if (t.gpid[s] != 0 && t.pod_id[s] == 0) {
    uint32_t gi;
    if (t64_find(d.t64, d.t64_mask, TID_GPROC, t.gpid[s], gi)) {
      const uint64_t vp = d.gp_vals[gi];
      const uint32_t p = (uint32_t)vp;
      if (p != 0 && (uint16_t)(vp >> 32) == t.vtap_id) {
        t.pod_id[s] = p;
        t.tag_source[s] |= TS_GPID;
      }
    }
  }
  if (t.pod_id[s] != 0) {
    has = query_pod(d, t.pod_id[s], info);
    t.tag_source[s] |= TS_PODID;
  }
  if (!has) {
    if (t.mac[s] != 0) {
      t.tag_source[s] |= TS_MAC;
      has = query_mac(d, t.mac[s] | (uint64_t)(int64_t)t.l3_epc[s] << 48, info);
    }
    if (!has) {
      bool by_agent = false;
      if (t.tap_side == 4 || t.tap_side == 41 || t.tap_side == 42) by_agent = is_local_ip(v6, a);
      if (by_agent) {
        t.tag_source[s] |= TS_AGENT;
        uint32_t vi;
        if (t64_find(d.t64, d.t64_mask, TID_VTAP, t.vtap_id, vi)) {
          const DVtap& v = d.vtaps[vi];
          IpAddr va;
          va.ip4 = v.ip4;
          for (int i = 0; i < 16; i++) va.ip6[i] = v.ip6[i];
          va.ip6len = v.has_ip6 ? 16 : 0;
          has = reget_info_from_ip(d, !v.is_v4, va, v.epc, info);
        }
      } else {
        t.tag_source[s] |= TS_EPCIP;
        has = reget_info_from_ip(d, v6, a, t.l3_epc[s], info);
      }
    }
  }
  return has;
}

__device__ __forceinline__ void m_copy_info(MTag& t, int s, const InfoVal& i) {
  t.region_id[s] = (uint16_t)i.region_id;
  t.host_id[s] = (uint16_t)i.host_id;
  t.l3_device_id[s] = i.device_id;
  t.l3_device_type[s] = (uint8_t)i.device_type;
  t.subnet_id[s] = (uint16_t)i.subnet_id;
  t.pod_node_id[s] = i.pod_node_id;
  t.pod_ns_id[s] = (uint16_t)i.pod_ns_id;
  t.az_id[s] = (uint16_t)i.az_id;
  t.pod_group_id[s] = i.pod_group_id;
  t.pod_id[s] = i.pod_id;
  t.pod_cluster_id[s] = (uint16_t)i.pod_cluster_id;
}
__device__ __forceinline__ bool m_is_multicast(const MTag& t, int s) {  // net.IP.IsMulticast
  return t.is_ipv4 == 0 ? t.ip6[s][0] == 0xff : (t.ip4[s] >> 28) == 0xe;
}
__device__ __forceinline__ bool m_pod_service_ip(uint8_t dev_type, uint32_t pod, uint32_t pod_node) {
  return dev_type == 11 || pod != 0 || pod_node != 0;  // common.go:195-198
}

// false = the document is dropped (DropDocCount)
__device__ __noinline__ bool m_expand(const DevDict& d, MTag& t, uint8_t meter_id) {
  {  // DatabaseSuffixID tag.go:1283-1295: usage documents with only the acl suffix are not expanded
    int suffix = 0;
    if (t.code & MC_ACLGID) suffix |= 1;
    if (t.code & MC_EDGE_FIELDS) suffix |= 2;
    if (t.code & MC_SERVERPORT) suffix |= 4;
    if (meter_id == MT_METER_USAGE && suffix == 1) return true;
  }
  const uint16_t my_region = (uint16_t)d.region_id;
  if (t.code & MC_SERVERPORT) t.code |= MC_PORTADDCODE;
  const bool edge = (t.code & MC_EDGECODE) == MC_EDGECODE;
  const bool v6 = t.is_ipv4 == 0;
  IpAddr a0, a1;
  m_addr(t, 0, a0);
  m_addr(t, 1, a1);
  InfoVal info, info1;
  bool has = false, has1 = false;
  if (t.l3_epc[0] != EPC_FROM_INTERNET) has = m_side_info(d, t, 0, a0, info);
  if (edge && t.l3_epc[1] != EPC_FROM_INTERNET) has1 = m_side_info(d, t, 1, a1, info1);
  t.code |= edge ? MC_EDGEADDCODE : MC_MAINADDCODE;
  uint8_t pgt = 0, pgt1 = 0;
  const bool default_org = t.org_id <= 1;
  if (has1) {  // :176-213
    m_copy_info(t, 1, info1);
    pgt1 = info1.pod_group_type;
    if (m_pod_service_ip(t.l3_device_type[1], t.pod_id[1], t.pod_node_id[1]))
      t.service_id[1] = query_pod_service(d, t.pod_id[1], t.pod_node_id[1], t.pod_cluster_id[1], t.pod_group_id[1],
                                          t.l3_epc[1], v6, a1, t.protocol, t.server_port);
    if (!has && m_is_multicast(t, 0)) {
      t.region_id[0] = t.region_id[1];
      t.subnet_id[0] = t.subnet_id[1];
      t.az_id[0] = t.az_id[1];
      t.tag_source[0] |= TS_PEER;
    }
    if (default_org && my_region != 0 && t.region_id[1] != 0 && t.tap_side == 2 && t.region_id[1] != my_region)
      return false;
  }
  get_auto_instance(t.pod_id[1], t.gpid[1], t.pod_node_id[1], t.l3_device_id[1], t.subnet_id[1], t.l3_device_type[1],
                    t.l3_epc[1], t.auto_instance_id[1], t.auto_instance_type[1]);
  {
    uint32_t custom = query_custom_service(d, t.l3_epc[1], v6, a1, t.server_port, t.service_id[1], t.pod_group_id[1],
                                           t.l3_device_id[1], t.pod_id[1], t.pod_cluster_id[1], t.l3_device_type[1],
                                           t.protocol);
    get_auto_service(custom, t.service_id[1], t.pod_group_id[1], t.gpid[1], t.pod_cluster_id[1], t.l3_device_id[1],
                     t.subnet_id[1], t.l3_device_type[1], pgt1, t.l3_epc[1], t.auto_service_id[1],
                     t.auto_service_type[1]);
  }
  if (has) {  // :218-268
    m_copy_info(t, 0, info);
    pgt = info.pod_group_type;
    if (m_pod_service_ip(t.l3_device_type[0], t.pod_id[0], t.pod_node_id[0])) {
      if (t.server_port > 0 && !edge)
        t.service_id[0] = query_pod_service(d, t.pod_id[0], t.pod_node_id[0], t.pod_cluster_id[0], t.pod_group_id[0],
                                            t.l3_epc[0], v6, a0, t.protocol, t.server_port);
      else if (m_pod_service_ip(t.l3_device_type[0], t.pod_id[0], 0))
        t.service_id[0] = query_pod_service(d, t.pod_id[0], t.pod_node_id[0], t.pod_cluster_id[0], t.pod_group_id[0],
                                            t.l3_epc[0], v6, a0, t.protocol, 0);
    }
    if (!has1 && edge && m_is_multicast(t, 1)) {
      t.region_id[1] = t.region_id[0];
      t.subnet_id[1] = t.subnet_id[0];
      t.az_id[1] = t.az_id[0];
      t.tag_source[1] |= TS_PEER;
    }
    if (default_org && my_region != 0 && t.region_id[0] != 0) {
      if (edge) {
        if (t.tap_side == 1 && t.region_id[0] != my_region) return false;
      } else if (t.region_id[0] != my_region) {
        return false;
      }
    }
  }
  get_auto_instance(t.pod_id[0], t.gpid[0], t.pod_node_id[0], t.l3_device_id[0], t.subnet_id[0], t.l3_device_type[0],
                    t.l3_epc[0], t.auto_instance_id[0], t.auto_instance_type[0]);
  {
    const uint16_t port = edge ? 0 : t.server_port;
    uint32_t custom = query_custom_service(d, t.l3_epc[0], v6, a0, port, t.service_id[0], t.pod_group_id[0],
                                           t.l3_device_id[0], t.pod_id[0], t.pod_cluster_id[0], t.l3_device_type[0],
                                           t.protocol);
    get_auto_service(custom, t.service_id[0], t.pod_group_id[0], t.gpid[0], t.pod_cluster_id[0], t.l3_device_id[0],
                     t.subnet_id[0], t.l3_device_type[0], pgt, t.l3_epc[0], t.auto_service_id[0],
                     t.auto_service_type[0]);
  }
  if (t.signal_source == 4) {  // SIGNAL_SOURCE_OTEL :280-303
    if (t.tap_side == 50 && t.server_port == 0) t.server_port = 65535;
    for (int s = 0; s < 2; s++) {
      if (t.l3_epc[s] == EPC_FROM_INTERNET) t.l3_epc[s] = 0;
      if (t.auto_service_type[s] == 0) t.auto_service_type[s] = 255;
      if (t.auto_instance_type[s] == 0) t.auto_instance_type[s] = 255;
    }
  }
  return true;
}

__device__ __forceinline__ int m_table_id(const MDev* dv, uint64_t code, bool is_second) {  // tag.go:920-931
  code &= ~(MC_MAC | MC_MACPATH);
  for (int i = 0; i < MT_TABLE_COUNT; i++)
    if (code == dv->table_code[i]) {
      const int id = is_second ? i + MT_NETWORK_1S : i;
      return id < MT_TABLE_COUNT ? id : -1;  // per-second traffic_policy has no table
    }
  return -1;
}

}  // namespace

// ---- the decode kernel ------------------------------------------------------------------------------
// One thread per record, persistent warps.  The record is decoded ONCE by the table-driven tokenizer
// (pb_tok.cuh) into shared-memory words; the tail — tag, SmartEncoding, table routing — runs with the warp
// converged and leaves, per record, the values of its table's tag columns and its raw meter fields in
// plane-major buffers ([plane][record]: every store is one coalesced line per warp).  k_m_emit, after the
// scans have fixed the row of every record, moves them into the columns.  (Round 1 decoded and expanded every
// record twice — once to size, once to write — out of a 1 KB per-thread slot array in local memory.)
constexpr int MD_THREADS = 64;  // MW_COUNT words per thread: 41.5 KB per CTA, 5 CTAs per SM

struct MSink {
  uint32_t* w;  // s_words + tid, stride MD_THREADS
  __device__ __forceinline__ void u32(uint32_t s, uint32_t v) { w[s * MD_THREADS] = v; }
  __device__ __forceinline__ void u64(uint32_t s, uint32_t lo, uint32_t hi) {
    w[s * MD_THREADS] = lo;
    w[(s + 1) * MD_THREADS] = hi;
  }
  __device__ __forceinline__ void add(uint32_t s, uint32_t v) { w[s * MD_THREADS] += v; }
  __device__ __forceinline__ uint32_t word(uint32_t s) const { return w[s * MD_THREADS]; }
  __device__ __forceinline__ void ref(uint32_t, uint32_t, uint32_t) {}
  __device__ __forceinline__ void elem(uint32_t, uint32_t, uint32_t, uint32_t) {}
  __device__ __forceinline__ void velem(uint32_t, uint32_t, uint32_t, uint32_t) {}
};
// the interpreter's slot array seen through the words: sl[MS_X] is the 64-bit slot value
struct MView {
  const uint32_t* w;
  struct At {
    const MView& v;
    int base;
    __device__ __forceinline__ uint64_t operator[](int i) const { return v[base + i]; }
  };
  __device__ __forceinline__ uint64_t operator[](int s) const {
    const int k = mw(s);
    const uint32_t lo = w[k * MD_THREADS];
    return mw64(s) ? (uint64_t)w[(k + 1) * MD_THREADS] << 32 | lo : (uint64_t)lo;
  }
  __device__ __forceinline__ At at(int s) const { return At{*this, s}; }
};

// the tag words of a record as k_m_decode left them: [word][record] planes
struct MViewG {
  const uint32_t* w;  // tagw + rec
  size_t cap;
  struct At {
    const MViewG& v;
    int base;
    __device__ __forceinline__ uint64_t operator[](int i) const { return v[base + i]; }
  };
  __device__ __forceinline__ uint64_t operator[](int s) const {
    const int k = mw(s) - MWB_TAG;
    const uint32_t lo = w[(size_t)k * cap];
    return mw64(s) ? (uint64_t)w[(size_t)(k + 1) * cap] << 32 | lo : (uint64_t)lo;
  }
  __device__ __forceinline__ At at(int s) const { return At{*this, s}; }
};
constexpr int M_TAG_WORDS = MWB_TRAFFIC - MWB_TAG;  // 38

__global__ void __launch_bounds__(MD_THREADS, 5) k_m_decode(const __grid_constant__ BatchParams p) {
  __shared__ uint32_t s_tbl[MM_COUNT * MM_MAXF];
  __shared__ uint32_t s_words[MW_COUNT * MD_THREADS];
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  for (uint32_t i = threadIdx.x; i < MM_COUNT * MM_MAXF; i += MD_THREADS) s_tbl[i] = dv->tok_tbl[i];
  __syncthreads();
  DynState* dyn = p.dyn;
  if (dyn->overflow) return;
  const uint32_t nrec = batch_nrec(p), nwarps = dyn->nwarps;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * (MD_THREADS / 32) + (threadIdx.x >> 5), total_warps = gridDim.x * (MD_THREADS / 32);
  uint32_t* my_words = s_words + threadIdx.x;
  const size_t cap = dv->rec_cap;
  TokWin win;
  win.init(p.in, p.in_bytes);
  for (uint32_t wtile = gwarp; wtile < nwarps; wtile += total_warps) {
    const uint32_t rec = wtile * 32 + lane;
    const bool in_range = rec < nrec;
#pragma unroll
    for (int i = 0; i < MW_COUNT; i++) my_words[i * MD_THREADS] = 0;
    MSink sink{my_words};
    const MView sl{my_words};
    bool ok = false;
    uint32_t frame = 0;
    if (in_range) {
      frame = p.rec_frame[rec];
      if (!(p.repair && rec >= p.frame_bad[frame])) {
        ok = tk_decode<MM_MAXF>(win, p.rec_off[rec], p.rec_len[rec], s_tbl, sink, 0u);
        if (ok) {  // unknown meter id: DecodePB fails, the frame is abandoned (codec.go:60-63)
          const uint8_t id = (uint8_t)sl_u32(sl, MS_METER_ID);
          ok = id == MT_METER_FLOW || id == MT_METER_USAGE || id == MT_METER_APP;
        }
        if (!ok && !p.repair) {
          atomicMin(&p.frame_bad[frame], rec);
          dyn->stats.need_repair = 1;
        }
      }
    }
    // ---- hand-over (the warp is converged: every store below is one line per plane) ----
    const uint32_t meter_id = ok ? sl_u32(sl, MS_METER_ID) & 0xffu : 0u;
    if (in_range) {
      uint32_t* T = dv->tagw + rec;
#pragma unroll
      for (int k = 0; k < M_TAG_WORDS; k++) T[(size_t)k * cap] = my_words[(MWB_TAG + k) * MD_THREADS];
      dv->meta[rec] = (ok ? 0x100u : 0u) | meter_id;  // k_m_tag adds the table
    }
    if (ok) {
      uint64_t* R = dv->rawv + rec;
      uint32_t i = 0;
#define X(name, op)                          \
  R[(size_t)i * cap] = (uint64_t)m.name;     \
  i++;
      if (meter_id == MT_METER_FLOW) {
        MFlowMeter m;
        m_read_flow(sl, m);
        M_FLOW_FIELDS(X)
      } else if (meter_id == MT_METER_APP) {
        MAppMeter m;
        m_read_app(sl, m);
        M_APP_FIELDS(X)
      } else {
        MUsageMeter m;
        m_read_usage(sl, m);
        M_USAGE_FIELDS(X)
      }
#undef X
    }
    __syncwarp();
  }
}

// ---- the tag kernel: Tag.ReadFromPB, DocumentExpand, table routing, sizes ---------------------------------
// Thread per record over the decoded tag words.  The dictionary probe chains of DocumentExpand want many warps in
// flight, the tokenizer wants shared memory: two kernels, each at the occupancy that suits it.
__global__ void __launch_bounds__(256) k_m_tag(const __grid_constant__ BatchParams p) {
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  DynState* dyn = p.dyn;
  MLayout* lay = dv->lay;
  if (dyn->overflow) return;
  const uint32_t nrec = batch_nrec(p), nwarps = dyn->nwarps;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), total_warps = gridDim.x * (blockDim.x >> 5);
  const size_t cap = dv->rec_cap;
  for (uint32_t wtile = gwarp; wtile < nwarps; wtile += total_warps) {
    const uint32_t rec = wtile * 32 + lane;
    const bool in_range = rec < nrec;
    const uint32_t meta0 = in_range ? dv->meta[rec] : 0u;
    const bool ok = (meta0 & 0x100u) != 0;
    const uint8_t meter_id = (uint8_t)(meta0 & 0xffu);
    const MViewG sl{dv->tagw + rec, cap};
    MTag t;
    int table = -1;
    if (ok) {
      m_read_tag(sl, p.in, t);
      const dfgpu_frame fr = p.frames[p.rec_frame[rec]];
      t.team_id = (uint16_t)fr.team_id;  // unmarshaller.go:242-243
      t.org_id = fr.org_id;
      if (m_expand(p.dict, t, meter_id)) {
        table = m_table_id(dv, t.code, (sl_u32(sl, MS_FLAGS) & 1) == 1);
        // a meter the table does not store cannot be packed by the reference either: dropped
        if (table >= 0 && dv->table_meter[table] != meter_id) table = -1;
      }
    }
    const bool valid = table >= 0;
    const int app = valid ? m_app_index(table) : -1;
    const int obs = valid ? m_obs_space(table) : -1;
    const uint32_t tside = obs >= 0 ? tapside_code(t.tap_side) : 0;
    const uint32_t same = __match_any_sync(0xffffffffu, valid ? table : 0xff);
    // ---- sizes: rows per table, string bytes per (application table, column) ----
    if (lane < M_SLOTS) p.warp_tot[(size_t)lane * p.warp_cap + wtile] = 0;
    __syncwarp();
    if (valid && (int)lane == __ffs(same) - 1) p.warp_tot[(size_t)(M_SLOT_ROWS + table) * p.warp_cap + wtile] = __popc(same);
    if (__any_sync(0xffffffffu, app >= 0)) {
#pragma unroll 1
      for (int a = 0; a < 4; a++)
#pragma unroll 1
        for (int w = 0; w < 3; w++) {
          const uint32_t tot = warp_sum(app == a ? t.str[w].n : 0u);
          if (lane == 0) p.warp_tot[(size_t)(M_SLOT_STR + a * 3 + w) * p.warp_cap + wtile] = tot;
        }
    }
    lc_note(&dyn->stats.lc_first[0][0], (uint32_t)(obs < 0 ? 0 : obs) * 256u + tside, rec, obs >= 0);
    const uint32_t n_ok = __popc(__ballot_sync(0xffffffffu, ok)), n_drop = __popc(__ballot_sync(0xffffffffu, ok && !valid));
    if (lane == 0) {
      if (n_ok) atomicAdd(&lay->count, n_ok);
      if (n_drop) atomicAdd(&lay->drop_count, n_drop);
    }
    // ---- hand-over to k_m_emit ----
    MCand cand;
    memset(&cand, 0, sizeof cand);
    if (valid) {
      const uint64_t tcode = dv->table_code[table];
      uint64_t* T = dv->tagv + rec;
      uint32_t k = 0;
#define PUT(v)                       \
  {                                  \
    T[(size_t)k * cap] = (uint64_t)(v); \
    k++;                             \
  }
#define F(bit, name, ch, ctype, expr) \
  if ((bit) == 0 || (tcode & (bit))) PUT((ctype)(expr))
      // an address column of a v4 row (or of side 1 of a single-ended tag) is all zeros: length 0
#define V6(bit, name, side) \
  if (tcode & (bit)) PUT((t.is_ipv4 == 0 && ((side) == 0 || (t.code & MC_IPPATH))) ? sl[(side) ? MS_IP1 : MS_IP] : 0ull)
#define OBS(bit, name) \
  if (tcode & (bit)) PUT(tside)
#define STRCOL(bit, name, which) \
  if (tcode & (bit)) PUT((uint64_t)t.str[which].n << 32 | (uint32_t)(t.str[which].p - p.in))
      MT_TAG_COLUMNS(F, V6, OBS, STRCOL, STRCOL)
#undef F
#undef V6
#undef OBS
#undef STRCOL
#undef PUT
      // appServiceTagWrite unmarshaller.go:309-327: only the 1m application tables
      if ((table == MT_APPLICATION_1M || table == MT_APPLICATION_MAP_1M) && (t.str[0].n | t.str[1].n)) {
        cand.svc_off = (uint32_t)(t.str[0].p - p.in);
        cand.svc_len = t.str[0].n;
        cand.inst_off = (uint32_t)(t.str[1].p - p.in);
        cand.inst_len = t.str[1].n;
        cand.time = t.time;
        cand.team = t.team_id;
        cand.table = table == MT_APPLICATION_1M ? 0 : 1;
        cand.state = 1;
      }
    }
    if (in_range) {
      dv->cand[rec] = cand;
      // table (0xff: no row) | meter id << 8 | tap-side code << 16
      dv->meta[rec] = (valid ? (uint32_t)table : 0xffu) | (uint32_t)meter_id << 8 | tside << 16;
      dv->strv[rec] = valid && app >= 0 ? make_uint4(t.str[0].n, t.str[1].n, t.str[2].n, 0u) : make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
  }
}

// ---- the emit kernel: planes -> columns, after the scans ------------------------------------------------
__global__ void __launch_bounds__(256) k_m_emit(const __grid_constant__ BatchParams p) {
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  DynState* dyn = p.dyn;
  const MLayout* lay = dv->lay;
  if ((dyn->stats.need_repair && !p.repair) || dyn->overflow || lay->overflow) return;
  const uint32_t nrec = batch_nrec(p), nwarps = dyn->nwarps;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), total_warps = gridDim.x * (blockDim.x >> 5);
  const size_t cap = dv->rec_cap;
  uint8_t* A = p.arena;
  for (uint32_t wtile = gwarp; wtile < nwarps; wtile += total_warps) {
    const uint32_t rec = wtile * 32 + lane;
    const bool in_range = rec < nrec;
    const uint32_t meta = in_range ? dv->meta[rec] : 0xffu;
    const int table = (meta & 0xffu) == 0xffu ? -1 : (int)(meta & 0xffu);
    const bool valid = table >= 0;
    const uint32_t meter_id = (meta >> 8) & 0xffu, tside = (meta >> 16) & 0xffu;
    const int app = valid ? m_app_index(table) : -1;
    const int obs = valid ? m_obs_space(table) : -1;
    const uint32_t same = __match_any_sync(0xffffffffu, valid ? table : 0xff);
    uint32_t e_str[3] = {0, 0, 0};
    if (__any_sync(0xffffffffu, app >= 0)) {
      const uint4 sn = in_range ? dv->strv[rec] : make_uint4(0u, 0u, 0u, 0u);
      const uint32_t n3[3] = {sn.x, sn.y, sn.z};
#pragma unroll 1
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int w = 0; w < 3; w++) {
          const uint32_t e = warp_excl_scan(app == a ? n3[w] : 0u);
          if (app == a) e_str[w] = e;
        }
    }
    uint32_t row = 0, d0 = 0, ntag = 0;
    if (valid) {
      row = p.warp_base[(size_t)(M_SLOT_ROWS + table) * p.warp_cap + wtile] + __popc(same & ((1u << lane) - 1));
      d0 = dv->desc_base[table];
      ntag = dv->ntag[table];
      dv->row_rec[lay->row_base[table] + row] = rec;
    }
    // tag columns: plane k of the record -> descriptor d0 + k of its table
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, ntag);
    for (uint32_t k = 0; k < kmax; k++) {
      if (k >= ntag) continue;
      const uint64_t v = dv->tagv[(size_t)k * cap + rec];
      const uint32_t d = d0 + k;
      const MColDesc c = dv->descs[d];
      uint8_t* col = A + lay->col_off[d];
      if (c.kind == MK_FIX) {
        if (c.elem == 1) col[row] = (uint8_t)v;
        else if (c.elem == 2) reinterpret_cast<uint16_t*>(col)[row] = (uint16_t)v;
        else if (c.elem == 4) reinterpret_cast<uint32_t*>(col)[row] = (uint32_t)v;
        else reinterpret_cast<uint64_t*>(col)[row] = v;
      } else if (c.kind == MK_IP6) {  // p.Field.Ip[:16], zero extended
        const uint32_t off = (uint32_t)v, len = (uint32_t)(v >> 32);
        uint64_t lo = 0, hi = 0;
        const uint32_t n = len > 16 ? 16 : len;
        for (uint32_t i = 0; i < n; i++) {
          const uint64_t b = (uint64_t)__ldg(p.in + off + i) << ((i & 7) * 8);
          if (i < 8) lo |= b; else hi |= b;
        }
        reinterpret_cast<ulonglong2*>(col)[row] = make_ulonglong2(lo, hi);
      } else if (c.kind == MK_OBS) {
        col[row] = dyn->lc_rank[obs < 0 ? 0 : obs][tside];
      } else {  // MK_LCS / MK_STR
        const uint32_t which = c.which, n = (uint32_t)(v >> 32);
        const uint64_t off = (uint64_t)p.warp_base[(size_t)(M_SLOT_STR + app * 3 + which) * p.warp_cap + wtile] + e_str[which];
        reinterpret_cast<dfgpu_pos*>(A + lay->aux_off[d])[row] = dfgpu_pos{(int64_t)off, (int64_t)(off + n)};
        uint8_t* dst = col + off;
        const uint8_t* src = p.in + (uint32_t)v;
        for (uint32_t i = 0; i < n; i++) dst[i] = __ldg(src + i);
      }
    }
    if (valid) {
      uint32_t d = d0 + ntag;
      const uint64_t* R = dv->rawv + rec;
      uint64_t* RW = (dv->raw_tables & (1u << table))
                         ? reinterpret_cast<uint64_t*>(A + lay->raw_off[table]) + (size_t)row * dv->raw_nf[table]
                         : nullptr;
      uint32_t i = 0;
#define X(name, op)                                                 \
  {                                                                 \
    const uint64_t r_ = R[(size_t)i * cap];                         \
    m.name = (decltype(m.name))r_;                                  \
    if (RW) RW[i] = r_;                                             \
    i++;                                                            \
  }
#define M(name, ch, ctype, expr)                                    \
  reinterpret_cast<ctype*>(A + lay->col_off[d])[row] = (ctype)(expr); \
  d++;
      if (meter_id == MT_METER_FLOW) {
        MFlowMeter m;
        M_FLOW_FIELDS(X)
        MT_FLOW_METER_COLUMNS(M)
      } else if (meter_id == MT_METER_APP) {
        MAppMeter m;
        M_APP_FIELDS(X)
        MT_APP_METER_COLUMNS(M)
      } else {
        MUsageMeter m;
        M_USAGE_FIELDS(X)
        MT_USAGE_METER_COLUMNS(M)
      }
#undef M
#undef X
    }
    __syncwarp();
  }
}

// ---- layout: one thread walks the ~700 column descriptors ----------------------------------------
__global__ void k_m_layout(const BatchParams p) {
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  MLayout* lay = dv->lay;
  const uint64_t* tot = p.dyn->stats.var_total;
  uint64_t cur = 0;
  for (int a = 0; a < 12; a++) lay->str_bytes[a] = tot[M_SLOT_STR + a];
  for (int t = 0; t < MT_TABLE_COUNT; t++) {
    const uint64_t nrows = tot[M_SLOT_ROWS + t];
    lay->nrows[t] = (uint32_t)nrows;
    lay->row_base[t] = t ? lay->row_base[t - 1] + lay->nrows[t - 1] : 0;
    const int app = m_app_index(t);
    for (uint32_t d = dv->desc_base[t]; d < dv->desc_base[t + 1]; d++) {
      const MColDesc c = dv->descs[d];
      if (c.kind == MK_LCS || c.kind == MK_STR) {
        lay->aux_off[d] = cur;
        cur += (nrows * 16 + 255) & ~255ull;
        lay->col_off[d] = cur;
        cur += (tot[M_SLOT_STR + app * 3 + c.which] + 255) & ~255ull;
      } else {
        lay->col_off[d] = cur;
        cur += (nrows * c.elem + 255) & ~255ull;
      }
    }
  }
  lay->end = cur;
  // raw meters of the tables the rollup is enabled for: device-only, behind the 9 tables (the app_service
  // tag table follows them; fetch copies the two host-bound ranges separately)
  for (int t = 0; t < MT_TABLE_COUNT; t++) {
    lay->raw_off[t] = cur;
    if (dv->raw_tables & (1u << t)) cur += ((uint64_t)lay->nrows[t] * dv->raw_nf[t] * 8 + 255) & ~255ull;
  }
  lay->raw_end = cur;
  if (cur > p.arena_bytes) {
    lay->overflow = 1;
    p.dyn->overflow = 1;  // dfgpu_fetch grows the arena and re-runs
    p.dyn->dev_end = cur + (cur >> 3);
  }
}

// ---- app_service tags -----------------------------------------------------------------------------
__global__ void k_m_as_prep(const BatchParams p) {
  if (p.dyn->stats.need_repair && !p.repair) return;  // this pass is re-run with the frames' cut points: leave the set alone
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  const uint32_t rec = blockIdx.x * blockDim.x + threadIdx.x;
  if (rec >= batch_nrec(p)) return;
  MCand& c = dv->cand[rec];
  if (c.state != 1) return;
  const StrHash a = str_hash(p.in + c.svc_off, c.svc_len), b = str_hash(p.in + c.inst_off, c.inst_len);
  const uint64_t y = (uint64_t)dv->org_id << 24 | (uint64_t)c.team << 8 | c.table;
  const U128 fp = fp_finish(a.a ^ mix64(b.a + y * 0x9E3779B97F4A7C15ull), a.b + mix64(b.b ^ (y * 0xC2B2AE3D27D4EB4Full)));
  FtEntry* e = ts_upsert(dv->as_set, dv->as_mask, fp, dv->as_state);
  atomicMax(&dv->as_state->now, c.time);
  if (!e) {  // saturated set: emit without recording (idempotent duplicate), never fail the batch
    c.state = 2;
    atomicAdd(&dv->as_state->unrecorded, 1u);
    return;
  }
  c.slot = (uint32_t)(e - dv->as_set);
}
// one lru round: every pending record that finds its key new or expired bids for the
// emission; the earliest record wins (k_m_as_decide) and stores its time, the others are
// looked at again in the next round against that time — the sequential outcome, because a
// key's stored time only grows (a record emits only if old + timeout < its time).
__global__ void k_m_as_mark(const BatchParams p, uint32_t epoch) {
  if (p.dyn->stats.need_repair && !p.repair) return;  // this pass is re-run with the frames' cut points: leave the set alone
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  const uint32_t rec = blockIdx.x * blockDim.x + threadIdx.x;
  if (rec >= batch_nrec(p)) return;
  MCand& c = dv->cand[rec];
  if (c.state != 1) return;
  FtEntry* e = dv->as_set + c.slot;
  if (ts_candidate(e, c.time, dv->as_timeout)) atomicMax(&e->first_pos, ts_pos_key(epoch, rec));
  else c.state = 3;
}
__global__ void k_m_as_decide(const BatchParams p, uint32_t epoch, int last) {
  if (p.dyn->stats.need_repair && !p.repair) return;  // this pass is re-run with the frames' cut points: leave the set alone
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  const uint32_t rec = blockIdx.x * blockDim.x + threadIdx.x;
  if (rec >= batch_nrec(p)) return;
  MCand& c = dv->cand[rec];
  if (c.state != 1) return;
  FtEntry* e = dv->as_set + c.slot;
  if (e->first_pos == ts_pos_key(epoch, rec)) {
    c.state = 2;
    e->time = c.time;  // lru Add(key, time)
    e->state = 1;
  } else if (last) {
    dv->lay->as_pending = 1;
  }
}
__global__ void k_m_as_count(const BatchParams p) {
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  const uint32_t rec = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = rec >> 5, lane = threadIdx.x & 31;
  MCand c;
  c.state = 0;
  if (rec < batch_nrec(p)) c = dv->cand[rec];
  const bool emit = c.state == 2;
  const uint32_t rows = __popc(__ballot_sync(0xffffffffu, emit));
  const uint32_t b0 = warp_sum(emit ? c.svc_len : 0u), b1 = warp_sum(emit ? c.inst_len : 0u);
  if (lane == 0 && warp < p.dyn->nwarps) {
    p.warp_tot[0 * (size_t)p.warp_cap + warp] = rows;
    p.warp_tot[1 * (size_t)p.warp_cap + warp] = b0;
    p.warp_tot[2 * (size_t)p.warp_cap + warp] = b1;
  }
  lc_note(dv->lay->as_lc_first, emit ? c.table : 0u, rec, emit);
}
__global__ void k_m_as_layout(const BatchParams p) {
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  MLayout* lay = dv->lay;
  const uint64_t rows = p.dyn->stats.ft_tot[0];
  const uint64_t bytes = p.dyn->stats.ft_tot[1], bytes1 = p.dyn->stats.ft_tot[2];
  uint64_t cur = (lay->raw_end + 255) & ~255ull;
  auto take = [&](int i, uint64_t n) {
    lay->as_off[i] = cur;
    cur += (n + 255) & ~255ull;
  };
  take(0, rows * 4);   // time
  take(1, rows);       // table key (rank)
  take(2, rows * 16);  // app_service pos
  take(3, rows * 16);  // app_instance pos
  take(4, rows * 2);   // team_id
  take(5, bytes);      // app_service bytes
  take(6, bytes1);     // app_instance bytes
  lay->as_rows = rows;
  lay->as_bytes[0] = bytes;
  lay->as_bytes[1] = bytes1;
  lay->as_end = cur;
  lay->as_count = dv->as_state->count;
  lay->as_unrecorded = dv->as_state->unrecorded;
  lay->as_evicted = dv->as_state->evicted;
  if (cur > p.arena_bytes) {
    lay->overflow = 1;
    p.dyn->overflow = 1;
    p.dyn->dev_end = cur + (cur >> 3);
  }
}
__global__ void k_m_as_write(const BatchParams p) {
  const MDev* dv = reinterpret_cast<const MDev*>(p.ext);
  const MLayout* lay = dv->lay;
  if (lay->overflow) return;
  const uint32_t rec = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = rec >> 5, lane = threadIdx.x & 31;
  MCand c;
  c.state = 0;
  c.svc_len = c.inst_len = 0;
  if (rec < batch_nrec(p)) c = dv->cand[rec];
  const bool emit = c.state == 2;
  const uint32_t m = __ballot_sync(0xffffffffu, emit);
  const uint32_t eb = warp_excl_scan(emit ? c.svc_len : 0u), eb1 = warp_excl_scan(emit ? c.inst_len : 0u);
  if (!emit) return;
  const uint32_t wb = warp < p.dyn->nwarps ? warp : 0;
  const uint32_t row = p.warp_base[0 * (size_t)p.warp_cap + wb] + __popc(m & ((1u << lane) - 1));
  const uint64_t off = (uint64_t)p.warp_base[1 * (size_t)p.warp_cap + wb] + eb;
  const uint64_t off1 = (uint64_t)p.warp_base[2 * (size_t)p.warp_cap + wb] + eb1;
  uint8_t* A = p.arena;
  reinterpret_cast<uint32_t*>(A + lay->as_off[0])[row] = c.time;
  // dictionary rank by first appearance: the other table ranks first only if it appeared earlier
  const uint32_t other = lay->as_lc_first[c.table ^ 1], mine = lay->as_lc_first[c.table];
  (A + lay->as_off[1])[row] = other < mine ? 1 : 0;
  reinterpret_cast<dfgpu_pos*>(A + lay->as_off[2])[row] = dfgpu_pos{(int64_t)off, (int64_t)(off + c.svc_len)};
  reinterpret_cast<dfgpu_pos*>(A + lay->as_off[3])[row] = dfgpu_pos{(int64_t)off1, (int64_t)(off1 + c.inst_len)};
  reinterpret_cast<uint16_t*>(A + lay->as_off[4])[row] = c.team;
  uint8_t* dst = A + lay->as_off[5] + off;
  for (uint32_t i = 0; i < c.svc_len; i++) dst[i] = __ldg(p.in + c.svc_off + i);
  dst = A + lay->as_off[6] + off1;
  for (uint32_t i = 0; i < c.inst_len; i++) dst[i] = __ldg(p.in + c.inst_off + i);
}
__global__ void k_ts_clear_pos(FtEntry* t, uint32_t mask) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= mask; i += gridDim.x * blockDim.x) t[i].first_pos = 0;
}

// ---- host side -------------------------------------------------------------------------------------
static int m_ensure_state(dfgpu_ctx* ctx) {
  if (ctx->mt) return 0;
  MHost* h = new MHost();
  MDev& d = h->dev;
  auto tbl = metrics_tok_table();
  memcpy(d.tok_tbl, tbl.data(), sizeof d.tok_tbl);
  const uint64_t codes[MT_TABLE_COUNT] = {MC_NETWORK, MC_NETWORK_MAP, MC_APPLICATION, MC_APPLICATION_MAP,
                                          MC_TRAFFIC_POLICY, MC_NETWORK, MC_NETWORK_MAP, MC_APPLICATION,
                                          MC_APPLICATION_MAP};
  uint32_t nd = 0;
  for (int t = 0; t < MT_TABLE_COUNT; t++) {
    d.table_code[t] = codes[t];
    d.table_meter[t] = codes[t] == MC_TRAFFIC_POLICY ? MT_METER_USAGE : (codes[t] & MC_L7PROTOCOL) ? MT_METER_APP : MT_METER_FLOW;
    d.desc_base[t] = nd;
    const uint64_t code = codes[t];
    auto add = [&](const char* name, int ch, uint8_t kind, uint8_t elem, uint8_t which) {
      d.descs[nd++] = MColDesc{kind, elem, which, 0};
      h->cols[t].push_back(ColSpec{name, ch, kind == MK_FIX || kind == MK_IP6 || kind == MK_OBS ? (int)elem : 0, -1});
    };
#define F(bit, name, ch, ctype, expr) if ((bit) == 0 || (code & (bit))) add(name, ch, MK_FIX, sizeof(ctype), 0);
#define V6(bit, name, side) if (code & (bit)) add(name, DFGPU_CH_IPV6, MK_IP6, 16, side);
#define OBS(bit, name) if (code & (bit)) add(name, DFGPU_CH_LC_STR, MK_OBS, 1, 0);
#define LCS(bit, name, which) if (code & (bit)) add(name, DFGPU_CH_LC_STR_ROWS, MK_LCS, 0, which);
#define STR(bit, name, which) if (code & (bit)) add(name, DFGPU_CH_STR, MK_STR, 0, which);
    MT_TAG_COLUMNS(F, V6, OBS, LCS, STR)
#undef F
#undef V6
#undef OBS
#undef LCS
#undef STR
    d.ntag[t] = nd - d.desc_base[t];
    h->max_ntag = std::max(h->max_ntag, d.ntag[t]);
#define M(name, ch, ctype, expr) add(name, ch, MK_FIX, sizeof(ctype), 0);
    if (d.table_meter[t] == MT_METER_FLOW) { MT_FLOW_METER_COLUMNS(M) }
    else if (d.table_meter[t] == MT_METER_APP) { MT_APP_METER_COLUMNS(M) }
    else { MT_USAGE_METER_COLUMNS(M) }
#undef M
  }
  d.desc_base[MT_TABLE_COUNT] = nd;
  d.raw_tables = ctx->cfg.rollup_tables & 0x1ffu;
  for (int t = 0; t < MT_TABLE_COUNT; t++) {
    uint32_t nf = 0;
#define X(name, op) nf++;
    if (d.table_meter[t] == MT_METER_FLOW) { M_FLOW_FIELDS(X) }
    else if (d.table_meter[t] == MT_METER_APP) { M_APP_FIELDS(X) }
    else { M_USAGE_FIELDS(X) }
#undef X
    d.raw_nf[t] = nf;
  }
  if (tagset_create(h->as_set, (uint64_t)ctx->cfg.flow_tag_cache_max_size * 2, ctx->cfg.flow_tag_cache_flush_timeout))
    return fail(ctx, DFGPU_ENOMEM, "app_service tag set allocation failed");
  d.as_timeout = ctx->cfg.flow_tag_cache_flush_timeout;
  CU(cudaMalloc(&h->d_dev, sizeof(MDev)));
  CU(cudaMalloc(&h->d_lay, sizeof(MLayout)));
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    h->grid_decode = (unsigned)sms * 5;
    h->grid_emit = (unsigned)sms * 6;
    cudaFuncSetAttribute(k_m_decode, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  CU(cudaHostAlloc(&h->h_lay, sizeof(MLayout), cudaHostAllocDefault));
  ctx->mt = h;
  return 0;
}

void metrics_free_state(dfgpu_ctx* ctx) {
  MHost* h = reinterpret_cast<MHost*>(ctx->mt);
  if (!h) return;
  tagset_destroy(h->as_set);
  cudaFree(h->d_dev);
  cudaFree(h->d_lay);
  cudaFree(h->d_cand);
  cudaFree(h->d_row_rec);
  cudaFree(h->d_planes);
  cudaFreeHost(h->h_lay);
  delete h;
  ctx->mt = nullptr;
}

// arena bytes that are always enough: every row of the widest table + the strings + the tag rows
uint64_t metrics_arena_need(uint32_t nrec, uint64_t nbytes) { return (uint64_t)nrec * 1100 + 3 * nbytes + (2u << 20); }

static uint32_t m_next_epoch(MHost* h, cudaStream_t s) {
  h->epoch++;
  if ((h->epoch & 0xffff) == 0) {
    k_ts_clear_pos<<<1024, 256, 0, s>>>(h->as_set.cur, h->as_set.mask);
    h->epoch = 1;
  }
  return h->epoch;
}

static int m_as_rounds(dfgpu_ctx* ctx, cudaStream_t s, int rounds) {
  MHost* h = reinterpret_cast<MHost*>(ctx->mt);
  BatchParams& bp = ctx->bp;
  dim3 grid(rec_grid(bp)), block(256);
  for (int r = 0; r < rounds; r++) {
    const uint32_t ep = m_next_epoch(h, s);
    k_m_as_mark<<<grid, block, 0, s>>>(bp, ep);
    k_m_as_decide<<<grid, block, 0, s>>>(bp, ep, r == rounds - 1);
    ctx->launches += 2;
  }
  return 0;
}
static int m_as_emit(dfgpu_ctx* ctx, cudaStream_t s) {
  MHost* h = reinterpret_cast<MHost*>(ctx->mt);
  BatchParams& bp = ctx->bp;
  CU(cudaMemsetAsync(&h->d_lay->as_lc_first, 0xff, sizeof(h->d_lay->as_lc_first), s));
  k_m_as_count<<<rec_grid(bp), 256, 0, s>>>(bp);
  launch_slot_scan_to(bp, 3, ctx->d_dyn->stats.ft_tot, s);
  k_m_as_layout<<<1, 1, 0, s>>>(bp);
  k_m_as_write<<<rec_grid(bp), 256, 0, s>>>(bp);
  ctx->launches += 5;
  return 0;
}

int metrics_ensure_state(dfgpu_ctx* ctx) { return m_ensure_state(ctx); }

// all kernels of one metrics batch (after framing); repair = 1 when re-running with known bad records
int metrics_passes(dfgpu_ctx* ctx, int repair) {
  int rc = m_ensure_state(ctx);
  if (rc) return rc;
  MHost* h = reinterpret_cast<MHost*>(ctx->mt);
  BatchParams& bp = ctx->bp;
  cudaStream_t s = ctx->stream;
  bp.repair = repair;
  bp.nslots = M_SLOTS;
  bp.nstreams = 0;
  bp.n_lc = 4;
  bp.ncols = 0;  // the 9 tables are laid out by k_m_layout, not k_prepare
  if (bp.rec_cap > h->cand_cap) {  // first metrics batch of this context (or the record index grew)
    if (h->d_cand) CU(cudaFree(h->d_cand));
    h->d_cand = nullptr;
    size_t cap = (size_t)bp.rec_cap;
    CU(cudaMalloc(&h->d_cand, cap * sizeof(MCand)));
    if (h->d_row_rec) CU(cudaFree(h->d_row_rec));
    h->d_row_rec = nullptr;
    CU(cudaMalloc(&h->d_row_rec, cap * sizeof(uint32_t)));
    // hand-over of the decode kernel: tag column values, raw meter fields, string lengths, meta word
    if (h->d_planes) CU(cudaFree(h->d_planes));
    h->d_planes = nullptr;
    const size_t per_rec = ((size_t)h->max_ntag + M_MAX_FIELDS) * 8 + 16 + 4 + (size_t)M_TAG_WORDS * 4;
    CU(cudaMalloc(&h->d_planes, per_rec * cap + 1024));
    uint8_t* q = h->d_planes;
    h->dev.strv = reinterpret_cast<uint4*>(q); q += 16 * cap;
    h->dev.tagv = reinterpret_cast<uint64_t*>(q); q += (size_t)h->max_ntag * 8 * cap;
    h->dev.rawv = reinterpret_cast<uint64_t*>(q); q += (size_t)M_MAX_FIELDS * 8 * cap;
    h->dev.meta = reinterpret_cast<uint32_t*>(q); q += 4 * cap;
    h->dev.tagw = reinterpret_cast<uint32_t*>(q);
    h->dev.rec_cap = (uint32_t)cap;
    h->cand_cap = cap;
  }
  tagset_maybe_rebuild(h->as_set, s, &ctx->launches);
  h->dev.as_set = h->as_set.cur;
  h->dev.as_mask = h->as_set.mask;
  h->dev.as_state = h->as_set.d_state;
  h->dev.lay = h->d_lay;
  h->dev.cand = h->d_cand;
  h->dev.row_rec = h->d_row_rec;
  h->dev.org_id = ctx->cur_org;
  CU(cudaMemcpyAsync(h->d_dev, &h->dev, sizeof(MDev), cudaMemcpyHostToDevice, s));
  bp.ext = h->d_dev;
  CU(cudaMemsetAsync(h->d_lay, 0, sizeof(MLayout), s));
  CU(cudaMemsetAsync(&ctx->d_dyn->stats.lc_first, 0xff, sizeof(ctx->d_dyn->stats.lc_first), s));
  CU(cudaMemsetAsync(&ctx->d_dyn->stats.var_total, 0, sizeof(ctx->d_dyn->stats.var_total), s));
  dim3 grid(rec_grid(bp)), block(256);
  ctx->mark(2);
  k_m_decode<<<h->grid_decode, MD_THREADS, 0, s>>>(bp);
  k_m_tag<<<h->grid_emit, 256, 0, s>>>(bp);
  ctx->launches++;
  ctx->mark(3);
  launch_slot_scan(bp, s);
  launch_finish(bp, s);  // LowCardinality ranks of the 4 observation_point spaces
  k_m_layout<<<1, 1, 0, s>>>(bp);
  ctx->mark(4);
  k_m_emit<<<h->grid_emit, 256, 0, s>>>(bp);
  ctx->mark(5);
  ctx->mark(6);
  k_m_as_prep<<<grid, block, 0, s>>>(bp);
  ctx->launches += 7;
  rc = m_as_rounds(ctx, s, AS_ROUNDS);
  if (rc) return rc;
  rc = m_as_emit(ctx, s);
  if (rc) return rc;
  CU(cudaGetLastError());
  return 0;
}

// fetch: D2H of the used arena + column views of the 9 tables and the app_service table
int metrics_fetch(dfgpu_ctx* ctx, dfgpu_batch* b, uint8_t* (*acquire)(dfgpu_ctx*, size_t)) {
  MHost* h = reinterpret_cast<MHost*>(ctx->mt);
  BatchParams& bp = ctx->bp;
  cudaStream_t s = ctx->stream;
  // sizes / overflow / repair were settled by dfgpu_fetch (DynState); here: the layout and the
  // app_service rounds
  for (int attempt = 0;; attempt++) {
    CU(cudaMemcpyAsync(h->h_lay, h->d_lay, sizeof(MLayout), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (h->h_lay->overflow) return fail(ctx, DFGPU_ENOMEM, "metrics arena overflow after the app_service rounds");
    if (h->h_lay->as_pending) {  // a key re-emits more than AS_ROUNDS times inside one batch: keep going
      if (attempt >= 64) return fail(ctx, DFGPU_ENOMEM, "app_service tag rounds did not converge");
      CU(cudaMemsetAsync(&h->d_lay->as_pending, 0, 4, s));
      int rc = m_as_rounds(ctx, s, AS_ROUNDS);
      if (rc) return rc;
      rc = m_as_emit(ctx, s);
      if (rc) return rc;
      continue;
    }
    break;
  }
  h->as_set.count_seen = h->h_lay->as_count;
  b->counters.flow_tag_unrecorded = h->h_lay->as_unrecorded;
  b->counters.flow_tag_evicted = h->h_lay->as_evicted;
  const MLayout& L = *h->h_lay;
  const uint64_t total = std::max<uint64_t>(L.as_end, 256);
  b->host = acquire(ctx, (size_t)total + 64);
  if (!b->host) return fail(ctx, DFGPU_ENOMEM, "pinned host allocation of %llu bytes failed", (unsigned long long)total);
  uint8_t* hb = b->host + 64;
  if (L.raw_end > L.end) {  // skip the device-only raw meters between the tables and the app_service table
    CU(cudaMemcpyAsync(hb, bp.arena, (size_t)L.end, cudaMemcpyDeviceToHost, s));
    if (total > L.raw_end) CU(cudaMemcpyAsync(hb + L.raw_end, bp.arena + L.raw_end, (size_t)(total - L.raw_end), cudaMemcpyDeviceToHost, s));
  } else {
    CU(cudaMemcpyAsync(hb, bp.arena, (size_t)total, cudaMemcpyDeviceToHost, s));
  }
  CU(cudaStreamSynchronize(s));
  static const std::vector<std::string> tap(kTapSideNames, kTapSideNames + 18);
  b->owned.reserve(16);
  b->owned_pos.reserve(16);
  auto lc_dict = [&](const uint32_t* first, uint32_t ncodes, const std::vector<std::string>& names, dfgpu_col& col) {
    std::vector<std::pair<uint32_t, uint32_t>> present;
    for (uint32_t c = 0; c < ncodes; c++)
      if (first[c] != NO_BAD) present.push_back({first[c], c});
    std::sort(present.begin(), present.end());
    b->owned.emplace_back();
    b->owned_pos.emplace_back();
    auto& buf = b->owned.back();
    auto& pos = b->owned_pos.back();
    for (auto& pr : present) {
      const std::string& str = names[pr.second];
      pos.push_back(dfgpu_pos{(int64_t)buf.size(), (int64_t)(buf.size() + str.size())});
      buf.insert(buf.end(), str.begin(), str.end());
    }
    col.dict_data = buf.data();
    col.dict_pos = pos.data();
    col.dict_n = pos.size();
  };
  for (int t = 0; t < MT_TABLE_COUNT; t++) {
    b->tables.emplace_back();
    dfgpu_batch::Table& tb = b->tables.back();
    tb.id = DFGPU_TABLE_METRICS_BASE + t;
    tb.nrows = L.nrows[t];
    const auto& cs = h->cols[t];
    tb.cols.resize(cs.size());
    const int app = m_app_index(t);
    for (size_t c = 0; c < cs.size(); c++) {
      const uint32_t d = h->dev.desc_base[t] + (uint32_t)c;
      const MColDesc& dc = h->dev.descs[d];
      dfgpu_col& col = tb.cols[c];
      memset(&col, 0, sizeof col);
      col.name = cs[c].name;
      col.ch_type = cs[c].ch_type;
      col.elem_size = cs[c].elem_size;
      col.nrows = tb.nrows;
      col.data = hb + L.col_off[d];
      if (dc.kind == MK_LCS || dc.kind == MK_STR) {
        col.pos = (const dfgpu_pos*)(hb + L.aux_off[d]);
        col.data_bytes = L.str_bytes[app * 3 + dc.which];
      } else {
        col.data_bytes = tb.nrows * dc.elem;
        if (dc.kind == MK_OBS) lc_dict(ctx->h_dyn->stats.lc_first[m_obs_space(t)], 256, tap, col);
      }
    }
    b->counters.rows[tb.id] = tb.nrows;
  }
  {  // app_service_tag_column_block.go:39-80
    b->tables.emplace_back();
    dfgpu_batch::Table& tb = b->tables.back();
    tb.id = DFGPU_TABLE_METRICS_APP_SERVICE;
    tb.nrows = L.as_rows;
    tb.cols.resize(5);
    static const char* names[5] = {"time", "table", "app_service", "app_instance", "team_id"};
    static const int chs[5] = {DFGPU_CH_DATETIME, DFGPU_CH_LC_STR, DFGPU_CH_LC_STR_ROWS, DFGPU_CH_LC_STR_ROWS, DFGPU_CH_UINT16};
    static const int es[5] = {4, 1, 0, 0, 2};
    static const std::vector<std::string> tnames = {"application", "application_map"};
    for (int c = 0; c < 5; c++) {
      dfgpu_col& col = tb.cols[c];
      memset(&col, 0, sizeof col);
      col.name = names[c];
      col.ch_type = chs[c];
      col.elem_size = es[c];
      col.nrows = tb.nrows;
      if (c == 2 || c == 3) {
        col.pos = (const dfgpu_pos*)(hb + L.as_off[c]);
        col.data = hb + L.as_off[c + 3];
        col.data_bytes = L.as_bytes[c - 2];
      } else {
        col.data = hb + L.as_off[c];
        col.data_bytes = tb.nrows * es[c];
        if (c == 1) lc_dict(L.as_lc_first, 2, tnames, col);
      }
    }
    b->counters.rows[tb.id] = tb.nrows;
  }
  b->counters.count = L.count;
  b->counters.drop_count = L.drop_count;
  return 0;
}

}  // namespace dfgpu
