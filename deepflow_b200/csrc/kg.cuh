/*
 * kg.cuh — SmartEncoding tag injection on the device: one thread resolves both
 * sides of one record against the DevDict tables.
 * Replaces  KnowledgeGraph.fill           flow_log/log_data/l4_flow_log.go:536-728
 *           ServiceTable.QueryPodService  libs/grpc/grpc_service.go:138-193
 *           ServiceTable.QueryCustomService               :195-323
 *           GetAutoInstance/GetAutoService/IsPodServiceIP ingester/common/common.go:160-198
 *           utils.IsLoopback / IsLocalIP  libs/utils/utils.go:116-130 (+ Go net.IP rules)
 */
#pragma once
#include "tables.cuh"

namespace dfgpu {

enum : uint8_t {  // flow-metrics/tag.go:255-266
  TS_GPID = 1, TS_PODID = 2, TS_MAC = 4, TS_EPCIP = 8, TS_PEER = 16, TS_AGENT = 32, TS_PROCESSID = 64
};

struct KG {  // l4_flow_log.go:175-226
  uint16_t region_id[2], az_id[2], host_id[2];
  uint8_t l3_device_type[2];
  uint32_t l3_device_id[2], pod_node_id[2];
  uint16_t pod_ns_id[2];
  uint32_t pod_group_id[2];
  uint8_t pod_group_type[2];
  uint32_t pod_id[2];
  uint16_t pod_cluster_id[2];
  int32_t l3_epc_id[2], epc_id[2];
  uint16_t subnet_id[2];
  uint32_t service_id[2];
  uint32_t auto_instance_id[2], auto_service_id[2];
  uint8_t auto_instance_type[2], auto_service_type[2];
  uint8_t tag_source[2];
};

struct IpAddr {  // one endpoint address as the wire carried it
  uint32_t ip4;
  uint8_t ip6[16];  // zero padded
  uint32_t ip6len;  // wire length (0 when absent)
};

// Go net.IP.To4() on a 16-byte slice
DF_HD bool go_to4(const IpAddr& a, uint8_t v4[4]) {
  if (a.ip6len == 4) {
    for (int i = 0; i < 4; i++) v4[i] = a.ip6[i];
    return true;
  }
  if (a.ip6len == 16 && ip6_is_v4mapped(a.ip6)) {
    for (int i = 0; i < 4; i++) v4[i] = a.ip6[12 + i];
    return true;
  }
  return false;
}
DF_HD bool v4_is_global_unicast(uint8_t a, uint8_t b, uint8_t c, uint8_t d) {
  if (a == 255 && b == 255 && c == 255 && d == 255) return false;  // IPv4bcast
  if ((a | b | c | d) == 0) return false;                           // unspecified
  if (a == 127) return false;                                       // loopback
  if ((a & 0xf0) == 0xe0) return false;                             // multicast
  if (a == 169 && b == 254) return false;                           // link-local unicast
  return true;
}
// utils.IsLoopback
DF_KG_FN bool is_loopback(bool v6, const IpAddr& a) {
  if (!v6) return (a.ip4 & 0xFF000000u) == 0x7F000000u;
  uint8_t v4[4];
  if (go_to4(a, v4)) return v4[0] == 127;
  if (a.ip6len != 16) return false;
  for (int i = 0; i < 15; i++)
    if (a.ip6[i]) return false;
  return a.ip6[15] == 1;
}
// utils.IsLocalIP = !ip.IsGlobalUnicast()
DF_KG_FN bool is_local_ip(bool v6, const IpAddr& a) {
  if (!v6)
    return !v4_is_global_unicast((uint8_t)(a.ip4 >> 24), (uint8_t)(a.ip4 >> 16), (uint8_t)(a.ip4 >> 8),
                                 (uint8_t)a.ip4);
  if (a.ip6len != 4 && a.ip6len != 16) return true;
  uint8_t v4[4];
  if (go_to4(a, v4)) return !v4_is_global_unicast(v4[0], v4[1], v4[2], v4[3]);
  bool zero = true;
  for (int i = 0; i < 16; i++) zero &= a.ip6[i] == 0;
  if (zero) return true;
  bool lo = true;
  for (int i = 0; i < 15; i++) lo &= a.ip6[i] == 0;
  if (lo && a.ip6[15] == 1) return true;
  if (a.ip6[0] == 0xff) return true;
  if (a.ip6[0] == 0xfe && (a.ip6[1] & 0xc0) == 0x80) return true;
  return false;
}

DF_HD bool reget_info_from_ip(const DevDict& d, bool v6, const IpAddr& a, int32_t epc, InfoVal& out) {
  return v6 ? query_ip6(d, epc, a.ip6, a.ip6len, out) : query_ip4(d, epc, a.ip4, out);
}

DF_HD uint32_t to_service_protocol(uint8_t p) { return p == 6 ? 1u : p == 17 ? 2u : 0u; }
DF_HD uint64_t key_epc_ip4_port(int32_t epc, uint32_t ip, uint16_t port) {
  return (uint64_t)(int64_t)epc << 48 | (uint64_t)ip << 16 | (uint64_t)port;
}
DF_HD uint64_t key_podx_port(uint32_t id, uint32_t proto, uint16_t port) {
  return (uint64_t)id << 32 | (uint64_t)proto << 16 | (uint64_t)port;
}
DF_HD uint32_t t64_getz(const DevDict& d, uint32_t tid, uint64_t key) {
  uint32_t v;
  return t64_find(d.t64, d.t64_mask, tid, key, v) ? v : 0u;
}
DF_HD uint32_t t6_getz(const DevDict& d, uint8_t tid, const uint8_t* ip6, int32_t epc, uint16_t port, uint8_t proto) {
  uint32_t v;
  return t6_find(d.t6, d.t6_mask, tid, ip6, epc, port, proto, v) ? v : 0u;
}

// grpc_service.go:138-193
DF_KG_FN uint32_t query_pod_service(const DevDict& d, uint32_t pod_id, uint32_t pod_node_id,
                                 uint32_t pod_cluster_id, uint32_t pod_group_id, int32_t epc, bool v6,
                                 const IpAddr& a, uint8_t protocol, uint16_t server_port) {
  if (server_port == 0) protocol = 0;
  uint32_t sp = to_service_protocol(protocol);
  uint32_t sid = 0;
  if (pod_id != 0) {
    sid = t64_getz(d, TID_PS_GROUPPORT, key_podx_port(pod_group_id, sp, server_port));
    if (sid == 0) {
      InfoVal info;
      if (reget_info_from_ip(d, v6, a, epc, info) && info.pod_id != pod_id) return 0;
    }
  } else if (pod_node_id != 0 && server_port != 0) {
    sid = t64_getz(d, TID_PS_CLUSTERPORT, key_podx_port(pod_cluster_id, sp, server_port));
  }
  if (sid != 0) return sid;
  if (epc <= 0) return 0;
  if (v6) return t6_getz(d, T6_PS_IP6PORT, a.ip6, epc, server_port, (uint8_t)sp);
  return t64_getz(d, TID_PS_IP4PORT0 + sp, key_epc_ip4_port(epc, a.ip4, server_port));
}

// grpc_service.go:195-323
DF_KG_FN uint32_t query_custom_service(const DevDict& d, int32_t epc, bool v6, const IpAddr& a,
                                    uint16_t server_port, uint32_t pod_service_id, uint32_t pod_group_id,
                                    uint32_t l3_device_id, uint32_t pod_id, uint32_t pod_cluster_id,
                                    uint8_t l3_device_type, uint8_t protocol) {
  if (epc < 0) return 0;
  if (server_port == 0) protocol = 0;
  uint32_t sp = to_service_protocol(protocol);
  uint32_t sid = 0;
  if (pod_service_id != 0 && d.counts[TID_CS_PODSVC]) {
    uint32_t v = t64_getz(d, TID_CS_PODSVC, pod_service_id);
    if (v) return v;
  }
  if (pod_group_id != 0 && d.counts[TID_CS_PODGROUP]) {
    uint32_t v = t64_getz(d, TID_CS_PODGROUP, pod_group_id);
    if (v) return v;
  }
  if (pod_id != 0 && d.counts[TID_CS_POD]) {
    uint32_t v = t64_getz(d, TID_CS_POD, pod_id);
    if (v) return v;
  }
  if (pod_cluster_id != 0 && server_port != 0) {
    if (d.counts[TID_CS_CLUSTERPORT]) {
      sid = t64_getz(d, TID_CS_CLUSTERPORT, key_podx_port(pod_cluster_id, sp, server_port));
      if (sid) return sid;
    }
    if (d.counts[TID_CS_CLUSTER]) {
      sid = t64_getz(d, TID_CS_CLUSTER, (uint64_t)pod_cluster_id << 32);
      if (sid) return sid;
    }
  }
  if (l3_device_id != 0 && l3_device_type == 1 && d.counts[TID_CS_CHOST]) {  // VMDevice
    uint32_t v = t64_getz(d, TID_CS_CHOST, l3_device_id);
    if (v) return v;
  }
  if (l3_device_id != 0 && l3_device_type == 6 && d.counts[TID_CS_HOST]) {  // HostDevice
    uint32_t v = t64_getz(d, TID_CS_HOST, l3_device_id);
    if (v) return v;
  }
  if (v6) {
    if (d.counts6[T6_CS_IP6PORT]) {
      sid = t6_getz(d, T6_CS_IP6PORT, a.ip6, epc, server_port, (uint8_t)sp);
      if (sid) return sid;
    }
    if (d.counts6[T6_CS_IP6] && server_port != 0) {
      sid = t6_getz(d, T6_CS_IP6, a.ip6, epc, 0, 0);
      if (sid) return sid;
    }
    if (d.counts6[T6_CS_IP6PORT]) {
      sid = t6_getz(d, T6_CS_IP6PORT, a.ip6, 0, server_port, (uint8_t)sp);
      if (sid) return sid;
    }
    if (d.counts6[T6_CS_IP6] && server_port != 0) sid = t6_getz(d, T6_CS_IP6, a.ip6, 0, 0, 0);
    return sid;
  }
  if (d.counts[TID_CS_IP4PORT0 + sp]) {
    sid = t64_getz(d, TID_CS_IP4PORT0 + sp, key_epc_ip4_port(epc, a.ip4, server_port));
    if (sid) return sid;
  }
  if (d.counts[TID_CS_IP4] && server_port != 0) {
    sid = t64_getz(d, TID_CS_IP4, key_epc_ip4_port(epc, a.ip4, 0));  // genEpcIDIP4Key(.., protocol 0)
    if (sid) return sid;
  }
  if (d.counts[TID_CS_IP4PORT0 + sp]) {
    sid = t64_getz(d, TID_CS_IP4PORT0 + sp, key_epc_ip4_port(0, a.ip4, server_port));
    if (sid) return sid;
  }
  if (d.counts[TID_CS_IP4] && server_port != 0) sid = t64_getz(d, TID_CS_IP4, key_epc_ip4_port(0, a.ip4, 0));
  return sid;
}

DF_HD void get_auto_instance(uint32_t pod, uint32_t gpid, uint32_t pod_node, uint32_t l3dev, uint32_t subnet,
                             uint8_t l3type, int32_t l3epc, uint32_t& id, uint8_t& type) {
  if (pod > 0) { id = pod; type = 10; }
  else if (gpid > 0) { id = gpid; type = 120; }
  else if (pod_node > 0) { id = pod_node; type = 14; }
  else if (l3dev > 0) { id = l3dev; type = l3type; }
  else if (l3epc == EPC_FROM_INTERNET) { id = 0; type = 0; }
  else { id = subnet; type = 255; }
}
DF_HD void get_auto_service(uint32_t custom, uint32_t pod_service, uint32_t pod_group, uint32_t gpid,
                            uint32_t pod_cluster, uint32_t l3dev, uint32_t subnet, uint8_t l3type,
                            uint8_t pod_group_type, int32_t l3epc, uint32_t& id, uint8_t& type) {
  if (custom > 0) { id = custom; type = 104; }
  else if (pod_service > 0) { id = pod_service; type = 11; }
  else if (pod_group > 0) { id = pod_group; type = pod_group_type; }
  else if (gpid > 0) { id = gpid; type = 120; }
  else if (pod_cluster > 0) { id = pod_cluster; type = 103; }
  else if (l3dev > 0) { id = l3dev; type = l3type; }
  else if (l3epc == EPC_FROM_INTERNET) { id = 0; type = 0; }
  else { id = subnet; type = 255; }
}

struct KGInput {
  bool v6;
  bool vip[2];
  int32_t l3_epc[2];
  IpAddr ip[2];
  uint64_t mac[2];
  uint32_t gpid[2];
  uint32_t pod_id[2];
  uint16_t vtap_id;
  uint16_t port;
  uint32_t tap_side;
  uint8_t protocol;
};

DF_HD void kg_copy_info(KG& k, int s, const InfoVal& i) {  // :677-704
  k.region_id[s] = (uint16_t)i.region_id;
  k.az_id[s] = (uint16_t)i.az_id;
  k.host_id[s] = (uint16_t)i.host_id;
  k.l3_device_type[s] = (uint8_t)i.device_type;
  k.l3_device_id[s] = i.device_id;
  k.pod_node_id[s] = i.pod_node_id;
  k.pod_ns_id[s] = (uint16_t)i.pod_ns_id;
  k.pod_group_id[s] = i.pod_group_id;
  k.pod_group_type[s] = i.pod_group_type;
  k.pod_id[s] = i.pod_id;
  k.pod_cluster_id[s] = (uint16_t)i.pod_cluster_id;
  k.subnet_id[s] = (uint16_t)i.subnet_id;
}

DF_KG_FN void kg_fill(const DevDict& d, const KGInput& in, KG& k) {
DF_KG_SIDE_LOOP
  for (int s = 0; s < 2; s++) {
    k.region_id[s] = k.az_id[s] = k.host_id[s] = 0;
    k.l3_device_type[s] = 0;
    k.l3_device_id[s] = k.pod_node_id[s] = 0;
    k.pod_ns_id[s] = 0;
    k.pod_group_id[s] = 0;
    k.pod_group_type[s] = 0;
    k.pod_id[s] = 0;
    k.pod_cluster_id[s] = 0;
    k.l3_epc_id[s] = k.epc_id[s] = 0;
    k.subnet_id[s] = 0;
    k.service_id[s] = 0;
    k.tag_source[s] = 0;
  }
  uint32_t pod[2] = {in.pod_id[0], in.pod_id[1]};
  bool by_mac[2] = {in.vip[0], in.vip[1]};
  bool by_agent[2] = {false, false};
  // :556-567 loopback end borrows the peer's pod (side 0 first)
  if (is_loopback(in.v6, in.ip[0]) && pod[0] == 0 && pod[1] != 0) {
    k.tag_source[0] |= TS_PEER;
    pod[0] = pod[1];
  }
  if (is_loopback(in.v6, in.ip[1]) && pod[1] == 0 && pod[0] != 0) {
    k.tag_source[1] |= TS_PEER;
    pod[1] = pod[0];
  }
  if (in.tap_side == 4) {  // Local :570-585
DF_KG_SIDE_LOOP
    for (int s = 0; s < 2; s++)
      if (is_local_ip(in.v6, in.ip[s])) {
        if (in.mac[s] != 0) by_mac[s] = true; else by_agent[s] = true;
      }
  } else if (in.tap_side == 41 || in.tap_side == 42) {  // ClientProcess / ServerProcess :586-598
DF_KG_SIDE_LOOP
    for (int s = 0; s < 2; s++) {
      if (in.mac[s] != 0) by_mac[s] = true;
      else if (is_local_ip(in.v6, in.ip[s])) by_agent[s] = true;
    }
  }
  uint64_t l3mac[2] = {in.mac[0] | (uint64_t)(int64_t)in.l3_epc[0] << 48,
                       in.mac[1] | (uint64_t)(int64_t)in.l3_epc[1] << 48};
DF_KG_SIDE_LOOP
  for (int s = 0; s < 2; s++) {  // :601-614 gpid -> pod
    if (in.gpid[s] != 0 && pod[s] == 0) {
      uint32_t gi;
      if (t64_find(d.t64, d.t64_mask, TID_GPROC, in.gpid[s], gi)) {
        uint64_t vp = d.gp_vals[gi];
        uint32_t p = (uint32_t)vp;
        if (p != 0 && (uint16_t)(vp >> 32) == in.vtap_id) {
          pod[s] = p;
          k.tag_source[s] |= TS_GPID;
        }
      }
    }
  }
  InfoVal info[2], agent;
  bool has[2] = {false, false}, has_agent = false;
DF_KG_SIDE_LOOP
  for (int s = 0; s < 2; s++)
    if (pod[s] != 0) {  // :617-624
      k.tag_source[s] |= TS_PODID;
      has[s] = query_pod(d, pod[s], info[s]);
    }
DF_KG_SIDE_LOOP
  for (int s = 0; s < 2; s++) {  // :626-666
    if (has[s]) continue;
    if (by_mac[s]) {
      k.tag_source[s] |= TS_MAC;
      has[s] = query_mac(d, l3mac[s], info[s]);
    }
    if (!has[s] && by_agent[s]) {
      k.tag_source[s] |= TS_AGENT;
      if (s == 1 && by_agent[0] && has_agent) {
        info[1] = agent;
        has[1] = true;
      } else {
        uint32_t vi;
        if (t64_find(d.t64, d.t64_mask, TID_VTAP, in.vtap_id, vi)) {
          const DVtap& v = d.vtaps[vi];
          IpAddr a;
          a.ip4 = v.ip4;
          for (int i = 0; i < 16; i++) a.ip6[i] = v.ip6[i];
          a.ip6len = v.has_ip6 ? 16 : 0;
          InfoVal t;
          bool f = reget_info_from_ip(d, !v.is_v4, a, v.epc, t);
          if (f) {
            info[s] = t;
            has[s] = true;
          }
          if (s == 0) {
            has_agent = f;
            if (f) agent = t;
          }
        }
      }
    }
    if (!has[s]) {
      k.tag_source[s] |= TS_EPCIP;
      has[s] = reget_info_from_ip(d, in.v6, in.ip[s], in.l3_epc[s], info[s]);
    }
  }
  // :668-675 L2 epc through the mac table (only sides with l3_epc > 0)
DF_KG_SIDE_LOOP
  for (int s = 0; s < 2; s++) {
    bool want = in.l3_epc[s] > 0;
    if (want) {
      InfoVal l2;
      if (query_mac(d, l3mac[s], l2)) k.epc_id[s] = l2.l2_epc;
    }
  }
DF_KG_SIDE_LOOP
  for (int s = 0; s < 2; s++) {
    if (has[s]) kg_copy_info(k, s, info[s]);
    k.l3_epc_id[s] = in.l3_epc[s];
  }
  // :713-719
  if (k.l3_device_type[0] == 11 || k.pod_id[0] != 0)
    k.service_id[0] = query_pod_service(d, k.pod_id[0], k.pod_node_id[0], k.pod_cluster_id[0],
                                        k.pod_group_id[0], in.l3_epc[0], in.v6, in.ip[0], in.protocol, 0);
  if (k.l3_device_type[1] == 11 || k.pod_id[1] != 0 || k.pod_node_id[1] != 0)
    k.service_id[1] = query_pod_service(d, k.pod_id[1], k.pod_node_id[1], k.pod_cluster_id[1],
                                        k.pod_group_id[1], in.l3_epc[1], in.v6, in.ip[1], in.protocol, in.port);
DF_KG_SIDE_LOOP
  for (int s = 0; s < 2; s++) {  // :721-727
    uint16_t port = s == 0 ? 0 : in.port;
    get_auto_instance(k.pod_id[s], in.gpid[s], k.pod_node_id[s], k.l3_device_id[s], k.subnet_id[s],
                      k.l3_device_type[s], k.l3_epc_id[s], k.auto_instance_id[s], k.auto_instance_type[s]);
    uint32_t custom = query_custom_service(d, in.l3_epc[s], in.v6, in.ip[s], port, k.service_id[s],
                                           k.pod_group_id[s], k.l3_device_id[s], k.pod_id[s],
                                           k.pod_cluster_id[s], k.l3_device_type[s], in.protocol);
    get_auto_service(custom, k.service_id[s], k.pod_group_id[s], in.gpid[s], k.pod_cluster_id[s],
                     k.l3_device_id[s], k.subnet_id[s], k.l3_device_type[s], k.pod_group_type[s],
                     k.l3_epc_id[s], k.auto_service_id[s], k.auto_service_type[s]);
  }
}

// ---- one SIDE at a time ------------------------------------------------------------------------
// KnowledgeGraph.fill treats its two endpoints independently except for one step: a loopback
// endpoint without a pod borrows the peer's pod (:556-567, side 0 first).  (The "reuse the agent
// info found for side 0" shortcut :651-660 is only a cache: the agent is looked up by vtap id, the
// same for both sides.)  l7_kg.cu runs one lane per (record, side) through kg_side.cuh; these are the
// pieces it shares with this file.
struct KGSide {
  uint16_t region_id, az_id, host_id;
  uint8_t l3_device_type;
  uint32_t l3_device_id, pod_node_id;
  uint16_t pod_ns_id;
  uint32_t pod_group_id;
  uint8_t pod_group_type;
  uint32_t pod_id;
  uint16_t pod_cluster_id;
  int32_t l3_epc_id, epc_id;
  uint16_t subnet_id;
  uint32_t service_id, auto_instance_id, auto_service_id;
  uint8_t auto_instance_type, auto_service_type, tag_source;
};

// :556-567 — both results from both inputs (side 0 is resolved first)
DF_HD void kg_peer_pods(bool lb0, bool lb1, uint32_t& pod0, uint32_t& pod1, uint8_t& ts0, uint8_t& ts1) {
  if (lb0 && pod0 == 0 && pod1 != 0) {
    ts0 |= TS_PEER;
    pod0 = pod1;
  }
  if (lb1 && pod1 == 0 && pod0 != 0) {
    ts1 |= TS_PEER;
    pod1 = pod0;
  }
}

}  // namespace dfgpu
