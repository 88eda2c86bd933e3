/*
 * dict_build.cu — host side of the SmartEncoding dictionary: turns one controller
 * snapshot (dfgpu_dict_snapshot) into the flat device tables of tables.cuh.
 * Replaces, with the same insertion order / overwrite rules:
 *   updatePlatformData   server/libs/grpc/grpc_platformdata.go:1090-1145
 *   updateInterfaceInfos :1494-1684      updateCidrInfos     :1422-1492
 *   updateGprocessInfos  :2008-2028      updateVtapIps       :1760-1791
 *   updatePodIps         :1861-1942      NewServiceTable     grpc_service.go:325-534
 * Runs once per dictionary version (every 60 s in the reference, :49), never per
 * record.
 */
#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.h"

namespace dfgpu {

namespace {

uint32_t pow2_at_least(uint64_t n) {
  uint32_t p = 16;
  while (p < n) p <<= 1;
  return p;
}

struct Builder {
  std::vector<T64Entry> t64;
  uint32_t mask = 0;
  std::vector<T6Entry> t6;
  uint32_t mask6 = 0;
  std::vector<DInfo> infos;
  std::vector<uint64_t> gp_vals;
  std::vector<DVtap> vtaps;
  uint32_t counts[TID_COUNT] = {0};
  uint32_t counts6[T6_COUNT] = {0};

  void init(uint64_t cap64, uint64_t cap6) {
    uint32_t n = pow2_at_least(cap64 * 2 + 16);
    T64Entry e;
    e.key = 0;
    e.val = 0;
    e.tid = TID_EMPTY;
    t64.assign(n, e);
    mask = n - 1;
    uint32_t n6 = pow2_at_least(cap6 * 2 + 16);
    T6Entry z;
    memset(&z, 0, sizeof z);
    t6.assign(n6, z);
    mask6 = n6 - 1;
  }
  // Go map assignment m[key] = val (insert or overwrite)
  void put(uint32_t tid, uint64_t key, uint32_t val) {
    uint32_t s = t64_slot(tid, key, mask);
    for (;;) {
      T64Entry& e = t64[s];
      if (e.tid == TID_EMPTY) {
        e.key = key;
        e.val = val;
        e.tid = tid;
        counts[tid]++;
        return;
      }
      if (e.tid == tid && e.key == key) {
        e.val = val;
        return;
      }
      s = (s + 1) & mask;
    }
  }
  bool get(uint32_t tid, uint64_t key, uint32_t& val) const {
    return t64_find(t64.data(), mask, tid, key, val);
  }
  void put6(uint8_t tid, const uint8_t* ip6, int32_t epc, uint16_t port, uint8_t proto, uint32_t val) {
    uint32_t s = (uint32_t)t6_hash(ip6, epc, port, proto, tid) & mask6;
    for (;;) {
      T6Entry& e = t6[s];
      if (e.tid == 0) {
        memcpy(e.ip6, ip6, 16);
        e.epc = epc;
        e.port = port;
        e.proto = proto;
        e.tid = tid;
        e.val = val;
        counts6[tid]++;
        return;
      }
      if (e.tid == tid && e.epc == epc && e.port == port && e.proto == proto && !memcmp(e.ip6, ip6, 16)) {
        e.val = val;
        return;
      }
      s = (s + 1) & mask6;
    }
  }
  uint32_t add_info(const DInfo& i) {
    infos.push_back(i);
    return (uint32_t)infos.size() - 1;
  }
};

struct CidrTmp {
  int32_t epc;
  bool v6;
  uint32_t plen, net4, mask4;
  uint8_t net6[16];
  uint32_t subnet_id, region_id, az_id;
  uint32_t order;
};

}  // namespace

int build_dict(const dfgpu_dict_snapshot& s, HostDict& out) {
  Builder b;
  uint64_t cap = 2 * s.n_ip_resources + 2 * s.n_interfaces + s.n_pod_ips + 2 * s.n_gprocesses +
                 s.n_vtap_ips + s.n_interfaces + 3 * s.n_cidrs + 64;
  uint64_t cap6 = 2 * s.n_ip_resources + 16;
  for (uint64_t i = 0; i < s.n_services; i++) {
    const dfgpu_service& v = s.services[i];
    uint64_t ports = (uint64_t)v.server_ports_cnt + 1;
    uint64_t ips = v.ips_cnt ? v.ips_cnt : 1;
    cap += 4 * ports * ips + v.pod_service_ids_cnt + v.pod_group_ids_cnt + v.pod_ids_cnt + v.chost_ids_cnt +
           v.host_ids_cnt + 8;
    cap6 += 4 * ports * ips + 8;
  }
  b.init(cap, cap6);
  b.infos.reserve(s.n_ip_resources * 2 + s.n_interfaces + s.n_pod_ips);

  // ---- interfaces (updateInterfaceInfos :1494-1684) ------------------------------
  for (uint64_t n = 0; n < s.n_interfaces; n++) {
    const dfgpu_interface& f = s.interfaces[n];
    int32_t epc = (int32_t)f.epc_id;
    if (epc == 0) epc = EPC_FROM_INTERNET;
    uint32_t tmp;
    if (!b.get(TID_EPC_BASE, (uint64_t)(uint32_t)epc, tmp)) b.put(TID_EPC_BASE, (uint64_t)(uint32_t)epc, f.region_id);
    DInfo base;
    memset(&base, 0, sizeof base);
    base.epc = epc;
    base.host_id = f.launch_server_id;
    base.region_id = f.region_id;
    base.device_type = f.device_type;
    base.device_id = f.device_id;
    base.pod_node_id = f.pod_node_id;
    base.pod_ns_id = f.pod_ns_id;
    base.pod_group_id = f.pod_group_id;
    base.pod_group_type = (uint8_t)f.pod_group_type;
    base.pod_id = f.pod_id;
    base.pod_cluster_id = f.pod_cluster_id;
    base.az_id = f.az_id;
    uint32_t first_subnet = 0;
    bool wan = f.if_type == 3;  // IF_TYPE_WAN
    for (uint32_t k = 0; k < f.ip_res_cnt; k++) {
      const dfgpu_ip_resource& r = s.ip_resources[f.ip_res_off + k];
      if (first_subnet == 0) first_subnet = r.subnet_id;
      if (r.ip.is_v6 && !r.ip.valid) continue;  // ParseIP failed
      DInfo i = base;
      i.subnet_id = r.subnet_id;
      uint32_t idx = b.add_info(i);
      if (!r.ip.is_v6) {
        b.put(TID_IP4, (uint64_t)(int64_t)epc << 32 | (uint64_t)r.ip.ip4, idx);
        if (f.pod_id > 0) b.put(TID_POD, f.pod_id, idx);
        if (wan) {  // extra entry with epc 0 carrying only EpcID (:1559-1565)
          DInfo w;
          memset(&w, 0, sizeof w);
          w.epc = epc;
          b.put(TID_IP4, (uint64_t)r.ip.ip4, b.add_info(w));
        }
      } else {
        b.put6(T6_IP6, r.ip.ip6, epc, 0, 0, idx);
        if (f.pod_id > 0) b.put(TID_POD, f.pod_id, idx);
        if (wan) {
          DInfo w;
          memset(&w, 0, sizeof w);
          w.epc = epc;
          b.put6(T6_IP6, r.ip.ip6, 0, 0, 0, b.add_info(w));
        }
      }
    }
    if (f.mac != 0) {  // :1618-1642
      uint64_t key = f.mac | (uint64_t)(int64_t)epc << 48;
      uint32_t idx;
      if (b.get(TID_MAC, key, idx)) {
        b.infos[idx].dup = 1;
      } else {
        DInfo m = base;
        m.l2_epc = epc;
        m.subnet_id = first_subnet;
        m.pod_ns_id = 0;  // the mac Info literal does not set PodNSID
        b.put(TID_MAC, key, b.add_info(m));
      }
    }
  }
  // ---- cidrs (updateCidrInfos :1422-1492) ----------------------------------------
  std::vector<CidrTmp> cidrs;
  for (uint64_t n = 0; n < s.n_cidrs; n++) {
    const dfgpu_cidr& c = s.cidrs[n];
    if (!c.prefix.valid) continue;
    CidrTmp t;
    memset(&t, 0, sizeof t);
    t.epc = c.epc_id == 0 ? EPC_FROM_INTERNET : c.epc_id;
    t.v6 = c.prefix.is_v6;
    t.plen = c.prefix_len;
    t.subnet_id = c.subnet_id;
    t.region_id = c.region_id;
    t.az_id = c.az_id;
    t.order = (uint32_t)cidrs.size();
    if (!t.v6) {
      t.mask4 = t.plen == 0 ? 0 : 0xffffffffu << (32 - t.plen);
      t.net4 = c.prefix.ip4 & t.mask4;
    } else {
      memcpy(t.net6, c.prefix.ip6, 16);
    }
    uint32_t tmp;
    if (!b.get(TID_EPC_BASE, (uint64_t)(uint32_t)t.epc, tmp)) b.put(TID_EPC_BASE, (uint64_t)(uint32_t)t.epc, c.region_id);
    cidrs.push_back(t);
    if (c.type == 1) {  // WAN: also listed under epc 0 (:1456-1462)
      CidrTmp w = t;
      w.epc = 0;
      w.order = (uint32_t)cidrs.size();
      cidrs.push_back(w);
    }
  }
  // group by (family, epc), longest prefix first; ties keep arrival order (the
  // reference's sort.Slice is unstable: ties must not matter)
  std::stable_sort(cidrs.begin(), cidrs.end(), [](const CidrTmp& a, const CidrTmp& c) {
    if (a.v6 != c.v6) return a.v6 < c.v6;
    if (a.epc != c.epc) return a.epc < c.epc;
    if (a.plen != c.plen) return a.plen > c.plen;
    return a.order < c.order;
  });
  out.cidr4.clear();
  out.cidr6.clear();
  out.cidr4_ranges.clear();
  out.cidr6_ranges.clear();
  for (size_t i = 0; i < cidrs.size();) {
    size_t j = i;
    while (j < cidrs.size() && cidrs[j].v6 == cidrs[i].v6 && cidrs[j].epc == cidrs[i].epc) j++;
    DRange rg;
    rg.count = (uint32_t)(j - i);
    if (!cidrs[i].v6) {
      rg.start = (uint32_t)out.cidr4.size();
      for (size_t k = i; k < j; k++)
        out.cidr4.push_back({cidrs[k].net4, cidrs[k].mask4, cidrs[k].subnet_id, cidrs[k].region_id, cidrs[k].az_id});
      b.put(TID_CIDR4, (uint64_t)(uint32_t)cidrs[i].epc, (uint32_t)out.cidr4_ranges.size());
      out.cidr4_ranges.push_back(rg);
    } else {
      rg.start = (uint32_t)out.cidr6.size();
      for (size_t k = i; k < j; k++) {
        DCidr6 c6;
        memcpy(c6.net, cidrs[k].net6, 16);
        c6.plen = cidrs[k].plen;
        c6.subnet_id = cidrs[k].subnet_id;
        c6.region_id = cidrs[k].region_id;
        c6.az_id = cidrs[k].az_id;
        out.cidr6.push_back(c6);
      }
      b.put(TID_CIDR6, (uint64_t)(uint32_t)cidrs[i].epc, (uint32_t)out.cidr6_ranges.size());
      out.cidr6_ranges.push_back(rg);
    }
    i = j;
  }
  // ---- gprocess infos (:2008-2028) -------------------------------------------------
  for (uint64_t n = 0; n < s.n_gprocesses; n++) {
    const dfgpu_gprocess& g = s.gprocesses[n];
    if (g.pid != 0) b.put(TID_VTAP_PID, (uint64_t)g.vtap_id << 32 | (uint64_t)g.pid, g.gprocess_id);
    if (g.vtap_id == 0 || g.gprocess_id == 0 || g.pod_id == 0) continue;
    uint32_t gi;
    uint64_t v = (uint64_t)g.vtap_id << 32 | (uint64_t)g.pod_id;
    if (b.get(TID_GPROC, g.gprocess_id, gi)) {
      b.gp_vals[gi] = v;
    } else {
      b.gp_vals.push_back(v);
      b.put(TID_GPROC, g.gprocess_id, (uint32_t)b.gp_vals.size() - 1);
    }
  }
  // ---- services (NewServiceTable grpc_service.go:325-386) --------------------------
  for (uint64_t n = 0; n < s.n_services; n++) {
    const dfgpu_service& v = s.services[n];
    uint32_t proto = v.protocol, sid = v.id;
    auto port_at = [&](uint32_t j) { return (uint16_t)s.u32_pool[v.server_ports_off + j]; };
    switch (v.type) {
      case 1: {  // POD_SERVICE_NODE
        for (uint32_t j = 0; j < v.server_ports_cnt; j++) {
          b.put(TID_PS_CLUSTERPORT, key_podx_port(v.pod_cluster_id, proto, port_at(j)), sid);
          b.put(TID_PS_CLUSTERPORT, key_podx_port(v.pod_cluster_id, 0, port_at(j)), sid);
        }
        b.put(TID_PS_CLUSTERPORT, key_podx_port(v.pod_cluster_id, proto, 0), sid);
        b.put(TID_PS_CLUSTERPORT, key_podx_port(v.pod_cluster_id, 0, 0), sid);
        break;
      }
      case 2: {  // POD_SERVICE_POD_GROUP (first group id only, :363)
        if (v.pod_group_ids_cnt == 0) break;
        uint32_t g = s.u32_pool[v.pod_group_ids_off];
        for (uint32_t j = 0; j < v.server_ports_cnt; j++) {
          b.put(TID_PS_GROUPPORT, key_podx_port(g, proto, port_at(j)), sid);
          b.put(TID_PS_GROUPPORT, key_podx_port(g, 0, port_at(j)), sid);
        }
        b.put(TID_PS_GROUPPORT, key_podx_port(g, proto, 0), sid);
        b.put(TID_PS_GROUPPORT, key_podx_port(g, 0, 0), sid);
        break;
      }
      case 3: {  // POD_SERVICE_IP addPodServiceIp :388-433
        if (proto > 2) break;
        int32_t epc = (int32_t)v.epc_id;
        for (uint32_t i = 0; i < v.ips_cnt; i++) {
          const dfgpu_ip& ip = s.ip_pool[v.ips_off + i];
          if (!ip.valid) continue;
          for (uint32_t j = 0; j <= v.server_ports_cnt; j++) {  // ports, then port 0
            uint16_t port = j < v.server_ports_cnt ? port_at(j) : 0;
            if (!ip.is_v6) {
              b.put(TID_PS_IP4PORT0 + proto, key_epc_ip4_port(epc, ip.ip4, port), sid);
              b.put(TID_PS_IP4PORT0, key_epc_ip4_port(epc, ip.ip4, port), sid);
            } else {
              b.put6(T6_PS_IP6PORT, ip.ip6, epc, port, (uint8_t)proto, sid);
              b.put6(T6_PS_IP6PORT, ip.ip6, epc, port, 0, sid);
            }
          }
        }
        break;
      }
      case 7: {  // CUSTOM_SERVICE addCustomService :435-534
        int32_t epc = (int32_t)v.epc_id;
        for (uint32_t i = 0; i < v.pod_service_ids_cnt; i++) b.put(TID_CS_PODSVC, s.u32_pool[v.pod_service_ids_off + i], sid);
        for (uint32_t i = 0; i < v.pod_group_ids_cnt; i++) b.put(TID_CS_PODGROUP, s.u32_pool[v.pod_group_ids_off + i], sid);
        for (uint32_t i = 0; i < v.pod_ids_cnt; i++) b.put(TID_CS_POD, s.u32_pool[v.pod_ids_off + i], sid);
        for (uint32_t i = 0; i < v.chost_ids_cnt; i++) b.put(TID_CS_CHOST, s.u32_pool[v.chost_ids_off + i], sid);
        for (uint32_t i = 0; i < v.host_ids_cnt; i++) b.put(TID_CS_HOST, s.u32_pool[v.host_ids_off + i], sid);
        if (v.pod_cluster_id != 0) {
          if (v.server_ports_cnt == 0) {
            b.put(TID_CS_CLUSTER, (uint64_t)v.pod_cluster_id << 32 | (uint64_t)proto << 16, sid);
            b.put(TID_CS_CLUSTERPORT, key_podx_port(v.pod_cluster_id, proto, 0), sid);
            b.put(TID_CS_CLUSTER, (uint64_t)v.pod_cluster_id << 32, sid);
            b.put(TID_CS_CLUSTERPORT, key_podx_port(v.pod_cluster_id, 0, 0), sid);
            break;
          }
          for (uint32_t j = 0; j <= v.server_ports_cnt; j++) {
            uint16_t port = j < v.server_ports_cnt ? port_at(j) : 0;
            b.put(TID_CS_CLUSTERPORT, key_podx_port(v.pod_cluster_id, proto, port), sid);
            b.put(TID_CS_CLUSTERPORT, key_podx_port(v.pod_cluster_id, 0, port), sid);
          }
          break;
        }
        if (v.ips_cnt == 0) break;
        const dfgpu_ip& ip = s.ip_pool[v.ips_off];
        if (!ip.valid) break;
        if (proto > 2) break;  // the reference would index past [3]map
        if (v.server_ports_cnt == 0) {
          if (!ip.is_v6) {
            b.put(TID_CS_IP4, key_epc_ip4_port(epc, ip.ip4, (uint16_t)proto), sid);
            b.put(TID_CS_IP4, key_epc_ip4_port(epc, ip.ip4, 0), sid);
            b.put(TID_CS_IP4PORT0 + proto, key_epc_ip4_port(epc, ip.ip4, 0), sid);
            b.put(TID_CS_IP4PORT0, key_epc_ip4_port(epc, ip.ip4, 0), sid);
          } else {
            b.put6(T6_CS_IP6, ip.ip6, epc, 0, (uint8_t)proto, sid);
            b.put6(T6_CS_IP6, ip.ip6, epc, 0, 0, sid);
            b.put6(T6_CS_IP6PORT, ip.ip6, epc, 0, (uint8_t)proto, sid);
            b.put6(T6_CS_IP6PORT, ip.ip6, epc, 0, 0, sid);
          }
          break;
        }
        for (uint32_t j = 0; j <= v.server_ports_cnt; j++) {
          uint16_t port = j < v.server_ports_cnt ? port_at(j) : 0;
          if (!ip.is_v6) {
            b.put(TID_CS_IP4PORT0 + proto, key_epc_ip4_port(epc, ip.ip4, port), sid);
            b.put(TID_CS_IP4PORT0, key_epc_ip4_port(epc, ip.ip4, port), sid);
          } else {
            b.put6(T6_CS_IP6PORT, ip.ip6, epc, port, (uint8_t)proto, sid);
            b.put6(T6_CS_IP6PORT, ip.ip6, epc, port, 0, sid);
          }
        }
        break;
      }
      default: break;
    }
  }
  // ---- vtap ips (:1760-1791) --------------------------------------------------------
  for (uint64_t n = 0; n < s.n_vtap_ips; n++) {
    const dfgpu_vtap_ip& v = s.vtap_ips[n];
    DVtap d;
    memset(&d, 0, sizeof d);
    d.epc = v.epc_id == 0 ? EPC_FROM_INTERNET : (int32_t)v.epc_id;
    d.is_v4 = 1;
    if (v.ip.valid) {
      if (!v.ip.is_v6) {
        d.ip4 = v.ip.ip4;
      } else {
        d.is_v4 = 0;
        d.has_ip6 = 1;
        memcpy(d.ip6, v.ip.ip6, 16);
      }
    }
    uint32_t vi;
    if (b.get(TID_VTAP, (uint16_t)v.vtap_id, vi)) {
      b.vtaps[vi] = d;
    } else {
      b.vtaps.push_back(d);
      b.put(TID_VTAP, (uint16_t)v.vtap_id, (uint32_t)b.vtaps.size() - 1);
    }
  }
  // ---- pod ips (:1861-1942): pods unknown to the interface list get an Info resolved
  // through the SAME lookup code the kernels run (host view of the tables) -----------
  out.gp_vals = b.gp_vals;
  out.vtaps = b.vtaps;
  {
    DevDict hv;
    memset(&hv, 0, sizeof hv);
    hv.t64 = b.t64.data();
    hv.t64_mask = b.mask;
    hv.t6 = b.t6.data();
    hv.t6_mask = b.mask6;
    hv.cidr4 = out.cidr4.data();
    hv.cidr4_ranges = out.cidr4_ranges.data();
    hv.cidr6 = out.cidr6.data();
    hv.cidr6_ranges = out.cidr6_ranges.data();
    hv.region_id = s.region_id;
    for (uint64_t n = 0; n < s.n_pod_ips; n++) {
      const dfgpu_pod_ip& p = s.pod_ips[n];
      uint32_t tmp;
      if (b.get(TID_POD, p.pod_id, tmp)) continue;
      int32_t epc = p.epc_id == 0 ? EPC_FROM_INTERNET : (int32_t)p.epc_id;
      const dfgpu_ip* ip = p.ip.valid ? &p.ip : &p.pod_node_ip;
      DInfo i;
      memset(&i, 0, sizeof i);
      if (ip->valid) {
        hv.infos = b.infos.data();  // vector may have grown
        InfoVal q;
        bool found = !ip->is_v6 ? query_ip4(hv, epc, ip->ip4, q) : query_ip6(hv, epc, ip->ip6, 16, q);
        if (found) {
          i.region_id = q.region_id;
          i.az_id = q.az_id;
          i.host_id = q.host_id;
          i.device_type = q.device_type;
          i.device_id = q.device_id;
          i.pod_node_id = q.pod_node_id;
          i.subnet_id = q.subnet_id;
          i.l2_epc = q.l2_epc;
        }
      }
      i.pod_id = p.pod_id;
      i.pod_cluster_id = p.pod_cluster_id;
      i.pod_ns_id = p.pod_ns_id;
      i.pod_group_id = p.pod_group_id;
      i.pod_group_type = (uint8_t)p.pod_group_type;
      i.epc = epc;
      b.put(TID_POD, p.pod_id, b.add_info(i));
    }
  }
  out.t64.swap(b.t64);
  out.t64_mask = b.mask;
  out.t6.swap(b.t6);
  out.t6_mask = b.mask6;
  out.infos.swap(b.infos);
  memcpy(out.counts, b.counts, sizeof b.counts);
  memcpy(out.counts6, b.counts6, sizeof b.counts6);
  out.region_id = s.region_id;
  out.analyzer_id = s.analyzer_id;
  out.version = s.version;
  return 0;
}

}  // namespace dfgpu
