/*
 * ORACLE — test infrastructure only.  See platform.h for the reference map.
 * Parity unpinned by the reference (no test builds a PlatformInfoTable / ServiceTable); cross-checked by the second
 * restatement tests/kg_second_opinion.py.  Pinned pieces: Go net.IP predicates (stdlib semantics, KATs in
 * tests/test_reference_kats_cpu.py), the netmask tree (libs/geo/netmask_tree_test.go).
 */
#include "platform.h"

namespace oracle {

// ---- Go net.IP predicates (stdlib net/ip.go) restated -----------------------
static bool to4(const uint8_t* ip, size_t len, uint8_t out[4]) {
  if (len == 4) {
    memcpy(out, ip, 4);
    return true;
  }
  if (len == 16) {
    for (int i = 0; i < 10; i++)
      if (ip[i]) return false;
    if (ip[10] != 0xff || ip[11] != 0xff) return false;
    memcpy(out, ip + 12, 4);
    return true;
  }
  return false;
}
static bool go_IsLoopback(const uint8_t* ip, size_t len) {
  uint8_t v4[4];
  if (to4(ip, len, v4)) return v4[0] == 127;
  if (len != 16) return false;
  for (int i = 0; i < 15; i++)
    if (ip[i]) return false;
  return ip[15] == 1;
}
static bool go_IsGlobalUnicast(const uint8_t* ip, size_t len) {
  if (len != 4 && len != 16) return false;
  uint8_t v4[4];
  if (to4(ip, len, v4)) {
    if (v4[0] == 255 && v4[1] == 255 && v4[2] == 255 && v4[3] == 255) return false;  // bcast
    if (!v4[0] && !v4[1] && !v4[2] && !v4[3]) return false;                            // 0.0.0.0
    if (v4[0] == 127) return false;                                                    // loopback
    if ((v4[0] & 0xf0) == 0xe0) return false;                                          // multicast
    if (v4[0] == 169 && v4[1] == 254) return false;                                    // link-local
    return true;
  }
  bool allzero = true;
  for (int i = 0; i < 16; i++)
    if (ip[i]) allzero = false;
  if (allzero) return false;                                   // ::
  if (go_IsLoopback(ip, len)) return false;                    // ::1
  if (ip[0] == 0xff) return false;                             // multicast ff00::/8
  if (ip[0] == 0xfe && (ip[1] & 0xc0) == 0x80) return false;   // fe80::/10
  return true;
}

bool IsLoopback(bool isIPv6, uint32_t ip4, const uint8_t* ip6, size_t ip6len) {
  if (isIPv6) return go_IsLoopback(ip6, ip6len);
  return (ip4 & 0xFF000000u) == 0x7F000000u;
}
bool IsLocalIP(bool isIPv6, uint32_t ip4, const uint8_t* ip6, size_t ip6len) {
  if (!isIPv6) {
    uint8_t b[4] = {(uint8_t)(ip4 >> 24), (uint8_t)(ip4 >> 16), (uint8_t)(ip4 >> 8), (uint8_t)ip4};
    return !go_IsGlobalUnicast(b, 4);
  }
  return !go_IsGlobalUnicast(ip6, ip6len);
}

// ---- ServiceTable -----------------------------------------------------------
static inline uint64_t genEpcIDIP4PortKey(int32_t epcID, uint32_t ipv4, uint16_t port) {
  return (uint64_t)(int64_t)epcID << 48 | (uint64_t)ipv4 << 16 | (uint64_t)port;  // :66-68
}
static inline uint64_t genEpcIDIP4Key(int32_t epcID, uint32_t ipv4, uint32_t protocol) {
  return (uint64_t)(int64_t)epcID << 48 | (uint64_t)ipv4 << 16 | (uint64_t)protocol;  // :74-76
}
static inline uint64_t genPodXIDKey(uint32_t id, uint32_t protocol) {
  return (uint64_t)id << 32 | (uint64_t)protocol << 16;  // :112-114
}
static inline uint64_t genPodXIDPortKey(uint32_t id, uint32_t protocol, uint16_t port) {
  return (uint64_t)id << 32 | (uint64_t)protocol << 16 | (uint64_t)port;  // :121-123
}
static EpcIDIP6PortKey genEpcIDIP6PortKey(int32_t epc, const uint8_t* ip6, size_t len,
                                          uint32_t protocol, uint16_t port) {
  EpcIDIP6PortKey k;
  k.epc = epc;
  k.port = port;
  k.protocol = (uint8_t)protocol;
  memcpy(k.ip.data(), ip6, len > 16 ? 16 : len);
  return k;
}
static EpcIDIP6Key genEpcIDIP6Key(int32_t epc, const uint8_t* ip6, size_t len, uint32_t protocol) {
  EpcIDIP6Key k;
  k.epc = epc;
  k.protocol = (uint8_t)protocol;
  memcpy(k.ip.data(), ip6, len > 16 ? 16 : len);
  return k;
}
static inline uint32_t toServiceProtocol(uint8_t protocol) {  // :127-136
  if (protocol == 6) return 1;
  if (protocol == 17) return 2;
  return 0;
}
template <class M, class K>
static inline uint32_t getz(const M& m, const K& k) {
  auto it = m.find(k);
  return it == m.end() ? 0 : it->second;
}

void ServiceTable::build(const dfgpu_dict_snapshot& s) {
  for (uint64_t i = 0; i < s.n_services; i++) {
    const dfgpu_service& svc = s.services[i];
    uint32_t protocol = svc.protocol, serviceId = svc.id;
    switch (svc.type) {
      case 1: {  // POD_SERVICE_NODE :348-357
        uint32_t podClusterId = svc.pod_cluster_id;
        for (uint32_t j = 0; j < svc.server_ports_cnt; j++) {
          uint16_t port = (uint16_t)s.u32_pool[svc.server_ports_off + j];
          podServiceClusterIDPortTable[genPodXIDPortKey(podClusterId, protocol, port)] = serviceId;
          podServiceClusterIDPortTable[genPodXIDPortKey(podClusterId, 0, port)] = serviceId;
        }
        podServiceClusterIDPortTable[genPodXIDPortKey(podClusterId, protocol, 0)] = serviceId;
        podServiceClusterIDPortTable[genPodXIDPortKey(podClusterId, 0, 0)] = serviceId;
        break;
      }
      case 2: {  // POD_SERVICE_POD_GROUP :359-373
        if (svc.pod_group_ids_cnt == 0) break;
        uint32_t podGroupId = s.u32_pool[svc.pod_group_ids_off];
        for (uint32_t j = 0; j < svc.server_ports_cnt; j++) {
          uint16_t port = (uint16_t)s.u32_pool[svc.server_ports_off + j];
          podServiceGroupIDPortTable[genPodXIDPortKey(podGroupId, protocol, port)] = serviceId;
          podServiceGroupIDPortTable[genPodXIDPortKey(podGroupId, 0, port)] = serviceId;
        }
        podServiceGroupIDPortTable[genPodXIDPortKey(podGroupId, protocol, 0)] = serviceId;
        podServiceGroupIDPortTable[genPodXIDPortKey(podGroupId, 0, 0)] = serviceId;
        break;
      }
      case 3: addPodServiceIp(s, svc); break;
      case 7: addCustomService(s, svc); break;
      default: break;
    }
  }
}

void ServiceTable::addPodServiceIp(const dfgpu_dict_snapshot& s, const dfgpu_service& svc) {
  uint32_t protocol = svc.protocol;
  if (protocol > 2) return;
  int32_t epcId = (int32_t)svc.epc_id;
  uint32_t serviceId = svc.id;
  for (uint32_t i = 0; i < svc.ips_cnt; i++) {
    const dfgpu_ip& ip = s.ip_pool[svc.ips_off + i];
    if (!ip.valid) continue;
    if (!ip.is_v6) {
      for (uint32_t j = 0; j < svc.server_ports_cnt; j++) {
        uint64_t key = genEpcIDIP4PortKey(epcId, ip.ip4, (uint16_t)s.u32_pool[svc.server_ports_off + j]);
        podServiceClusterIP4PortTable[protocol][key] = serviceId;
        podServiceClusterIP4PortTable[0][key] = serviceId;
      }
      uint64_t key = genEpcIDIP4PortKey(epcId, ip.ip4, 0);
      podServiceClusterIP4PortTable[protocol][key] = serviceId;
      podServiceClusterIP4PortTable[0][key] = serviceId;
    } else {
      for (uint32_t j = 0; j < svc.server_ports_cnt; j++) {
        uint16_t port = (uint16_t)s.u32_pool[svc.server_ports_off + j];
        podServiceClusterIP6PortTable[genEpcIDIP6PortKey(epcId, ip.ip6, 16, protocol, port)] = serviceId;
        podServiceClusterIP6PortTable[genEpcIDIP6PortKey(epcId, ip.ip6, 16, 0, port)] = serviceId;
      }
      podServiceClusterIP6PortTable[genEpcIDIP6PortKey(epcId, ip.ip6, 16, protocol, 0)] = serviceId;
      podServiceClusterIP6PortTable[genEpcIDIP6PortKey(epcId, ip.ip6, 16, 0, 0)] = serviceId;
    }
  }
}

void ServiceTable::addCustomService(const dfgpu_dict_snapshot& s, const dfgpu_service& svc) {
  int32_t epcId = (int32_t)svc.epc_id;
  uint32_t serviceId = svc.id;
  for (uint32_t i = 0; i < svc.pod_service_ids_cnt; i++)
    customServicePodService[s.u32_pool[svc.pod_service_ids_off + i]] = serviceId;
  for (uint32_t i = 0; i < svc.pod_group_ids_cnt; i++)
    customServicePodGroup[s.u32_pool[svc.pod_group_ids_off + i]] = serviceId;
  for (uint32_t i = 0; i < svc.pod_ids_cnt; i++)
    customServicePod[s.u32_pool[svc.pod_ids_off + i]] = serviceId;
  for (uint32_t i = 0; i < svc.chost_ids_cnt; i++)
    customServiceChost[s.u32_pool[svc.chost_ids_off + i]] = serviceId;
  for (uint32_t i = 0; i < svc.host_ids_cnt; i++)
    customServiceHost[s.u32_pool[svc.host_ids_off + i]] = serviceId;

  uint32_t protocol = svc.protocol;
  uint32_t podClusterId = svc.pod_cluster_id;
  if (podClusterId != 0) {  // Default_ServiceInfo_PodClusterId = 0 (trident.proto ServiceInfo)
    if (svc.server_ports_cnt == 0) {
      customServiceClusterIDTable[genPodXIDKey(podClusterId, protocol)] = serviceId;
      customServiceClusterIDPortTable[genPodXIDPortKey(podClusterId, protocol, 0)] = serviceId;
      customServiceClusterIDTable[genPodXIDKey(podClusterId, 0)] = serviceId;
      customServiceClusterIDPortTable[genPodXIDPortKey(podClusterId, 0, 0)] = serviceId;
      return;
    }
    for (uint32_t j = 0; j <= svc.server_ports_cnt; j++) {  // ports + appended 0
      uint16_t port = j < svc.server_ports_cnt ? (uint16_t)s.u32_pool[svc.server_ports_off + j] : 0;
      customServiceClusterIDPortTable[genPodXIDPortKey(podClusterId, protocol, port)] = serviceId;
      customServiceClusterIDPortTable[genPodXIDPortKey(podClusterId, 0, port)] = serviceId;
    }
    return;
  }
  if (svc.ips_cnt == 0) return;
  const dfgpu_ip& ip = s.ip_pool[svc.ips_off];
  if (!ip.valid) return;
  bool v4 = !ip.is_v6;
  if (protocol > 2) return;  // Go: index out of range panic; generator stays in range
  if (svc.server_ports_cnt == 0) {
    if (v4) {
      customServiceClusterIP4Table[genEpcIDIP4Key(epcId, ip.ip4, protocol)] = serviceId;
      customServiceClusterIP4Table[genEpcIDIP4Key(epcId, ip.ip4, 0)] = serviceId;
      customServiceClusterIP4PortTable[protocol][genEpcIDIP4PortKey(epcId, ip.ip4, 0)] = serviceId;
      customServiceClusterIP4PortTable[0][genEpcIDIP4PortKey(epcId, ip.ip4, 0)] = serviceId;
    } else {
      customServiceClusterIP6Table[genEpcIDIP6Key(epcId, ip.ip6, 16, protocol)] = serviceId;
      customServiceClusterIP6Table[genEpcIDIP6Key(epcId, ip.ip6, 16, 0)] = serviceId;
      customServiceClusterIP6PortTable[genEpcIDIP6PortKey(epcId, ip.ip6, 16, protocol, 0)] = serviceId;
      customServiceClusterIP6PortTable[genEpcIDIP6PortKey(epcId, ip.ip6, 16, 0, 0)] = serviceId;
    }
    return;
  }
  for (uint32_t j = 0; j <= svc.server_ports_cnt; j++) {
    uint16_t port = j < svc.server_ports_cnt ? (uint16_t)s.u32_pool[svc.server_ports_off + j] : 0;
    if (v4) {
      uint64_t key = genEpcIDIP4PortKey(epcId, ip.ip4, port);
      customServiceClusterIP4PortTable[protocol][key] = serviceId;
      customServiceClusterIP4PortTable[0][key] = serviceId;
    } else {
      customServiceClusterIP6PortTable[genEpcIDIP6PortKey(epcId, ip.ip6, 16, protocol, port)] = serviceId;
      customServiceClusterIP6PortTable[genEpcIDIP6PortKey(epcId, ip.ip6, 16, 0, port)] = serviceId;
    }
  }
}

uint32_t ServiceTable::QueryPodService(const PlatformInfoTable* t, uint32_t podID,
                                       uint32_t podNodeID, uint32_t podClusterID,
                                       uint32_t podGroupID, int32_t epcID, bool isIPv6,
                                       uint32_t ipv4, const uint8_t* ipv6, size_t ip6len,
                                       uint8_t protocol, uint16_t serverPort) const {
  if (serverPort == 0) protocol = 0;
  uint32_t serviceProtocol = toServiceProtocol(protocol);
  uint32_t serviceID = 0;
  if (podID != 0) {
    serviceID = getz(podServiceGroupIDPortTable, genPodXIDPortKey(podGroupID, serviceProtocol, serverPort));
    if (serviceID == 0) {
      Info info;
      bool found = isIPv6 ? t->QueryIPV6Infos(epcID, ipv6, ip6len, info)
                          : t->QueryIPV4Infos(epcID, ipv4, info);
      if (found && info.PodID != podID) return serviceID;
    }
  } else if (podNodeID != 0 && serverPort != 0) {
    serviceID = getz(podServiceClusterIDPortTable, genPodXIDPortKey(podClusterID, serviceProtocol, serverPort));
  }
  if (serviceID != 0) return serviceID;
  if (epcID <= 0) return 0;
  if (isIPv6)
    return getz(podServiceClusterIP6PortTable,
                genEpcIDIP6PortKey(epcID, ipv6, ip6len, serviceProtocol, serverPort));
  return getz(podServiceClusterIP4PortTable[serviceProtocol], genEpcIDIP4PortKey(epcID, ipv4, serverPort));
}

uint32_t ServiceTable::QueryCustomService(int32_t epcID, bool isIPv6, uint32_t ip4,
                                          const uint8_t* ip6, size_t ip6len,
                                          uint16_t serverPort, uint32_t podServiceId,
                                          uint32_t podGroupId, uint32_t l3DeviceId,
                                          uint32_t podId, uint32_t podClusterID,
                                          uint8_t l3DeviceType, uint8_t protocol) const {
  if (epcID < 0) return 0;
  if (serverPort == 0) protocol = 0;
  uint32_t sp = toServiceProtocol(protocol);
  uint32_t serviceId = 0;
  if (podServiceId != 0 && !customServicePodService.empty()) {
    uint32_t v = getz(customServicePodService, podServiceId);
    if (v != 0) return v;
  }
  if (podGroupId != 0 && !customServicePodGroup.empty()) {
    uint32_t v = getz(customServicePodGroup, podGroupId);
    if (v != 0) return v;
  }
  if (podId != 0 && !customServicePod.empty()) {
    uint32_t v = getz(customServicePod, podId);
    if (v != 0) return v;
  }
  if (podClusterID != 0 && serverPort != 0) {
    if (!customServiceClusterIDPortTable.empty()) {
      serviceId = getz(customServiceClusterIDPortTable, genPodXIDPortKey(podClusterID, sp, serverPort));
      if (serviceId > 0) return serviceId;
    }
    if (!customServiceClusterIDTable.empty()) {
      serviceId = getz(customServiceClusterIDTable, genPodXIDKey(podClusterID, 0));
      if (serviceId > 0) return serviceId;
    }
  }
  if (l3DeviceId != 0 && l3DeviceType == 1 /*VMDevice*/ && !customServiceChost.empty()) {
    uint32_t v = getz(customServiceChost, l3DeviceId);
    if (v != 0) return v;
  }
  if (l3DeviceId != 0 && l3DeviceType == 6 /*HostDevice*/ && !customServiceHost.empty()) {
    uint32_t v = getz(customServiceHost, l3DeviceId);
    if (v != 0) return v;
  }
  if (isIPv6) {
    if (!customServiceClusterIP6PortTable.empty()) {
      serviceId = getz(customServiceClusterIP6PortTable, genEpcIDIP6PortKey(epcID, ip6, ip6len, sp, serverPort));
      if (serviceId > 0) return serviceId;
    }
    if (!customServiceClusterIP6Table.empty() && serverPort != 0) {
      serviceId = getz(customServiceClusterIP6Table, genEpcIDIP6Key(epcID, ip6, ip6len, 0));
      if (serviceId > 0) return serviceId;
    }
    if (!customServiceClusterIP6PortTable.empty()) {
      serviceId = getz(customServiceClusterIP6PortTable, genEpcIDIP6PortKey(0, ip6, ip6len, sp, serverPort));
      if (serviceId > 0) return serviceId;
    }
    if (!customServiceClusterIP6Table.empty() && serverPort != 0)
      serviceId = getz(customServiceClusterIP6Table, genEpcIDIP6Key(0, ip6, ip6len, 0));
    return serviceId;
  }
  if (!customServiceClusterIP4PortTable[sp].empty()) {
    serviceId = getz(customServiceClusterIP4PortTable[sp], genEpcIDIP4PortKey(epcID, ip4, serverPort));
    if (serviceId > 0) return serviceId;
  }
  if (!customServiceClusterIP4Table.empty() && serverPort != 0) {
    serviceId = getz(customServiceClusterIP4Table, genEpcIDIP4Key(epcID, ip4, 0));
    if (serviceId > 0) return serviceId;
  }
  if (!customServiceClusterIP4PortTable[sp].empty()) {
    serviceId = getz(customServiceClusterIP4PortTable[sp], genEpcIDIP4PortKey(0, ip4, serverPort));
    if (serviceId > 0) return serviceId;
  }
  if (!customServiceClusterIP4Table.empty() && serverPort != 0)
    serviceId = getz(customServiceClusterIP4Table, genEpcIDIP4Key(0, ip4, 0));
  return serviceId;
}

// ---- PlatformInfoTable build --------------------------------------------------
void PlatformInfoTable::updateInterfaceInfos(const dfgpu_dict_snapshot& s, const dfgpu_interface& intf) {
  int32_t epcID = (int32_t)intf.epc_id;
  if (epcID == 0) epcID = EPC_FROM_INTERNET;  // :1498-1500
  if (!epcIDBaseInfos.count(epcID)) epcIDBaseInfos[epcID] = intf.region_id;  // :1513-1517
  uint32_t firstSubnetID = 0;
  bool isWan = intf.if_type == IF_TYPE_WAN;
  for (uint32_t i = 0; i < intf.ip_res_cnt; i++) {
    const dfgpu_ip_resource& r = s.ip_resources[intf.ip_res_off + i];
    uint32_t subnetID = r.subnet_id;
    if (firstSubnetID == 0) firstSubnetID = subnetID;
    if (!r.ip.is_v6 && !r.ip.valid) {
      // isIPV4(ipStr) is decided from the string; an unparsable v4 string yields
      // ip 0 in Go (utils.ParserStringIpV4 -> nil -> IpToUint32 0).  The generator
      // never emits one; treat as 0.0.0.0.
    }
    if (r.ip.is_v6 && !r.ip.valid) continue;  // net.ParseIP == nil -> continue :1569-1572
    auto info = std::make_shared<Info>();
    info->EpcID = epcID;
    info->HostID = intf.launch_server_id;
    info->Mac = intf.mac;
    info->RegionID = intf.region_id;
    info->DeviceType = intf.device_type;
    info->DeviceID = intf.device_id;
    info->SubnetID = subnetID;
    info->PodNodeID = intf.pod_node_id;
    info->PodNSID = intf.pod_ns_id;
    info->PodGroupID = intf.pod_group_id;
    info->PodGroupType = (uint8_t)intf.pod_group_type;
    info->PodID = intf.pod_id;
    info->PodClusterID = intf.pod_cluster_id;
    info->AZID = intf.az_id;
    info->IsWan = isWan;
    info->VtapID = intf.vtap_id;
    if (!r.ip.is_v6) {
      info->IsIPv4 = true;
      info->IP4 = r.ip.ip4;
      epcIDIPV4Infos[(uint64_t)(int64_t)epcID << 32 | (uint64_t)r.ip.ip4] = info;
      if (intf.pod_id > 0) podIDInfosPlatformData[intf.pod_id] = info;
      if (isWan) {
        auto w = std::make_shared<Info>();
        w->EpcID = epcID;
        epcIDIPV4Infos[(uint64_t)r.ip.ip4] = w;  // :1559-1565
      }
    } else {
      info->IsIPv4 = false;
      memcpy(info->IP6v.data(), r.ip.ip6, 16);
      epcIDIPV6Infos[make_key20(epcID, r.ip.ip6, 16)] = info;
      if (intf.pod_id > 0) podIDInfosPlatformData[intf.pod_id] = info;
      if (isWan) {
        auto w = std::make_shared<Info>();
        w->EpcID = epcID;
        epcIDIPV6Infos[make_key20(0, r.ip.ip6, 16)] = w;  // :1604-1611
      }
    }
  }
  if (intf.mac != 0) {
    uint64_t l3EpcMac = intf.mac | (uint64_t)(int64_t)epcID << 48;  // :1619
    auto it = macInfos.find(l3EpcMac);
    if (it != macInfos.end()) {
      it->second->DupCount += 1;
    } else {
      auto m = std::make_shared<Info>();
      m->EpcID = epcID;
      m->L2EpcID = epcID;
      m->DeviceType = intf.device_type;
      m->DeviceID = intf.device_id;
      m->HostID = intf.launch_server_id;
      m->Mac = intf.mac;
      m->RegionID = intf.region_id;
      m->SubnetID = firstSubnetID;
      m->PodNodeID = intf.pod_node_id;
      m->PodGroupID = intf.pod_group_id;
      m->PodGroupType = (uint8_t)intf.pod_group_type;
      m->PodID = intf.pod_id;
      m->PodClusterID = intf.pod_cluster_id;
      m->AZID = intf.az_id;
      m->VtapID = intf.vtap_id;
      macInfos[l3EpcMac] = m;  // note: PodNSID is NOT set on mac infos (:1623-1640)
    }
  }
}

void PlatformInfoTable::updateCidrInfos(const dfgpu_cidr& c) {
  if (!c.prefix.valid) return;  // ParseCIDR error -> return
  int32_t epcID = c.epc_id;
  if (epcID == 0) epcID = EPC_FROM_INTERNET;  // :1433-1435
  auto ci = std::make_shared<CidrInfo>();
  ci->v6 = c.prefix.is_v6;
  ci->plen = (int)c.prefix_len;
  ci->EpcID = epcID;
  ci->AZID = c.az_id;
  ci->RegionID = c.region_id;
  ci->SubnetID = c.subnet_id;
  ci->IsWan = c.type == 1;
  if (!ci->v6) {
    ci->mask4 = ci->plen == 0 ? 0 : 0xffffffffu << (32 - ci->plen);
    ci->net4 = c.prefix.ip4 & ci->mask4;
  } else {
    memcpy(ci->net6.data(), c.prefix.ip6, 16);
  }
  if (!epcIDBaseInfos.count(epcID)) epcIDBaseInfos[epcID] = c.region_id;  // :1446-1450
  auto& tbl = ci->v6 ? epcIDIPV6CidrInfos : epcIDIPV4CidrInfos;
  tbl[epcID].push_back(ci);
  if (ci->IsWan) tbl[0].push_back(ci);  // :1456-1462, :1468-1474
}

void PlatformInfoTable::updateGprocessInfos(const dfgpu_dict_snapshot& s) {
  for (uint64_t i = 0; i < s.n_gprocesses; i++) {
    const dfgpu_gprocess& g = s.gprocesses[i];
    if (g.pid != 0) vtapIDProcessInfos[(uint64_t)g.vtap_id << 32 | (uint64_t)g.pid] = g.gprocess_id;
    if (g.vtap_id == 0 || g.gprocess_id == 0 || g.pod_id == 0) continue;
    gprocessInfos[g.gprocess_id] = (uint64_t)g.vtap_id << 32 | (uint64_t)g.pod_id;
  }
}

void PlatformInfoTable::updateVtapIps(const dfgpu_dict_snapshot& s) {
  for (uint64_t i = 0; i < s.n_vtap_ips; i++) {
    const dfgpu_vtap_ip& v = s.vtap_ips[i];
    int32_t epcId = (int32_t)v.epc_id;
    if (epcId == 0) epcId = EPC_FROM_INTERNET;
    VtapInfo info;
    info.VtapId = (uint16_t)v.vtap_id;
    info.EpcId = epcId;
    info.IsIPv4 = true;
    if (v.ip.valid) {
      if (!v.ip.is_v6) {
        info.IP4 = v.ip.ip4;
        info.IsIPv4 = true;
      } else {
        memcpy(info.IP6v.data(), v.ip.ip6, 16);
        info.has_ip6 = true;
        info.IsIPv4 = false;
      }
    }
    vtapIdInfos[(uint16_t)v.vtap_id] = info;
  }
}

void PlatformInfoTable::updatePodIps(const dfgpu_dict_snapshot& s) {
  podIDInfos = podIDInfosPlatformData;  // :1865-1869
  for (uint64_t i = 0; i < s.n_pod_ips; i++) {
    const dfgpu_pod_ip& p = s.pod_ips[i];
    int32_t epcId = (int32_t)p.epc_id;
    if (epcId == 0) epcId = EPC_FROM_INTERNET;
    if (podIDInfos.count(p.pod_id)) continue;  // :1907
    const dfgpu_ip* ip = &p.ip;
    if (!ip->valid) ip = &p.pod_node_ip;  // ip == "" -> podNodeIp
    Info info;
    if (ip->valid) {
      Info q;
      bool found = !ip->is_v6 ? QueryIPV4Infos(epcId, ip->ip4, q) : QueryIPV6Infos(epcId, ip->ip6, 16, q);
      if (found) info = q;
      info.IsIPv4 = !ip->is_v6;
      info.IP4 = ip->is_v6 ? 0 : ip->ip4;
      if (ip->is_v6) memcpy(info.IP6v.data(), ip->ip6, 16);
    }
    info.PodID = p.pod_id;
    info.PodClusterID = p.pod_cluster_id;
    info.PodNSID = p.pod_ns_id;
    info.PodGroupID = p.pod_group_id;
    info.PodGroupType = (uint8_t)p.pod_group_type;
    info.EpcID = epcId;
    podIDInfos[p.pod_id] = std::make_shared<Info>(info);
  }
}

void PlatformInfoTable::load(const dfgpu_dict_snapshot& s) {
  *this = PlatformInfoTable();
  for (uint64_t i = 0; i < s.n_interfaces; i++) updateInterfaceInfos(s, s.interfaces[i]);
  for (uint64_t i = 0; i < s.n_cidrs; i++) updateCidrInfos(s.cidrs[i]);
  // sort by prefix length, longest first (:1476-1491). sort.Slice is unstable; the
  // generator never emits two identical CIDRs in one epc, so ties cannot both match.
  for (auto* tbl : {&epcIDIPV4CidrInfos, &epcIDIPV6CidrInfos})
    for (auto& kv : *tbl)
      std::stable_sort(kv.second.begin(), kv.second.end(),
                       [](const std::shared_ptr<CidrInfo>& a, const std::shared_ptr<CidrInfo>& b) {
                         return a->plen > b->plen;
                       });
  updateGprocessInfos(s);
  services.build(s);  // updateServices :1085
  // updateOthers :1147-1164 (region/analyzer first is irrelevant for updateVtapIps;
  // updatePodIps consults queryEpcIDBaseInfo -> regionID, which Go reads from the
  // PREVIOUS sync's t.regionID on the very first load. We model the steady state.)
  regionID = s.region_id;
  analyzerID = s.analyzer_id;
  updateVtapIps(s);
  updatePodIps(s);
}

// ---- queries -------------------------------------------------------------------
bool PlatformInfoTable::queryEpcIDBaseInfo(int32_t epcID, uint32_t& region) const {
  if (epcID == EPC_UNKNOWN) {
    region = regionID;
    return true;
  }
  auto it = epcIDBaseInfos.find(epcID);
  if (it == epcIDBaseInfos.end()) return false;
  region = it->second;
  return true;
}

bool PlatformInfoTable::queryIPV4Cidr(int32_t epcID, uint32_t ipv4, Info& out) const {
  auto it = epcIDIPV4CidrInfos.find(epcID);
  if (it == epcIDIPV4CidrInfos.end()) return false;
  for (auto& c : it->second)
    if (c->contains4(ipv4)) {
      out = Info();
      out.SubnetID = c->SubnetID;
      out.RegionID = c->RegionID;
      out.AZID = c->AZID;
      return true;
    }
  return false;
}
bool PlatformInfoTable::queryIPV6Cidr(int32_t epcID, const uint8_t* ip6, Info& out) const {
  auto it = epcIDIPV6CidrInfos.find(epcID);
  if (it == epcIDIPV6CidrInfos.end()) return false;
  for (auto& c : it->second)
    if (c->contains6(ip6)) {
      out = Info();
      out.SubnetID = c->SubnetID;
      out.RegionID = c->RegionID;
      out.AZID = c->AZID;
      return true;
    }
  return false;
}

bool PlatformInfoTable::QueryIPV4Infos(int32_t epcID, uint32_t ipv4, Info& out) const {
  if (epcID == EPC_FROM_INTERNET) return false;
  uint64_t key = (uint64_t)(int64_t)epcID << 32 | (uint64_t)ipv4;  // :581
  auto it = epcIDIPV4Infos.find(key);
  if (it != epcIDIPV4Infos.end()) {
    out = *it->second;
    return true;
  }
  if (queryIPV4Cidr(epcID, ipv4, out)) return true;
  uint32_t region;
  if (!queryEpcIDBaseInfo(epcID, region)) return false;
  out = Info();
  out.RegionID = region;
  return true;
}

bool PlatformInfoTable::QueryIPV6Infos(int32_t epcID, const uint8_t* ip6, size_t len, Info& out) const {
  if (epcID == EPC_FROM_INTERNET) return false;
  Key20 key = make_key20(epcID, ip6, len);
  auto it = epcIDIPV6Infos.find(key);
  if (it != epcIDIPV6Infos.end()) {
    out = *it->second;
    return true;
  }
  // net.IPNet.Contains on a non-16/4-byte ip returns false
  // (IPNet.Contains first reduces a v4-mapped address with To4(); its length then
  //  differs from a v6 network's and the match fails)
  uint8_t ip16[16] = {0}, v4[4];
  if (len == 16 && !to4(ip6, len, v4)) {
    memcpy(ip16, ip6, 16);
    if (queryIPV6Cidr(epcID, ip16, out)) return true;
  }
  uint32_t region;
  if (!queryEpcIDBaseInfo(epcID, region)) return false;
  out = Info();
  out.RegionID = region;
  return true;
}

bool PlatformInfoTable::QueryMacInfo(uint64_t mac, Info& out) const {
  auto it = macInfos.find(mac);
  if (it == macInfos.end()) return false;
  if (it->second->DupCount > 0) return false;
  out = *it->second;
  return true;
}
bool PlatformInfoTable::QueryPodIdInfo(uint32_t podId, Info& out) const {
  auto it = podIDInfos.find(podId);
  if (it == podIDInfos.end()) return false;
  out = *it->second;
  return true;
}
bool PlatformInfoTable::QueryVtapInfo(uint16_t vtapId, VtapInfo& out) const {
  auto it = vtapIdInfos.find(vtapId);
  if (it == vtapIdInfos.end()) return false;
  out = it->second;
  return true;
}
void PlatformInfoTable::QueryGprocessInfo(uint32_t gpid, uint16_t& vtap, uint32_t& pod) const {
  auto it = gprocessInfos.find(gpid);
  if (it == gprocessInfos.end()) {
    vtap = 0;
    pod = 0;
    return;
  }
  vtap = (uint16_t)(it->second >> 32);
  pod = (uint32_t)(it->second << 32 >> 32);
}
uint32_t PlatformInfoTable::QueryProcessInfo(uint16_t vtapId, uint32_t pid) const {
  auto it = vtapIDProcessInfos.find((uint64_t)vtapId << 32 | (uint64_t)pid);
  return it == vtapIDProcessInfos.end() ? 0 : it->second;
}

}  // namespace oracle
