/*
 * ORACLE — test infrastructure only.  See kg.h for the reference map.
 * Parity unpinned by the reference (KnowledgeGraph.fill has no reference test); cross-checked by the second
 * restatement tests/kg_second_opinion.py (tests/test_kg_crosscheck_cpu.py).  Pinned pieces: MarshalInt32WithSpecialID
 * (tag_test.go:45-56), TAP-side strings (tag.go:211-233), auto instance / service type codes (trident.proto enums).
 */
#include "kg.h"

namespace oracle {

int32_t MarshalInt32WithSpecialID(int32_t v) {  // tag.go:560-565
  if (v > 0 || v == -1 || v == -2) return v;
  return (int32_t)((uint64_t)(int64_t)v & 0xffff);
}

const char* TAPSideString(uint32_t v, bool& ok) {  // tag.go:211-233 (sparse slice literal)
  ok = true;
  switch (v) {
    case 0: return "rest";
    case 1: return "c";
    case 2: return "s";
    case 4: return "local";
    case 9: return "c-nd";
    case 10: return "s-nd";
    case 17: return "c-hv";
    case 18: return "s-hv";
    case 25: return "c-gw-hv";
    case 26: return "s-gw-hv";
    case 33: return "c-gw";
    case 34: return "s-gw";
    case 41: return "c-p";
    case 42: return "s-p";
    case 49: return "c-app";
    case 50: return "s-app";
    case 48: return "app";
  }
  ok = v <= 50;  // inside the slice but unnamed -> ""; beyond it Go panics
  return "";
}

// common.go:160-198
void GetAutoInstance(uint32_t podID, uint32_t gpID, uint32_t podNodeID, uint32_t l3DeviceID,
                            uint32_t subnetID, uint8_t l3DeviceType, int32_t l3EpcID, uint32_t& id,
                            uint8_t& type) {
  if (podID > 0) { id = podID; type = 10; }
  else if (gpID > 0) { id = gpID; type = 120; }
  else if (podNodeID > 0) { id = podNodeID; type = 14; }
  else if (l3DeviceID > 0) { id = l3DeviceID; type = l3DeviceType; }
  else if (l3EpcID == EPC_FROM_INTERNET) { id = 0; type = 0; }
  else { id = subnetID; type = 255; }
}
void GetAutoService(uint32_t customServiceID, uint32_t podServiceID, uint32_t podGroupID,
                           uint32_t gpID, uint32_t podClusterID, uint32_t l3DeviceID,
                           uint32_t subnetID, uint8_t l3DeviceType, uint8_t podGroupType,
                           int32_t l3EpcID, uint32_t& id, uint8_t& type) {
  if (customServiceID > 0) { id = customServiceID; type = 104; }
  else if (podServiceID > 0) { id = podServiceID; type = 11; /* AUTO_SERVICE_TYPE_POD_SERVICE (trident.proto:363; the "12" in the Go comment is stale) */ }
  else if (podGroupID > 0) { id = podGroupID; type = podGroupType; }
  else if (gpID > 0) { id = gpID; type = 120; }
  else if (podClusterID > 0) { id = podClusterID; type = 103; }
  else if (l3DeviceID > 0) { id = l3DeviceID; type = l3DeviceType; }
  else if (l3EpcID == EPC_FROM_INTERNET) { id = 0; type = 0; }
  else { id = subnetID; type = 255; }
}
bool IsPodServiceIP(uint8_t deviceType, uint32_t podId, uint32_t podNodeId) {
  return deviceType == 11 /*trident DEVICE_TYPE_POD_SERVICE*/ || podId != 0 || podNodeId != 0;
}
// common.go:137-143
bool RegetInfoFromIP(const PlatformInfoTable* pd, bool isIPv6, const uint8_t* ip6, size_t ip6len,
                            uint32_t ip4, int32_t epcID, Info& out) {
  return isIPv6 ? pd->QueryIPV6Infos(epcID, ip6, ip6len, out) : pd->QueryIPV4Infos(epcID, ip4, out);
}

void KnowledgeGraph::fill(const PlatformInfoTable* pd, bool isIPv6, bool isVipInterface0,
                          bool isVipInterface1, int32_t l3EpcID0, int32_t l3EpcID1, uint32_t ip40,
                          uint32_t ip41, std::string_view ip60sv, std::string_view ip61sv,
                          uint64_t mac0, uint64_t mac1, uint32_t gpID0, uint32_t gpID1,
                          uint16_t vtapId, uint32_t podId0, uint32_t podId1, uint16_t port,
                          uint32_t tapSide, uint8_t protocol) {
  const uint8_t* ip60 = (const uint8_t*)ip60sv.data();
  const uint8_t* ip61 = (const uint8_t*)ip61sv.data();
  size_t l60 = ip60sv.size(), l61 = ip61sv.size();
  Info info0, info1, agentInfo;
  bool has0 = false, has1 = false, hasAgent = false;
  bool lookupByMac0 = isVipInterface0, lookupByMac1 = isVipInterface1;
  bool lookupByAgent0 = false, lookupByAgent1 = false;

  if (IsLoopback(isIPv6, ip40, ip60, l60)) {  // :556-561
    if (podId0 == 0 && podId1 != 0) {
      TagSource0 |= TS_Peer;
      podId0 = podId1;
    }
  }
  if (IsLoopback(isIPv6, ip41, ip61, l61)) {  // :562-567
    if (podId1 == 0 && podId0 != 0) {
      TagSource1 |= TS_Peer;
      podId1 = podId0;
    }
  }
  if (tapSide == TAP_Local) {  // :570-585
    if (IsLocalIP(isIPv6, ip40, ip60, l60)) {
      if (mac0 != 0) lookupByMac0 = true; else lookupByAgent0 = true;
    }
    if (IsLocalIP(isIPv6, ip41, ip61, l61)) {
      if (mac1 != 0) lookupByMac1 = true; else lookupByAgent1 = true;
    }
  } else if (tapSide == TAP_ClientProcess || tapSide == TAP_ServerProcess) {  // :586-598
    if (mac0 != 0) lookupByMac0 = true;
    else if (IsLocalIP(isIPv6, ip40, ip60, l60)) lookupByAgent0 = true;
    if (mac1 != 0) lookupByMac1 = true;
    else if (IsLocalIP(isIPv6, ip41, ip61, l61)) lookupByAgent1 = true;
  }
  uint64_t l3EpcMac0 = mac0 | (uint64_t)(int64_t)l3EpcID0 << 48;  // :599
  uint64_t l3EpcMac1 = mac1 | (uint64_t)(int64_t)l3EpcID1 << 48;

  if (gpID0 != 0 && podId0 == 0) {  // :601-607
    uint16_t v; uint32_t p;
    pd->QueryGprocessInfo(gpID0, v, p);
    if (p != 0 && v == vtapId) { podId0 = p; TagSource0 |= TS_GpId; }
  }
  if (gpID1 != 0 && podId1 == 0) {  // :608-614
    uint16_t v; uint32_t p;
    pd->QueryGprocessInfo(gpID1, v, p);
    if (p != 0 && v == vtapId) { podId1 = p; TagSource1 |= TS_GpId; }
  }
  if (podId0 != 0) {  // :617-620
    TagSource0 |= TS_PodId;
    has0 = pd->QueryPodIdInfo(podId0, info0);
  }
  if (podId1 != 0) {
    TagSource1 |= TS_PodId;
    has1 = pd->QueryPodIdInfo(podId1, info1);
  }
  if (!has0) {  // :626-643
    if (lookupByMac0) {
      TagSource0 |= TS_Mac;
      has0 = pd->QueryMacInfo(l3EpcMac0, info0);
    }
    if (!has0 && lookupByAgent0) {
      TagSource0 |= TS_Agent;
      VtapInfo vi;
      if (pd->QueryVtapInfo(vtapId, vi)) {
        hasAgent = RegetInfoFromIP(pd, !vi.IsIPv4, vi.IP6v.data(), vi.has_ip6 ? 16 : 0, vi.IP4, vi.EpcId, agentInfo);
        if (hasAgent) { info0 = agentInfo; has0 = true; }
      }
    }
    if (!has0) {
      TagSource0 |= TS_EpcIP;
      has0 = RegetInfoFromIP(pd, isIPv6, ip60, l60, ip40, l3EpcID0, info0);
    }
  }
  if (!has1) {  // :645-666
    if (lookupByMac1) {
      TagSource1 |= TS_Mac;
      has1 = pd->QueryMacInfo(l3EpcMac1, info1);
    }
    if (!has1 && lookupByAgent1) {
      TagSource1 |= TS_Agent;
      if (lookupByAgent0 && hasAgent) {
        info1 = agentInfo;
        has1 = true;
      } else {
        VtapInfo vi;
        if (pd->QueryVtapInfo(vtapId, vi))
          has1 = RegetInfoFromIP(pd, !vi.IsIPv4, vi.IP6v.data(), vi.has_ip6 ? 16 : 0, vi.IP4, vi.EpcId, info1);
      }
    }
    if (!has1) {
      TagSource1 |= TS_EpcIP;
      has1 = RegetInfoFromIP(pd, isIPv6, ip61, l61, ip41, l3EpcID1, info1);
    }
  }
  Info l2Info0, l2Info1;
  bool hasL20 = false, hasL21 = false;
  if (l3EpcID0 > 0 && l3EpcID1 > 0) {  // :668-675 (pair lookup == two single lookups)
    hasL20 = pd->QueryMacInfo(l3EpcMac0, l2Info0);
    hasL21 = pd->QueryMacInfo(l3EpcMac1, l2Info1);
  } else if (l3EpcID0 > 0) {
    hasL20 = pd->QueryMacInfo(l3EpcMac0, l2Info0);
  } else if (l3EpcID1 > 0) {
    hasL21 = pd->QueryMacInfo(l3EpcMac1, l2Info1);
  }
  if (has0) {  // :677-690
    RegionID0 = (uint16_t)info0.RegionID;
    AZID0 = (uint16_t)info0.AZID;
    HostID0 = (uint16_t)info0.HostID;
    L3DeviceType0 = (uint8_t)info0.DeviceType;
    L3DeviceID0 = info0.DeviceID;
    PodNodeID0 = info0.PodNodeID;
    PodNSID0 = (uint16_t)info0.PodNSID;
    PodGroupID0 = info0.PodGroupID;
    PodGroupType0 = info0.PodGroupType;
    PodID0 = info0.PodID;
    PodClusterID0 = (uint16_t)info0.PodClusterID;
    SubnetID0 = (uint16_t)info0.SubnetID;
  }
  if (has1) {  // :691-704
    RegionID1 = (uint16_t)info1.RegionID;
    AZID1 = (uint16_t)info1.AZID;
    HostID1 = (uint16_t)info1.HostID;
    L3DeviceType1 = (uint8_t)info1.DeviceType;
    L3DeviceID1 = info1.DeviceID;
    PodNodeID1 = info1.PodNodeID;
    PodNSID1 = (uint16_t)info1.PodNSID;
    PodGroupID1 = info1.PodGroupID;
    PodGroupType1 = info1.PodGroupType;
    PodID1 = info1.PodID;
    PodClusterID1 = (uint16_t)info1.PodClusterID;
    SubnetID1 = (uint16_t)info1.SubnetID;
  }
  L3EpcID0 = l3EpcID0;
  L3EpcID1 = l3EpcID1;
  if (hasL20) EpcID0 = l2Info0.L2EpcID;
  if (hasL21) EpcID1 = l2Info1.L2EpcID;

  if (IsPodServiceIP(L3DeviceType0, PodID0, 0))  // :714-716
    ServiceID0 = pd->services.QueryPodService(pd, PodID0, PodNodeID0, PodClusterID0, PodGroupID0,
                                              l3EpcID0, isIPv6, ip40, ip60, l60, protocol, 0);
  if (IsPodServiceIP(L3DeviceType1, PodID1, PodNodeID1))  // :717-719
    ServiceID1 = pd->services.QueryPodService(pd, PodID1, PodNodeID1, PodClusterID1, PodGroupID1,
                                              l3EpcID1, isIPv6, ip41, ip61, l61, protocol, port);

  GetAutoInstance(PodID0, gpID0, PodNodeID0, L3DeviceID0, SubnetID0, L3DeviceType0, L3EpcID0,
                  AutoInstanceID0, AutoInstanceType0);
  uint32_t custom0 = pd->services.QueryCustomService(l3EpcID0, isIPv6, ip40, ip60, l60, 0, ServiceID0,
                                                     PodGroupID0, L3DeviceID0, PodID0, PodClusterID0,
                                                     L3DeviceType0, protocol);
  GetAutoService(custom0, ServiceID0, PodGroupID0, gpID0, PodClusterID0, L3DeviceID0, SubnetID0,
                 L3DeviceType0, PodGroupType0, L3EpcID0, AutoServiceID0, AutoServiceType0);
  GetAutoInstance(PodID1, gpID1, PodNodeID1, L3DeviceID1, SubnetID1, L3DeviceType1, L3EpcID1,
                  AutoInstanceID1, AutoInstanceType1);
  uint32_t custom1 = pd->services.QueryCustomService(l3EpcID1, isIPv6, ip41, ip61, l61, port, ServiceID1,
                                                     PodGroupID1, L3DeviceID1, PodID1, PodClusterID1,
                                                     L3DeviceType1, protocol);
  GetAutoService(custom1, ServiceID1, PodGroupID1, gpID1, PodClusterID1, L3DeviceID1, SubnetID1,
                 L3DeviceType1, PodGroupType1, L3EpcID1, AutoServiceID1, AutoServiceType1);
}

void KnowledgeGraph::add_columns(Block& b) {
  struct { const char* n; int t; int es; } c[] = {
      {"region_id_0", DFGPU_CH_UINT16, 2}, {"region_id_1", DFGPU_CH_UINT16, 2},
      {"az_id_0", DFGPU_CH_UINT16, 2}, {"az_id_1", DFGPU_CH_UINT16, 2},
      {"host_id_0", DFGPU_CH_UINT16, 2}, {"host_id_1", DFGPU_CH_UINT16, 2},
      {"l3_device_type_0", DFGPU_CH_UINT8, 1}, {"l3_device_type_1", DFGPU_CH_UINT8, 1},
      {"l3_device_id_0", DFGPU_CH_UINT32, 4}, {"l3_device_id_1", DFGPU_CH_UINT32, 4},
      {"pod_node_id_0", DFGPU_CH_UINT32, 4}, {"pod_node_id_1", DFGPU_CH_UINT32, 4},
      {"pod_ns_id_0", DFGPU_CH_UINT16, 2}, {"pod_ns_id_1", DFGPU_CH_UINT16, 2},
      {"pod_group_id_0", DFGPU_CH_UINT32, 4}, {"pod_group_id_1", DFGPU_CH_UINT32, 4},
      {"pod_id_0", DFGPU_CH_UINT32, 4}, {"pod_id_1", DFGPU_CH_UINT32, 4},
      {"pod_cluster_id_0", DFGPU_CH_UINT16, 2}, {"pod_cluster_id_1", DFGPU_CH_UINT16, 2},
      {"l3_epc_id_0", DFGPU_CH_INT32, 4}, {"l3_epc_id_1", DFGPU_CH_INT32, 4},
      {"epc_id_0", DFGPU_CH_INT32, 4}, {"epc_id_1", DFGPU_CH_INT32, 4},
      {"subnet_id_0", DFGPU_CH_UINT16, 2}, {"subnet_id_1", DFGPU_CH_UINT16, 2},
      {"service_id_0", DFGPU_CH_UINT32, 4}, {"service_id_1", DFGPU_CH_UINT32, 4},
      {"auto_instance_id_0", DFGPU_CH_UINT32, 4}, {"auto_instance_type_0", DFGPU_CH_UINT8, 1},
      {"auto_service_id_0", DFGPU_CH_UINT32, 4}, {"auto_service_type_0", DFGPU_CH_UINT8, 1},
      {"auto_instance_id_1", DFGPU_CH_UINT32, 4}, {"auto_instance_type_1", DFGPU_CH_UINT8, 1},
      {"auto_service_id_1", DFGPU_CH_UINT32, 4}, {"auto_service_type_1", DFGPU_CH_UINT8, 1},
      {"tag_source_0", DFGPU_CH_UINT8, 1}, {"tag_source_1", DFGPU_CH_UINT8, 1},
      {"team_id", DFGPU_CH_UINT16, 2}};
  for (auto& x : c) b.add(x.n, x.t, x.es);
}

void KnowledgeGraph::append(Block& b, size_t i) const {
  b[i++].append<uint16_t>(RegionID0); b[i++].append<uint16_t>(RegionID1);
  b[i++].append<uint16_t>(AZID0); b[i++].append<uint16_t>(AZID1);
  b[i++].append<uint16_t>(HostID0); b[i++].append<uint16_t>(HostID1);
  b[i++].append<uint8_t>(L3DeviceType0); b[i++].append<uint8_t>(L3DeviceType1);
  b[i++].append<uint32_t>(L3DeviceID0); b[i++].append<uint32_t>(L3DeviceID1);
  b[i++].append<uint32_t>(PodNodeID0); b[i++].append<uint32_t>(PodNodeID1);
  b[i++].append<uint16_t>(PodNSID0); b[i++].append<uint16_t>(PodNSID1);
  b[i++].append<uint32_t>(PodGroupID0); b[i++].append<uint32_t>(PodGroupID1);
  b[i++].append<uint32_t>(PodID0); b[i++].append<uint32_t>(PodID1);
  b[i++].append<uint16_t>(PodClusterID0); b[i++].append<uint16_t>(PodClusterID1);
  b[i++].append<int32_t>(L3EpcID0); b[i++].append<int32_t>(L3EpcID1);
  b[i++].append<int32_t>(EpcID0); b[i++].append<int32_t>(EpcID1);
  b[i++].append<uint16_t>(SubnetID0); b[i++].append<uint16_t>(SubnetID1);
  b[i++].append<uint32_t>(ServiceID0); b[i++].append<uint32_t>(ServiceID1);
  b[i++].append<uint32_t>(AutoInstanceID0); b[i++].append<uint8_t>(AutoInstanceType0);
  b[i++].append<uint32_t>(AutoServiceID0); b[i++].append<uint8_t>(AutoServiceType0);
  b[i++].append<uint32_t>(AutoInstanceID1); b[i++].append<uint8_t>(AutoInstanceType1);
  b[i++].append<uint32_t>(AutoServiceID1); b[i++].append<uint8_t>(AutoServiceType1);
  b[i++].append<uint8_t>(TagSource0); b[i++].append<uint8_t>(TagSource1);
  b[i++].append<uint16_t>(TeamID);
}

}  // namespace oracle
