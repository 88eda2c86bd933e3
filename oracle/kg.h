/*
 * ORACLE — test infrastructure only.
 * kg.h: CPU restatement of SmartEncoding tag injection
 *   KnowledgeGraph        server/ingester/flow_log/log_data/l4_flow_log.go:175-226
 *   KnowledgeGraph.fill   l4_flow_log.go:536-728
 *   helpers               server/ingester/common/common.go:137-198
 *   TagSource bits        server/libs/flow-metrics/tag.go:255-266
 */
#pragma once
#include "columns.h"
#include "platform.h"

namespace oracle {

enum TagSource : uint8_t {  // tag.go:257-266
  TS_GpId = 1, TS_PodId = 2, TS_Mac = 4, TS_EpcIP = 8, TS_Peer = 16, TS_Agent = 32, TS_ProcessId = 64
};
enum TapSideEnum : uint32_t {  // tag.go:187-205
  TAP_Rest = 0, TAP_Client = 1, TAP_Server = 2, TAP_Local = 4,
  TAP_ClientProcess = 1 | 40, TAP_ServerProcess = 2 | 40
};
const char* TAPSideString(uint32_t v, bool& ok);  // tag.go:211-233

struct KnowledgeGraph {
  uint16_t RegionID0 = 0, RegionID1 = 0, AZID0 = 0, AZID1 = 0, HostID0 = 0, HostID1 = 0;
  uint8_t L3DeviceType0 = 0, L3DeviceType1 = 0;
  uint32_t L3DeviceID0 = 0, L3DeviceID1 = 0, PodNodeID0 = 0, PodNodeID1 = 0;
  uint16_t PodNSID0 = 0, PodNSID1 = 0;
  uint32_t PodGroupID0 = 0, PodGroupID1 = 0;
  uint8_t PodGroupType0 = 0, PodGroupType1 = 0;
  uint32_t PodID0 = 0, PodID1 = 0;
  uint16_t PodClusterID0 = 0, PodClusterID1 = 0;
  int32_t L3EpcID0 = 0, L3EpcID1 = 0, EpcID0 = 0, EpcID1 = 0;
  uint16_t SubnetID0 = 0, SubnetID1 = 0;
  uint32_t ServiceID0 = 0, ServiceID1 = 0;
  uint32_t AutoInstanceID0 = 0;
  uint8_t AutoInstanceType0 = 0;
  uint32_t AutoServiceID0 = 0;
  uint8_t AutoServiceType0 = 0;
  uint32_t AutoInstanceID1 = 0;
  uint8_t AutoInstanceType1 = 0;
  uint32_t AutoServiceID1 = 0;
  uint8_t AutoServiceType1 = 0;
  uint8_t TagSource0 = 0, TagSource1 = 0;
  uint16_t OrgId = 0, TeamID = 0;

  void fill(const PlatformInfoTable* pd, bool isIPv6, bool isVipInterface0, bool isVipInterface1,
            int32_t l3EpcID0, int32_t l3EpcID1, uint32_t ip40, uint32_t ip41,
            std::string_view ip60, std::string_view ip61, uint64_t mac0, uint64_t mac1,
            uint32_t gpID0, uint32_t gpID1, uint16_t vtapId, uint32_t podId0, uint32_t podId1,
            uint16_t port, uint32_t tapSide, uint8_t protocol);

  static void add_columns(Block& b);    // KnowledgeGraphBlock.ToInput l4_column_block.go:362-403
  void append(Block& b, size_t base) const;  // AppendToColumnBlock :410-451
};

int32_t MarshalInt32WithSpecialID(int32_t v);  // flow-metrics/tag.go:560-565

// server/ingester/common/common.go:137-198 (shared with the metrics twin, metrics.cc)
void GetAutoInstance(uint32_t podID, uint32_t gpID, uint32_t podNodeID, uint32_t l3DeviceID, uint32_t subnetID,
                     uint8_t l3DeviceType, int32_t l3EpcID, uint32_t& id, uint8_t& type);
void GetAutoService(uint32_t customServiceID, uint32_t podServiceID, uint32_t podGroupID, uint32_t gpID,
                    uint32_t podClusterID, uint32_t l3DeviceID, uint32_t subnetID, uint8_t l3DeviceType,
                    uint8_t podGroupType, int32_t l3EpcID, uint32_t& id, uint8_t& type);
bool IsPodServiceIP(uint8_t deviceType, uint32_t podId, uint32_t podNodeId);
bool RegetInfoFromIP(const PlatformInfoTable* pd, bool isIPv6, const uint8_t* ip6, size_t ip6len, uint32_t ip4,
                     int32_t epcID, Info& out);

}  // namespace oracle
