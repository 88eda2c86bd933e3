/*
 * ORACLE — test infrastructure only.
 * platform.h: CPU restatement of the SmartEncoding dictionary
 *   server/libs/grpc/grpc_platformdata.go  (PlatformInfoTable: build :1090-1145,
 *     :1422-1684, :1760-1791, :1861-1942, :2008-2028; queries :263-413, :574-861,
 *     :2047-2057)
 *   server/libs/grpc/grpc_service.go       (ServiceTable :32-534)
 * Go maps become std::unordered_map; the U64LRU/U160LRU front caches
 * (:574-593) are transparent and omitted; hit counters are omitted.
 */
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../include/dfgpu.h"

namespace oracle {

constexpr int32_t EPC_FROM_DEEPFLOW = -1;  // libs/datatype/endpoint.go:28
constexpr int32_t EPC_FROM_INTERNET = -2;  // :29
constexpr int32_t EPC_UNKNOWN = 0;         // :30
constexpr uint32_t IF_TYPE_WAN = 3;        // libs/datatype/platform.go:31

using IP6 = std::array<uint8_t, 16>;

// grpc_platformdata.go:64-90 (fields the hot path reads)
struct Info {
  int32_t EpcID = 0;
  int32_t L2EpcID = 0;
  uint32_t HostID = 0;
  uint64_t Mac = 0;
  uint32_t RegionID = 0;
  uint32_t DeviceType = 0;
  uint32_t DeviceID = 0;
  uint32_t SubnetID = 0;
  uint32_t PodNodeID = 0;
  uint32_t PodNSID = 0;
  uint32_t PodGroupID = 0;
  uint8_t PodGroupType = 0;
  uint32_t PodID = 0;
  uint32_t PodClusterID = 0;
  uint32_t AZID = 0;
  bool IsWan = false;
  uint32_t VtapID = 0;
  bool IsIPv4 = false;
  uint32_t IP4 = 0;
  IP6 IP6v{};
  uint32_t DupCount = 0;
};

struct CidrInfo {  // :92-100
  bool v6 = false;
  uint32_t net4 = 0, mask4 = 0;
  IP6 net6{};
  int plen = 0;
  int32_t EpcID = 0;
  uint32_t RegionID = 0, SubnetID = 0, AZID = 0;
  bool IsWan = false;
  bool contains4(uint32_t ip) const { return (ip & mask4) == net4; }
  bool contains6(const uint8_t* ip) const {
    int full = plen / 8, rem = plen % 8;
    if (memcmp(ip, net6.data(), full) != 0) return false;
    if (rem) {
      uint8_t m = (uint8_t)(0xff << (8 - rem));
      if ((ip[full] & m) != (net6[full] & m)) return false;
    }
    return true;
  }
};

struct VtapInfo {  // :115-125
  uint16_t VtapId = 0;
  int32_t EpcId = 0;
  bool IsIPv4 = true;
  uint32_t IP4 = 0;
  IP6 IP6v{};
  bool has_ip6 = false;
};

struct Key20 {  // [EpcIDIPV6_LEN]byte: epc LE u32 + 16 bytes
  std::array<uint8_t, 20> b{};
  bool operator==(const Key20& o) const { return b == o.b; }
};
struct Key20Hash {
  size_t operator()(const Key20& k) const {
    uint64_t h = 1469598103934665603ull;
    for (auto c : k.b) h = (h ^ c) * 1099511628211ull;
    return (size_t)h;
  }
};
inline Key20 make_key20(int32_t epc, const uint8_t* ip, size_t iplen) {
  Key20 k;
  uint32_t e = (uint32_t)epc;
  memcpy(k.b.data(), &e, 4);  // binary.LittleEndian.PutUint32
  memcpy(k.b.data() + 4, ip, iplen > 16 ? 16 : iplen);  // copy(key[4:], ipv6)
  return k;
}

// grpc_service.go:51-63
struct EpcIDIP6PortKey {
  IP6 ip{};
  int32_t epc = 0;
  uint16_t port = 0;
  uint8_t protocol = 0;
  bool operator<(const EpcIDIP6PortKey& o) const {
    return std::tie(ip, epc, port, protocol) < std::tie(o.ip, o.epc, o.port, o.protocol);
  }
};
struct EpcIDIP6Key {
  IP6 ip{};
  int32_t epc = 0;
  uint8_t protocol = 0;
  bool operator<(const EpcIDIP6Key& o) const {
    return std::tie(ip, epc, protocol) < std::tie(o.ip, o.epc, o.protocol);
  }
};

class PlatformInfoTable;

class ServiceTable {  // grpc_service.go:32-49
 public:
  std::unordered_map<uint64_t, uint32_t> podServiceClusterIP4PortTable[3];
  std::map<EpcIDIP6PortKey, uint32_t> podServiceClusterIP6PortTable;
  std::unordered_map<uint64_t, uint32_t> podServiceClusterIDPortTable;
  std::unordered_map<uint64_t, uint32_t> podServiceGroupIDPortTable;
  std::unordered_map<uint64_t, uint32_t> customServiceClusterIP4PortTable[3];
  std::map<EpcIDIP6PortKey, uint32_t> customServiceClusterIP6PortTable;
  std::unordered_map<uint64_t, uint32_t> customServiceClusterIP4Table;
  std::map<EpcIDIP6Key, uint32_t> customServiceClusterIP6Table;
  std::unordered_map<uint64_t, uint32_t> customServiceClusterIDPortTable;
  std::unordered_map<uint64_t, uint32_t> customServiceClusterIDTable;
  std::unordered_map<uint32_t, uint32_t> customServicePodService, customServicePodGroup,
      customServicePod, customServiceChost, customServiceHost;

  void build(const dfgpu_dict_snapshot& s);  // NewServiceTable :325-386
  uint32_t QueryPodService(const PlatformInfoTable* t, uint32_t podID, uint32_t podNodeID,
                           uint32_t podClusterID, uint32_t podGroupID, int32_t epcID,
                           bool isIPv6, uint32_t ipv4, const uint8_t* ipv6, size_t ip6len,
                           uint8_t protocol, uint16_t serverPort) const;  // :138-193
  uint32_t QueryCustomService(int32_t epcID, bool isIPv6, uint32_t ip4, const uint8_t* ip6,
                              size_t ip6len, uint16_t serverPort, uint32_t podServiceId,
                              uint32_t podGroupId, uint32_t l3DeviceId, uint32_t podId,
                              uint32_t podClusterID, uint8_t l3DeviceType,
                              uint8_t protocol) const;  // :195-323

 private:
  void addPodServiceIp(const dfgpu_dict_snapshot& s, const dfgpu_service& svc);  // :388-433
  void addCustomService(const dfgpu_dict_snapshot& s, const dfgpu_service& svc); // :435-534
};

class PlatformInfoTable {
 public:
  uint32_t regionID = 0;
  uint32_t analyzerID = 0;
  std::unordered_map<uint64_t, std::shared_ptr<Info>> epcIDIPV4Infos;
  std::unordered_map<Key20, std::shared_ptr<Info>, Key20Hash> epcIDIPV6Infos;
  std::unordered_map<int32_t, std::vector<std::shared_ptr<CidrInfo>>> epcIDIPV4CidrInfos;
  std::unordered_map<int32_t, std::vector<std::shared_ptr<CidrInfo>>> epcIDIPV6CidrInfos;
  std::unordered_map<uint64_t, std::shared_ptr<Info>> macInfos;
  std::unordered_map<int32_t, uint32_t> epcIDBaseInfos;  // BaseInfo.RegionID
  std::unordered_map<uint32_t, uint64_t> gprocessInfos;
  std::unordered_map<uint64_t, uint32_t> vtapIDProcessInfos;
  std::unordered_map<uint32_t, std::shared_ptr<Info>> podIDInfosPlatformData;
  std::unordered_map<uint32_t, std::shared_ptr<Info>> podIDInfos;
  std::unordered_map<uint16_t, VtapInfo> vtapIdInfos;
  ServiceTable services;

  // updatePlatformData :1090-1145 + updateServices + updateOthers :1147-1164
  void load(const dfgpu_dict_snapshot& s);

  uint32_t QueryAnalyzerID() const { return analyzerID; }
  // Returned Info is by value with a found flag (Go returns *Info / nil).
  bool QueryIPV4Infos(int32_t epcID, uint32_t ipv4, Info& out) const;               // :292-308
  bool QueryIPV6Infos(int32_t epcID, const uint8_t* ip6, size_t len, Info& out) const;  // :310-326
  bool QueryMacInfo(uint64_t mac, Info& out) const;                                   // :613-623
  bool QueryPodIdInfo(uint32_t podId, Info& out) const;                               // :376-382
  bool QueryVtapInfo(uint16_t vtapId, VtapInfo& out) const;                           // :1693-1698
  void QueryGprocessInfo(uint32_t gpid, uint16_t& vtap, uint32_t& pod) const;         // :2047-2052
  uint32_t QueryProcessInfo(uint16_t vtapId, uint32_t pid) const;                     // :2055-2057

 private:
  bool queryEpcIDBaseInfo(int32_t epcID, uint32_t& region) const;  // :686-700
  bool queryIPV4Cidr(int32_t epcID, uint32_t ipv4, Info& out) const;  // :821-840
  bool queryIPV6Cidr(int32_t epcID, const uint8_t* ip6, Info& out) const;  // :843-861
  void updateInterfaceInfos(const dfgpu_dict_snapshot& s, const dfgpu_interface& intf);
  void updateCidrInfos(const dfgpu_cidr& c);
  void updateGprocessInfos(const dfgpu_dict_snapshot& s);
  void updateVtapIps(const dfgpu_dict_snapshot& s);
  void updatePodIps(const dfgpu_dict_snapshot& s);
};

// server/libs/utils/utils.go:116-130 + Go net.IP predicates
bool IsLoopback(bool isIPv6, uint32_t ip4, const uint8_t* ip6, size_t ip6len);
bool IsLocalIP(bool isIPv6, uint32_t ip4, const uint8_t* ip6, size_t ip6len);

}  // namespace oracle
