/*
 * ORACLE — test infrastructure only.  See l7.cc for the reference map.
 */
#pragma once
#include <list>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>

#include "l4.h"

namespace oracle {

// server/libs/lru/lru.go:23-76 (map + container/list), Key -> uint32 value
template <class Key, class Hash>
class LruCache {
 public:
  explicit LruCache(int cap = 1 << 18) : capacity(cap) {}
  // AddOrGet :44-60 — returns {value, found}
  std::pair<uint32_t, bool> AddOrGet(const Key& k, uint32_t v) {
    auto it = cache.find(k);
    if (it != cache.end()) {
      lru.splice(lru.begin(), lru, it->second);
      return {it->second->second, true};
    }
    lru.emplace_front(k, v);
    cache[k] = lru.begin();
    if ((int)lru.size() > capacity) removeOldest();
    return {v, false};
  }
  void Add(const Key& k, uint32_t v) {  // :62-76
    auto it = cache.find(k);
    if (it != cache.end()) {
      lru.splice(lru.begin(), lru, it->second);
      it->second->second = v;
      return;
    }
    lru.emplace_front(k, v);
    cache[k] = lru.begin();
    if ((int)lru.size() > capacity) removeOldest();
  }
  size_t Len() const { return lru.size(); }
  int Keys(Key* out) const {  // oldest -> newest (lru.go:105-113)
    int i = 0;
    for (auto it = lru.rbegin(); it != lru.rend(); ++it) out[i++] = it->first;
    return i;
  }

 private:
  void removeOldest() {
    auto last = std::prev(lru.end());
    cache.erase(last->first);
    lru.pop_back();
  }
  int capacity;
  std::list<std::pair<Key, uint32_t>> lru;
  std::unordered_map<Key, typename std::list<std::pair<Key, uint32_t>>::iterator, Hash> cache;
};

// flow_tag.FlowTagInfo (flow_tag.go:93-113) — only the members this path sets
struct FlowTagInfo {
  std::string Table, FieldName, FieldValue;
  int32_t VpcId = 0;
  uint16_t PodNsId = 0;
  uint8_t FieldType = 0;  // 0 tag, 1 metrics, 2 custom_tag
  uint16_t OrgId = 0, TeamID = 0;
  bool operator==(const FlowTagInfo& o) const {
    return VpcId == o.VpcId && PodNsId == o.PodNsId && FieldType == o.FieldType && OrgId == o.OrgId &&
           TeamID == o.TeamID && FieldName == o.FieldName && FieldValue == o.FieldValue && Table == o.Table;
  }
};
inline uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
  const uint8_t* b = (const uint8_t*)p;
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}
struct FlowTagInfoHash {  // stands in for Go's runtime map hash of the struct key
  size_t operator()(const FlowTagInfo& k) const {
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(h, k.FieldName.data(), k.FieldName.size());
    h = fnv1a(h, "\0", 1);
    h = fnv1a(h, k.FieldValue.data(), k.FieldValue.size());
    uint64_t t = (uint64_t)(uint32_t)k.VpcId << 32 | (uint64_t)k.PodNsId << 16 | (uint64_t)k.FieldType << 8;
    h = fnv1a(h, &t, 8);
    uint32_t u = (uint32_t)k.OrgId << 16 | k.TeamID;
    h = fnv1a(h, &u, 4);
    return (size_t)h;
  }
};
struct AppServiceTagKey {  // app_service_tag.go: Table, AppService, AppInstance, OrgId, TeamID
  std::string Table, AppService, AppInstance;
  uint16_t OrgId = 0, TeamID = 0;
  bool operator==(const AppServiceTagKey& o) const {
    return OrgId == o.OrgId && TeamID == o.TeamID && AppService == o.AppService &&
           AppInstance == o.AppInstance && Table == o.Table;
  }
};
struct AppServiceTagKeyHash {
  size_t operator()(const AppServiceTagKey& k) const {
    uint64_t h = 1469598103934665603ull;
    h = fnv1a(h, k.AppService.data(), k.AppService.size());
    h = fnv1a(h, "\0", 1);
    h = fnv1a(h, k.AppInstance.data(), k.AppInstance.size());
    uint32_t u = (uint32_t)k.OrgId << 16 | k.TeamID;
    h = fnv1a(h, &u, 4);
    return (size_t)h;
  }
};

struct L7Config {
  bool trace_id_index_disabled = false;
  uint32_t flow_tag_cache_flush_timeout = 1800;
  uint32_t flow_tag_cache_max_size = 1 << 18;
};

struct L7Counters {
  int64_t http = 0, dns = 0, sql = 0, nosql = 0, rpc = 0, mq = 0;
  int64_t new_field = 0, new_field_value = 0;
};

struct L7State {
  L7Config cfg;
  LruCache<FlowTagInfo, FlowTagInfoHash> FieldCache, FieldValueCache;  // flow_tag_writer.go:65-77
  LruCache<AppServiceTagKey, AppServiceTagKeyHash> AppServiceCache;         // app_service_tag_writer.go:52
  explicit L7State(const L7Config& c)
      : cfg(c), FieldCache((int)c.flow_tag_cache_max_size),
        FieldValueCache((int)c.flow_tag_cache_max_size), AppServiceCache((int)c.flow_tag_cache_max_size) {}
};

struct L7Blocks {
  Block l7;          // flow_log.l7_flow_log
  Block field;       // flow_tag.l7_flow_log_custom_field
  Block field_value; // flow_tag.l7_flow_log_custom_field_value
  Block app_service; // flow_tag.l7_flow_log_app_service
  void init();
  // sendProto (decoder.go:474-496) for one decoded AppProtoLogsData
  void append_record(const Msg& m, uint16_t orgId, uint16_t teamId, const PlatformInfoTable* pd,
                     L7State& st, uint32_t* id_counter, uint64_t* lastTraceIdIndex, L7Counters& cnt);
};

uint64_t DJBHash(uint64_t base, std::string_view s);  // libs/utils/hash.go:37-43
std::string_view L7ProtocolString(uint8_t p, bool isTLS);  // datatype/flow.go:623-803

}  // namespace oracle
