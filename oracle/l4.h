/*
 * ORACLE — test infrastructure only.  See l4.cc for the reference map.
 */
#pragma once
#include <string>
#include <unordered_map>

#include "kg.h"
#include "schema.h"

namespace oracle {

// server/libs/geo/netmask_tree.go restated (maxMask + cache map).
struct GeoTree {
  std::vector<uint8_t> maxMask;
  std::unordered_map<uint32_t, uint16_t> cache;
  std::vector<std::string> region_names;
  std::string unknown = "\xe6\x9c\xaa\xe7\x9f\xa5";  // "未知" (libs/geo/utils.go:38)
  void build(const dfgpu_geo_entry* e, uint64_t n, const char* const* names, uint32_t n_names);
  uint8_t query_region(uint32_t ip) const;
  std::string_view province(uint32_t ip) const;
};

uint64_t genID(uint32_t time, uint32_t* counter, uint32_t analyzerID);
int getFirstMask(uint32_t start, uint32_t end);  // libs/geo/netmask_tree.go:37-50

struct L4Block {
  Block b;
  size_t kg_base = 0, fi_base = 0, m_base = 0;
  void init();
  void append_record(const Msg& tagged_flow, uint16_t orgId, uint16_t teamId,
                     const PlatformInfoTable* pd, const GeoTree* geo, uint32_t* id_counter);
};

}  // namespace oracle
