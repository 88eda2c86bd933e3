/*
 * metrics_host.h — host-side state of the flow_metrics path (metrics.cu, rollup.cu).
 */
#pragma once
#include "metrics_dev.h"

namespace dfgpu {

struct MHost {
  MDev dev{};
  MDev* d_dev = nullptr;
  MLayout* d_lay = nullptr;
  MLayout* h_lay = nullptr;  // pinned
  MCand* d_cand = nullptr;
  size_t cand_cap = 0;
  TagSet as_set;
  uint32_t* d_row_rec = nullptr;
  uint8_t* d_planes = nullptr;  // MDev::tagv / rawv / strv / meta
  uint32_t max_ntag = 0;
  unsigned grid_decode = 0, grid_emit = 0;
  uint32_t epoch = 0;
  std::vector<ColSpec> cols[MT_TABLE_COUNT];
};


}  // namespace dfgpu
