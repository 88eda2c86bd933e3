/*
 * ORACLE — test infrastructure only.
 * metrics.h: flow_metrics Document path.  Interface; see metrics.cc.
 */
#pragma once
#include "l4.h"

namespace oracle {

struct MetricSchema {
  MsgDesc document{"Document", {}};
};
const MetricSchema& metric_schema();

struct MetricsBlocks {
  Block tables[9];
  Block reduced;
  void init();
  void append_record(const Msg& doc, uint16_t orgId, uint16_t teamId, const PlatformInfoTable* pd,
                     dfgpu_counters& cnt);
  int reduce(int table, uint32_t bucket_seconds);
};

}  // namespace oracle
