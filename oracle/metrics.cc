/*
 * ORACLE — test infrastructure only.
 * metrics.cc: placeholder until the Document path lands.
 */
#include "metrics.h"

namespace oracle {
const MetricSchema& metric_schema() {
  static MetricSchema s;
  return s;
}
void MetricsBlocks::init() {}
void MetricsBlocks::append_record(const Msg&, uint16_t, uint16_t, const PlatformInfoTable*, dfgpu_counters&) {}
int MetricsBlocks::reduce(int, uint32_t) { return DFGPU_ENOENT; }
}  // namespace oracle
