"""deepflow_b200 — B200-native ingest hot path for DeepFlow (decode -> SmartEncoding
-> flow_tag / flow_metrics -> ClickHouse column pack), behind the C ABI in
include/dfgpu.h.  See DESIGN.md."""
__all__ = ["abi", "build", "native", "synth"]
