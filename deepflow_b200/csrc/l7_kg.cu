/*
 * l7_kg.cu — SmartEncoding of the L7 rows, one lane per (record, SIDE).
 *   KnowledgeGraph.FillL7   flow_log/log_data/l7_flow_log.go:599-613 -> kg_side.cuh kgs::fill_side
 *                           (KnowledgeGraph.fill l4_flow_log.go:536-728)
 *   QueryProcessInfo use    l7_flow_log.go:589-596
 * SmartEncoding is a chain of ~6 dependent dictionary probes per endpoint, walked at ~10 of 32
 * lanes active; it was 0.82 ms of the 1.50 ms EMIT pass (ablation, profiles/).  The two endpoints of
 * a record are independent except for the loopback "borrow the peer's pod" step, so each gets its own
 * lane: twice the chains in flight per SM, half the state per thread.  The per-side routine is
 * kg_side.cuh: register-only (no by-reference structs, no byte arrays), first probes issued together.
 * The pair exchanges (is_loopback, pod) with one shuffle.  Inputs come from the slot buffer the SIZE
 * pass wrote, the row index from the EMIT pass (rec_row).
 */
#include "l7_dev.h"
#include "kg_side.cuh"

namespace dfgpu {

namespace {
template <class T>
__device__ __forceinline__ void st(const BatchParams& p, int col, uint32_t row, T v) {
  reinterpret_cast<T*>(p.arena + p.dyn->fix_off[col])[row] = v;
}
}  // namespace

#ifndef L7_KG_MIN_BLOCKS
#define L7_KG_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, L7_KG_MIN_BLOCKS) k_l7_kg(const __grid_constant__ BatchParams p) {
  const DynState* dyn = p.dyn;
  const L7Dev* dv = reinterpret_cast<const L7Dev*>(p.ext);
  if ((dyn->stats.need_repair && !p.repair) || dyn->overflow) return;
  const uint32_t nrec = batch_nrec(p);
  const size_t cap = dv->rec_cap;
  // persistent threads: one lane per (record, side)
  const uint64_t total = (uint64_t)nrec * 2;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < total; base += (uint64_t)gridDim.x * blockDim.x) {
  const uint64_t t = base + threadIdx.x;  // the loop bound is uniform across the CTA: warps stay whole
  const uint32_t rec = (uint32_t)(t >> 1);
  const int s = (int)(t & 1);
  const bool in_range = rec < nrec;
  const uint32_t row = in_range ? dv->rec_row[rec] : 0xffffffffu;
  const bool ok = row != 0xffffffffu;  // both lanes of a pair agree
  const uint32_t* K = dv->kgin + rec;
  const uint32_t w0 = ok ? K[L7KG_VTAP_SIDE_PROTO * cap] : 0u, w1 = ok ? K[L7KG_PORT_FLAGS * cap] : 0u;
  const uint32_t vtap_id = w0 & 0xffffu;

  kgs::SideIn in;
  in.v6 = (w1 >> 16) & 1u;
  in.vip = (w1 >> (17 + s)) & 1u;
  in.l3_epc = ok ? (int32_t)K[(L7KG_EPC_SRC + s) * cap] : 0;
  in.ip.ip4 = ok ? K[(L7KG_IP_SRC + s) * cap] : 0u;
  {
    const unsigned long long r6 = ok ? reinterpret_cast<const unsigned long long*>(dv->refs)[(size_t)(L7REF_IP6_SRC + s) * cap + rec] : 0ull;
    const uint32_t n6 = (uint32_t)(r6 >> 32);
    in.ip.len6 = n6;
    uint64_t lo = 0, hi = 0;  // zero padded, like copy(key[4:], ipv6)
    const uint8_t* q = p.in + (uint32_t)r6;
    for (uint32_t i = 0; i < (n6 < 16 ? n6 : 16u); i++) {
      const uint64_t v = (uint64_t)__ldg(q + i) << ((i & 7) * 8);
      if (i < 8) lo |= v; else hi |= v;
    }
    in.ip.lo = lo;
    in.ip.hi = hi;
  }
  in.mac = ok ? (uint64_t)K[(L7KG_MAC_SRC_HI + 2 * s) * cap] << 32 | K[(L7KG_MAC_SRC_LO + 2 * s) * cap] : 0ull;
  in.gpid = ok ? K[(L7KG_GPID_0 + s) * cap] : 0u;
  in.vtap_id = (uint16_t)vtap_id;
  in.port = s ? (uint16_t)(w1 & 0xffffu) : (uint16_t)0;
  in.tap_side = (w0 >> 16) & 0xffu;
  in.protocol = (uint8_t)(w0 >> 24);
  in.server_side = s == 1;
  {  // :556-567 with the partner lane (the pair sits in adjacent lanes of one warp)
    const uint32_t my_pod = ok ? K[(L7KG_POD_0 + s) * cap] : 0u;
    const bool my_lb = ok && kgs::loopback(in.v6, in.ip);
    const uint32_t other_pod = __shfl_xor_sync(0xffffffffu, my_pod, 1);
    const bool other_lb = __shfl_xor_sync(0xffffffffu, (int)my_lb, 1) != 0;
    uint32_t pod0 = s ? other_pod : my_pod, pod1 = s ? my_pod : other_pod;
    uint8_t ts0 = 0, ts1 = 0;
    kg_peer_pods(s ? other_lb : my_lb, s ? my_lb : other_lb, pod0, pod1, ts0, ts1);
    in.pod = s ? pod1 : pod0;
    in.tag_source = s ? ts1 : ts0;
  }
  if (ok) {
  KGSide k;
  kgs::fill_side(p.dict, in, k);

  uint32_t gpid = in.gpid;
  const uint32_t pid = K[(L7KG_PID_0 + s) * cap];
  if (pid != 0 && gpid == 0) {  // l7_flow_log.go:589-596
    gpid = kgs::t64_z(p.dict, TID_VTAP_PID, (uint64_t)(uint16_t)vtap_id << 32 | (uint64_t)pid);
    k.tag_source |= TS_PROCESSID;
  }
  // the _0 / _1 columns of a pair are adjacent in the schema (auto_* come in blocks of four)
  static_assert(L7C_REGION_ID_1 == L7C_REGION_ID_0 + 1 && L7C_AZ_ID_1 == L7C_AZ_ID_0 + 1 &&
                    L7C_HOST_ID_1 == L7C_HOST_ID_0 + 1 && L7C_L3_DEVICE_TYPE_1 == L7C_L3_DEVICE_TYPE_0 + 1 &&
                    L7C_L3_DEVICE_ID_1 == L7C_L3_DEVICE_ID_0 + 1 && L7C_POD_NODE_ID_1 == L7C_POD_NODE_ID_0 + 1 &&
                    L7C_POD_NS_ID_1 == L7C_POD_NS_ID_0 + 1 && L7C_POD_GROUP_ID_1 == L7C_POD_GROUP_ID_0 + 1 &&
                    L7C_POD_ID_1 == L7C_POD_ID_0 + 1 && L7C_POD_CLUSTER_ID_1 == L7C_POD_CLUSTER_ID_0 + 1 &&
                    L7C_L3_EPC_ID_1 == L7C_L3_EPC_ID_0 + 1 && L7C_EPC_ID_1 == L7C_EPC_ID_0 + 1 &&
                    L7C_SUBNET_ID_1 == L7C_SUBNET_ID_0 + 1 && L7C_SERVICE_ID_1 == L7C_SERVICE_ID_0 + 1 &&
                    L7C_AUTO_INSTANCE_ID_1 == L7C_AUTO_INSTANCE_ID_0 + 4 &&
                    L7C_AUTO_SERVICE_TYPE_1 == L7C_AUTO_SERVICE_TYPE_0 + 4 &&
                    L7C_TAG_SOURCE_1 == L7C_TAG_SOURCE_0 + 1 && L7C_GPROCESS_ID_1 == L7C_GPROCESS_ID_0 + 1,
                "column pairs");
  st<uint16_t>(p, L7C_REGION_ID_0 + s, row, k.region_id);
  st<uint16_t>(p, L7C_AZ_ID_0 + s, row, k.az_id);
  st<uint16_t>(p, L7C_HOST_ID_0 + s, row, k.host_id);
  st<uint8_t>(p, L7C_L3_DEVICE_TYPE_0 + s, row, k.l3_device_type);
  st<uint32_t>(p, L7C_L3_DEVICE_ID_0 + s, row, k.l3_device_id);
  st<uint32_t>(p, L7C_POD_NODE_ID_0 + s, row, k.pod_node_id);
  st<uint16_t>(p, L7C_POD_NS_ID_0 + s, row, k.pod_ns_id);
  st<uint32_t>(p, L7C_POD_GROUP_ID_0 + s, row, k.pod_group_id);
  st<uint32_t>(p, L7C_POD_ID_0 + s, row, k.pod_id);
  st<uint16_t>(p, L7C_POD_CLUSTER_ID_0 + s, row, k.pod_cluster_id);
  st<int32_t>(p, L7C_L3_EPC_ID_0 + s, row, k.l3_epc_id);
  st<int32_t>(p, L7C_EPC_ID_0 + s, row, k.epc_id);
  st<uint16_t>(p, L7C_SUBNET_ID_0 + s, row, k.subnet_id);
  st<uint32_t>(p, L7C_SERVICE_ID_0 + s, row, k.service_id);
  st<uint32_t>(p, L7C_AUTO_INSTANCE_ID_0 + 4 * s, row, k.auto_instance_id);
  st<uint8_t>(p, L7C_AUTO_INSTANCE_TYPE_0 + 4 * s, row, k.auto_instance_type);
  st<uint32_t>(p, L7C_AUTO_SERVICE_ID_0 + 4 * s, row, k.auto_service_id);
  st<uint8_t>(p, L7C_AUTO_SERVICE_TYPE_0 + 4 * s, row, k.auto_service_type);
  st<uint8_t>(p, L7C_TAG_SOURCE_0 + s, row, k.tag_source);
  st<uint32_t>(p, L7C_GPROCESS_ID_0 + s, row, gpid);
  }
  }
}

void launch_l7_kg(const BatchParams& p, unsigned grid, cudaStream_t s) { k_l7_kg<<<grid, 256, 0, s>>>(p); }

}  // namespace dfgpu
