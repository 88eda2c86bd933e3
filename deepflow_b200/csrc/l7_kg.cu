/*
 * l7_kg.cu — SmartEncoding of the L7 rows: KnowledgeGraph.FillL7 (flow_log/log_data/l7_flow_log.go:599-613) +
 * QueryProcessInfo (:589-596), one lane per (record, side) — kg_kernel.cuh.  Inputs come from the decode kernel
 * (l7_pass.cu), the row index from rec_row.
 */
#include "kg_kernel.cuh"
#include "l7_dev.h"

namespace dfgpu {

static_assert(L7C_REGION_ID_1 == L7C_REGION_ID_0 + 1 && L7C_AZ_ID_0 == L7C_REGION_ID_0 + 2 && L7C_HOST_ID_0 == L7C_REGION_ID_0 + 4 &&
                  L7C_L3_DEVICE_TYPE_0 == L7C_REGION_ID_0 + 6 && L7C_L3_DEVICE_ID_0 == L7C_REGION_ID_0 + 8 &&
                  L7C_POD_NODE_ID_0 == L7C_REGION_ID_0 + 10 && L7C_POD_NS_ID_0 == L7C_REGION_ID_0 + 12 &&
                  L7C_POD_GROUP_ID_0 == L7C_REGION_ID_0 + 14 && L7C_POD_ID_0 == L7C_REGION_ID_0 + 16 &&
                  L7C_POD_CLUSTER_ID_0 == L7C_REGION_ID_0 + 18 && L7C_L3_EPC_ID_0 == L7C_REGION_ID_0 + 20 &&
                  L7C_EPC_ID_0 == L7C_REGION_ID_0 + 22 && L7C_SUBNET_ID_0 == L7C_REGION_ID_0 + 24 &&
                  L7C_SERVICE_ID_0 == L7C_REGION_ID_0 + 26 && L7C_AUTO_INSTANCE_ID_0 == L7C_REGION_ID_0 + 28 &&
                  L7C_AUTO_INSTANCE_TYPE_0 == L7C_REGION_ID_0 + 29 && L7C_AUTO_SERVICE_ID_0 == L7C_REGION_ID_0 + 30 &&
                  L7C_AUTO_SERVICE_TYPE_0 == L7C_REGION_ID_0 + 31 && L7C_AUTO_INSTANCE_ID_1 == L7C_REGION_ID_0 + 32 &&
                  L7C_TAG_SOURCE_0 == L7C_REGION_ID_0 + 36 && L7C_TAG_SOURCE_1 == L7C_REGION_ID_0 + 37 &&
                  L7C_TEAM_ID == L7C_REGION_ID_0 + 38 && L7C_GPROCESS_ID_1 == L7C_GPROCESS_ID_0 + 1,
              "KG_COLUMNS order (kg_kernel.cuh)");
static_assert(KG_WORDS == L7KG_WORDS, "kgin layout");

#ifndef L7_KG_MIN_BLOCKS
#define L7_KG_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, L7_KG_MIN_BLOCKS) k_l7_kg(const __grid_constant__ BatchParams p) {
  __shared__ uint64_t s_fix[L7C_COUNT];
  for (uint32_t i = threadIdx.x; i < L7C_COUNT; i += blockDim.x) s_fix[i] = p.dyn->fix_off[i];
  __syncthreads();
  const L7Dev* dv = reinterpret_cast<const L7Dev*>(p.ext);
  KgSrc q;
  q.kgin = dv->kgin;
  q.ip6_src = dv->refs + (size_t)L7REF_IP6_SRC * dv->rec_cap;
  q.ip6_dst = dv->refs + (size_t)L7REF_IP6_DST * dv->rec_cap;
  q.rec_row = dv->rec_row;
  q.rec_cap = dv->rec_cap;
  kg_rows<L7C_REGION_ID_0, L7C_GPROCESS_ID_0>(p, q, s_fix);
}

void launch_l7_kg(const BatchParams& p, unsigned grid, cudaStream_t s) { k_l7_kg<<<grid, 256, 0, s>>>(p); }

}  // namespace dfgpu
