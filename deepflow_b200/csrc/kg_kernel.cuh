/*
 * kg_kernel.cuh — SmartEncoding (KnowledgeGraph.fill, flow_log/log_data/l4_flow_log.go:536-728) of the rows of
 * one flow_log batch, one lane per (record, SIDE): the body shared by the L4 (FillL4 :730-743) and the L7
 * (FillL7 l7_flow_log.go:599-613, QueryProcessInfo :589-596) kernels.  The decode kernel of the message type left
 * the inputs as coalesced word arrays (KG_* words, [word][rec_cap]); kg_side.cuh does the look-ups.
 * The two endpoints of a record are independent except for the loopback "borrow the peer's pod" step, which the
 * pair of adjacent lanes settles with one shuffle.
 */
#pragma once
#ifndef KG_SIDES_IN_SEQUENCE
#define KG_SIDES_IN_SEQUENCE 0
#endif
#include "kg_side.cuh"
#include "pipeline.cuh"

namespace dfgpu {

// SmartEncoding inputs of a record: [KG_WORDS][rec_cap] u32
enum : int {
  KG_VTAP_SIDE_PROTO = 0,  // vtap_id | tap_side << 16 | protocol << 24
  KG_PORT_FLAGS,           // port_dst | v6 << 16 | vip_src << 17 | vip_dst << 18
  KG_IP_SRC, KG_IP_DST, KG_EPC_SRC, KG_EPC_DST, KG_MAC_SRC_LO, KG_MAC_SRC_HI, KG_MAC_DST_LO, KG_MAC_DST_HI, KG_GPID_0,
  KG_GPID_1, KG_POD_0, KG_POD_1, KG_PID_0, KG_PID_1,
  KG_WORDS
};

struct KgSrc {
  const uint32_t* kgin;     // [KG_WORDS][rec_cap]
  const uint64_t* ip6_src;  // [rec_cap] offset | length << 32 of the v6 address bytes inside the input
  const uint64_t* ip6_dst;
  const uint32_t* rec_row;  // [rec_cap] row of the record (0xffffffff: none)
  uint32_t rec_cap;
};

__device__ __forceinline__ void kg_write_inputs(uint32_t* kgin, size_t st, uint32_t rec, uint32_t vtap, uint32_t tap_side,
                                                uint32_t proto, uint32_t port_dst, bool v6, bool vip0, bool vip1, uint32_t ip0,
                                                uint32_t ip1, uint32_t epc0, uint32_t epc1, uint64_t mac0, uint64_t mac1,
                                                uint32_t gpid0, uint32_t gpid1, uint32_t pod0, uint32_t pod1, uint32_t pid0,
                                                uint32_t pid1) {
  uint32_t* k = kgin + rec;
  k[KG_VTAP_SIDE_PROTO * st] = (vtap & 0xffffu) | (tap_side & 0xffu) << 16 | (proto & 0xffu) << 24;
  k[KG_PORT_FLAGS * st] = (port_dst & 0xffffu) | (v6 ? 1u << 16 : 0u) | (vip0 ? 1u << 17 : 0u) | (vip1 ? 1u << 18 : 0u);
  k[KG_IP_SRC * st] = ip0;
  k[KG_IP_DST * st] = ip1;
  k[KG_EPC_SRC * st] = epc0;
  k[KG_EPC_DST * st] = epc1;
  k[KG_MAC_SRC_LO * st] = (uint32_t)mac0;
  k[KG_MAC_SRC_HI * st] = (uint32_t)(mac0 >> 32);
  k[KG_MAC_DST_LO * st] = (uint32_t)mac1;
  k[KG_MAC_DST_HI * st] = (uint32_t)(mac1 >> 32);
  k[KG_GPID_0 * st] = gpid0;
  k[KG_GPID_1 * st] = gpid1;
  k[KG_POD_0 * st] = pod0;
  k[KG_POD_1 * st] = pod1;
  k[KG_PID_0 * st] = pid0;
  k[KG_PID_1 * st] = pid1;
}

// COL0: column index of region_id_0 (the 39 KnowledgeGraph columns follow in KG_COLUMNS order in both tables);
// GPID0: column of gprocess_id_0 when the kernel also resolves it from the pid (L7), -1 otherwise
template <int COL0, int GPID0>
__device__ __forceinline__ void kg_rows(const BatchParams& p, const KgSrc& q, const uint64_t* s_fix) {
  const DynState* dyn = p.dyn;
  if ((dyn->stats.need_repair && !p.repair) || dyn->overflow) return;
  const uint32_t nrec = batch_nrec(p);
  const size_t cap = q.rec_cap;
#if KG_SIDES_IN_SEQUENCE
  // one lane per RECORD, its two sides one after the other through the same code: all 32 lanes of a warp are on
  // the same side at the same time (a lane per (record, side) puts the client and the server path — different
  // branches of fill — into alternating lanes), and every column store is 32 consecutive rows
  const uint64_t total = (uint64_t)nrec;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < total; base += (uint64_t)gridDim.x * blockDim.x)
#pragma unroll 1
  for (int s = 0; s < 2; s++) {
  const uint32_t rec = (uint32_t)(base + threadIdx.x);
  const bool in_range = rec < nrec;
#else
  const uint64_t total = (uint64_t)nrec * 2;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < total; base += (uint64_t)gridDim.x * blockDim.x) {
  const uint64_t t = base + threadIdx.x;  // the loop bound is uniform across the CTA: warps stay whole
  const uint32_t rec = (uint32_t)(t >> 1);
  const int s = (int)(t & 1);
  const bool in_range = rec < nrec;
#endif
  const uint32_t row = in_range ? q.rec_row[rec] : 0xffffffffu;
  const bool ok = row != 0xffffffffu;  // both lanes of a pair agree
  const uint32_t* K = q.kgin + rec;
  const uint32_t w0 = ok ? K[KG_VTAP_SIDE_PROTO * cap] : 0u, w1 = ok ? K[KG_PORT_FLAGS * cap] : 0u;
  const uint32_t vtap_id = w0 & 0xffffu;

  kgs::SideIn in;
  in.v6 = (w1 >> 16) & 1u;
  in.vip = (w1 >> (17 + s)) & 1u;
  in.l3_epc = ok ? (int32_t)K[(KG_EPC_SRC + s) * cap] : 0;
  in.ip.ip4 = ok ? K[(KG_IP_SRC + s) * cap] : 0u;
  {
    const unsigned long long r6 = ok ? (s ? q.ip6_dst : q.ip6_src)[rec] : 0ull;
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
  in.mac = ok ? (uint64_t)K[(KG_MAC_SRC_HI + 2 * s) * cap] << 32 | K[(KG_MAC_SRC_LO + 2 * s) * cap] : 0ull;
  in.gpid = ok ? K[(KG_GPID_0 + s) * cap] : 0u;
  in.vtap_id = (uint16_t)vtap_id;
  in.port = s ? (uint16_t)(w1 & 0xffffu) : (uint16_t)0;
  in.tap_side = (w0 >> 16) & 0xffu;
  in.protocol = (uint8_t)(w0 >> 24);
  in.server_side = s == 1;
  {  // :556-567 with the partner lane (the pair sits in adjacent lanes of one warp)
    const uint32_t my_pod = ok ? K[(KG_POD_0 + s) * cap] : 0u;
    const bool my_lb = ok && kgs::loopback(in.v6, in.ip);
#if KG_SIDES_IN_SEQUENCE
    const uint32_t other_pod = ok ? K[(KG_POD_0 + (1 - s)) * cap] : 0u;
    bool other_lb = false;
    if (ok) {
      kgs::IpV oip;
      oip.ip4 = K[(KG_IP_SRC + (1 - s)) * cap];
      const unsigned long long o6 = (s ? q.ip6_src : q.ip6_dst)[rec];
      const uint32_t on6 = (uint32_t)(o6 >> 32);
      oip.len6 = on6;
      uint64_t lo = 0, hi = 0;
      const uint8_t* ob = p.in + (uint32_t)o6;
      for (uint32_t i = 0; i < (on6 < 16 ? on6 : 16u); i++) {
        const uint64_t v = (uint64_t)__ldg(ob + i) << ((i & 7) * 8);
        if (i < 8) lo |= v; else hi |= v;
      }
      oip.lo = lo;
      oip.hi = hi;
      other_lb = kgs::loopback(in.v6, oip);
    }
#else
    const uint32_t other_pod = __shfl_xor_sync(0xffffffffu, my_pod, 1);
    const bool other_lb = __shfl_xor_sync(0xffffffffu, (int)my_lb, 1) != 0;
#endif
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
  if (GPID0 >= 0) {
    const uint32_t pid = K[(KG_PID_0 + s) * cap];
    if (pid != 0 && gpid == 0) {  // l7_flow_log.go:589-596
      gpid = kgs::t64_z(p.dict, TID_VTAP_PID, (uint64_t)(uint16_t)vtap_id << 32 | (uint64_t)pid);
      k.tag_source |= TS_PROCESSID;
    }
  }
  // the _0 / _1 columns of a pair are adjacent in the schema (auto_* come in blocks of four); offsets relative
  // to region_id_0 are those of KG_COLUMNS (schema_l4.h), shared by both tables
#define KGST(T, REL, v) reinterpret_cast<T*>(p.arena + s_fix[COL0 + (REL)])[row] = (v)
  KGST(uint16_t, 0 + s, k.region_id);
  KGST(uint16_t, 2 + s, k.az_id);
  KGST(uint16_t, 4 + s, k.host_id);
  KGST(uint8_t, 6 + s, k.l3_device_type);
  KGST(uint32_t, 8 + s, k.l3_device_id);
  KGST(uint32_t, 10 + s, k.pod_node_id);
  KGST(uint16_t, 12 + s, k.pod_ns_id);
  KGST(uint32_t, 14 + s, k.pod_group_id);
  KGST(uint32_t, 16 + s, k.pod_id);
  KGST(uint16_t, 18 + s, k.pod_cluster_id);
  KGST(int32_t, 20 + s, k.l3_epc_id);
  KGST(int32_t, 22 + s, k.epc_id);
  KGST(uint16_t, 24 + s, k.subnet_id);
  KGST(uint32_t, 26 + s, k.service_id);
  KGST(uint32_t, 28 + 4 * s, k.auto_instance_id);
  KGST(uint8_t, 29 + 4 * s, k.auto_instance_type);
  KGST(uint32_t, 30 + 4 * s, k.auto_service_id);
  KGST(uint8_t, 31 + 4 * s, k.auto_service_type);
  KGST(uint8_t, 36 + s, k.tag_source);
  if (GPID0 >= 0) reinterpret_cast<uint32_t*>(p.arena + s_fix[GPID0 + s])[row] = gpid;
#undef KGST
  }
  }
}

}  // namespace dfgpu
