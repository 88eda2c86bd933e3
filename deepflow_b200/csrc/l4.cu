/*
 * l4.cu — the L4 flow-log kernel: TaggedFlow record -> one l4_flow_log row.
 * One thread per record (a warp = 32 consecutive records, so every column store
 * of the warp is one contiguous, coalesced segment).  Fuses what the reference
 * does in three goroutine hops:
 *   decode            gogo Unmarshal via SimpleDecoder.ReadPB (simple_codec.go:229-248)
 *   row build         TaggedFlowToL4FlowLog  flow_log/log_data/l4_flow_log.go:955-971
 *                     (*.Fill :467-534, FlowInfo.Fill :779-815, Metrics.Fill :817-884,
 *                      genID :949-953, geo.QueryProvince flow_log/geo/geo.go:29)
 *   SmartEncoding     KnowledgeGraph.FillL4 :730-743 -> kg.cuh
 *   column pack       L4FlowLog.AppendToColumnBlock l4_column_block.go:848-859
 */
#include "l4_decode.cuh"
#include "pipeline.cuh"
#include "schema_l4.h"

namespace dfgpu {

namespace {

template <class T>
__device__ __forceinline__ void st(const BatchParams& p, int col, uint32_t row, T v) {
  reinterpret_cast<T*>(p.arena + p.dyn->fix_off[col])[row] = v;
}
__device__ __forceinline__ void st_ip6(const BatchParams& p, int col, uint32_t row, const uint8_t* src,
                                       uint32_t len) {
  // ckdb.AppendIPv6 (ckdb.go:46-56): nil -> zeros, short -> zero padded
  uint32_t w[4] = {0, 0, 0, 0};
  uint8_t* b = reinterpret_cast<uint8_t*>(w);
  uint32_t n = len > 16 ? 16 : len;
  for (uint32_t i = 0; i < n; i++) b[i] = __ldg(src + i);
  reinterpret_cast<uint4*>(p.arena + p.dyn->fix_off[col])[row] = make_uint4(w[0], w[1], w[2], w[3]);
}

struct EmitSink {
  uint16_t* acl_out;  // next acl_gids slot
  uint8_t* agg_out;   // next aggregated_flow_ids character
  uint32_t n_agg;
  __device__ __forceinline__ void acl(uint32_t v) { *acl_out++ = (uint16_t)v; }
  __device__ void agg(uint64_t v) {  // strconv.FormatUint(n, 10) joined by ','
    if (n_agg++) *agg_out++ = ',';
    char tmp[20];
    int n = 0;
    do {
      tmp[n++] = (char)('0' + v % 10);
      v /= 10;
    } while (v);
    while (n) *agg_out++ = (uint8_t)tmp[--n];
  }
};

}  // namespace

template <int EMIT>
__global__ void __launch_bounds__(256) k_l4_pass(const __grid_constant__ BatchParams p) {
  const uint32_t rec = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = rec >> 5;
  const bool in_range = rec < batch_nrec(p);
  DynState* dyn = p.dyn;
  if (EMIT && (dyn->stats.need_repair && !p.repair)) return;  // sizes are stale: host re-runs
  if (EMIT && dyn->overflow) return;

  L4Rec r;
  memset(&r, 0, sizeof r);
  bool ok = false;
  uint32_t frame = 0;
  const uint8_t* base = nullptr;
  if (in_range) {
    frame = p.rec_frame[rec];
    base = p.in + p.rec_off[rec];
    uint32_t len = p.rec_len[rec];
    if (p.repair && rec >= p.frame_bad[frame]) {
      ok = false;  // a previous record of this frame failed: the frame was abandoned there
    } else {
      NoSink ns;
      ok = dec_tagged_flow(base, len, r, ns);
      if (!ok && !p.repair) {
        atomicMin(&p.frame_bad[frame], rec);
        dyn->stats.need_repair = 1;
      }
    }
  }
  // TaggedFlow.IsValid() (datatype/pb/flow_log.go:91-101) cannot fail: ResetAll keeps
  // every sub-message allocated.

  // ---- sizes of the variable-length outputs ------------------------------------------
  uint32_t sz_rows = ok ? 1u : 0u;
  uint32_t sz_agg = ok ? r.agg_ids_chars : 0u;
  uint32_t sz_acl = ok ? r.n_acl_gids : 0u;
  uint32_t sz_dom = ok ? r.domain_len : 0u;

  // canonical LowCardinality codes
  uint32_t prov0 = 0, prov1 = 0, tside = 0;
  if (ok) {
    prov0 = geo_code(p.geo, r.ip_src);  // Internet.Fill :531-534: always the v4 fields
    prov1 = geo_code(p.geo, r.ip_dst);
    tside = tapside_code(r.tap_side);
  }

  if (!EMIT) {
    uint32_t t0 = warp_sum(sz_rows), t1 = warp_sum(sz_agg), t2 = warp_sum(sz_acl), t3 = warp_sum(sz_dom);
    if ((threadIdx.x & 31) == 0 && warp < p.dyn->nwarps) {
      p.warp_tot[(size_t)0 * p.warp_cap + warp] = t0;
      p.warp_tot[(size_t)1 * p.warp_cap + warp] = t1;
      p.warp_tot[(size_t)2 * p.warp_cap + warp] = t2;
      p.warp_tot[(size_t)3 * p.warp_cap + warp] = t3;
    }
    lc_note(dyn->stats.lc_first[LC_PROVINCE_0], prov0, rec, ok);
    lc_note(dyn->stats.lc_first[LC_PROVINCE_1], prov1, rec, ok);
    lc_note(dyn->stats.lc_first[LC_TAPSIDE], tside, rec, ok);
    return;
  }

  // ---- EMIT ---------------------------------------------------------------------------
  uint32_t e_rows = warp_excl_scan(sz_rows), e_agg = warp_excl_scan(sz_agg),
           e_acl = warp_excl_scan(sz_acl), e_dom = warp_excl_scan(sz_dom);
  if (!ok) return;
  const uint32_t wb = warp < p.dyn->nwarps ? warp : 0;
  const uint32_t row = p.warp_base[(size_t)0 * p.warp_cap + wb] + e_rows;
  const uint64_t agg_off = (uint64_t)p.warp_base[(size_t)1 * p.warp_cap + wb] + e_agg;
  const uint64_t acl_off = (uint64_t)p.warp_base[(size_t)2 * p.warp_cap + wb] + e_acl;
  const uint64_t dom_off = (uint64_t)p.warp_base[(size_t)3 * p.warp_cap + wb] + e_dom;
  const dfgpu_frame fr = p.frames[frame];

  const bool v6 = r.eth_type == 0x86DD;  // :956
  const uint32_t time_s = (uint32_t)(r.end_time / 1000000000ull);
  {  // genID :949-953, counter pre-incremented per valid record in arrival order
    uint32_t count = p.persist->id_counter + row + 1;
    uint64_t id = (uint64_t)time_s << 32 | (uint64_t)(p.dict.analyzer_id & 0x3ff) << 22 | ((uint64_t)count & 0x3fffff);
    st<uint64_t>(p, L4C__ID, row, id);
  }
  st<uint64_t>(p, L4C_MAC_0, row, r.mac_src);
  st<uint64_t>(p, L4C_MAC_1, row, r.mac_dst);
  st<uint16_t>(p, L4C_ETH_TYPE, row, (uint16_t)r.eth_type);
  st<uint16_t>(p, L4C_VLAN, row, (uint16_t)r.vlan);
  st<uint32_t>(p, L4C_IP4_0, row, v6 ? 0u : r.ip_src);
  st<uint32_t>(p, L4C_IP4_1, row, v6 ? 0u : r.ip_dst);
  st_ip6(p, L4C_IP6_0, row, base + r.ip6_src_off, v6 ? r.ip6_src_len : 0);
  st_ip6(p, L4C_IP6_1, row, base + r.ip6_dst_off, v6 ? r.ip6_dst_len : 0);
  st<uint8_t>(p, L4C_IS_IPV4, row, v6 ? 0 : 1);
  st<uint8_t>(p, L4C_PROTOCOL, row, (uint8_t)r.proto);
  const bool tun = r.tunnel_type != 0;  // :497-511
  st<uint8_t>(p, L4C_TUNNEL_TIER, row, tun ? (uint8_t)r.tier : 0);
  st<uint16_t>(p, L4C_TUNNEL_TYPE, row, tun ? (uint16_t)r.tunnel_type : 0);
  st<uint32_t>(p, L4C_TUNNEL_TX_ID, row, tun ? r.tx_id : 0);
  st<uint32_t>(p, L4C_TUNNEL_RX_ID, row, tun ? r.rx_id : 0);
  st<uint32_t>(p, L4C_TUNNEL_TX_IP4_0, row, tun ? r.tx_ip0 : 0);
  st<uint32_t>(p, L4C_TUNNEL_TX_IP4_1, row, tun ? r.tx_ip1 : 0);
  st<uint32_t>(p, L4C_TUNNEL_RX_IP4_0, row, tun ? r.rx_ip0 : 0);
  st<uint32_t>(p, L4C_TUNNEL_RX_IP4_1, row, tun ? r.rx_ip1 : 0);
  st_ip6(p, L4C_TUNNEL_TX_IP6_0, row, nullptr, 0);
  st_ip6(p, L4C_TUNNEL_TX_IP6_1, row, nullptr, 0);
  st_ip6(p, L4C_TUNNEL_RX_IP6_0, row, nullptr, 0);
  st_ip6(p, L4C_TUNNEL_RX_IP6_1, row, nullptr, 0);
  st<uint8_t>(p, L4C_TUNNEL_IS_IPV4, row, tun ? 1 : 0);
  st<uint32_t>(p, L4C_TUNNEL_TX_MAC_0, row, tun ? r.tx_mac0 : 0);
  st<uint32_t>(p, L4C_TUNNEL_TX_MAC_1, row, tun ? r.tx_mac1 : 0);
  st<uint32_t>(p, L4C_TUNNEL_RX_MAC_0, row, tun ? r.rx_mac0 : 0);
  st<uint32_t>(p, L4C_TUNNEL_RX_MAC_1, row, tun ? r.rx_mac1 : 0);
  st<uint16_t>(p, L4C_CLIENT_PORT, row, (uint16_t)r.port_src);
  st<uint16_t>(p, L4C_SERVER_PORT, row, (uint16_t)r.port_dst);
  st<uint16_t>(p, L4C_TCP_FLAGS_BIT_0, row, (uint16_t)r.peer[0].tcp_flags);
  st<uint16_t>(p, L4C_TCP_FLAGS_BIT_1, row, (uint16_t)r.peer[1].tcp_flags);
  st<uint32_t>(p, L4C_SYN_SEQ, row, r.syn_seq);
  st<uint32_t>(p, L4C_SYN_ACK_SEQ, row, r.synack_seq);
  st<uint32_t>(p, L4C_LAST_KEEPALIVE_SEQ, row, r.last_keepalive_seq);
  st<uint32_t>(p, L4C_LAST_KEEPALIVE_ACK, row, r.last_keepalive_ack);
  const bool perf = r.has_perf_stats == 1;
  st<uint8_t>(p, L4C_L7_PROTOCOL, row, perf ? (uint8_t)r.l7_protocol : 0);
  st<uint8_t>(p, L4C_PROVINCE_0, row, dyn->lc_rank[LC_PROVINCE_0][prov0]);
  st<uint8_t>(p, L4C_PROVINCE_1, row, dyn->lc_rank[LC_PROVINCE_1][prov1]);

  {  // KnowledgeGraph.FillL4 :730-743
    KGInput in;
    in.v6 = v6;
    in.vip[0] = r.peer[0].is_vip_interface == 1;
    in.vip[1] = r.peer[1].is_vip_interface == 1;
    in.l3_epc[0] = marshal_epc(r.peer[0].l3_epc_id);
    in.l3_epc[1] = marshal_epc(r.peer[1].l3_epc_id);
    in.ip[0].ip4 = r.ip_src;
    in.ip[1].ip4 = r.ip_dst;
    in.ip[0].ip6len = r.ip6_src_len;
    in.ip[1].ip6len = r.ip6_dst_len;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      in.ip[0].ip6[i] = (uint32_t)i < r.ip6_src_len ? __ldg(base + r.ip6_src_off + i) : 0;
      in.ip[1].ip6[i] = (uint32_t)i < r.ip6_dst_len ? __ldg(base + r.ip6_dst_off + i) : 0;
    }
    in.mac[0] = r.mac_src;
    in.mac[1] = r.mac_dst;
    in.gpid[0] = r.peer[0].gpid;
    in.gpid[1] = r.peer[1].gpid;
    in.pod_id[0] = in.pod_id[1] = 0;
    in.vtap_id = (uint16_t)r.vtap_id;
    in.port = (uint16_t)r.port_dst;
    in.tap_side = r.tap_side;
    in.protocol = (uint8_t)r.proto;
    KG k;
    kg_fill(p.dict, in, k);
    st<uint16_t>(p, L4C_REGION_ID_0, row, k.region_id[0]);
    st<uint16_t>(p, L4C_REGION_ID_1, row, k.region_id[1]);
    st<uint16_t>(p, L4C_AZ_ID_0, row, k.az_id[0]);
    st<uint16_t>(p, L4C_AZ_ID_1, row, k.az_id[1]);
    st<uint16_t>(p, L4C_HOST_ID_0, row, k.host_id[0]);
    st<uint16_t>(p, L4C_HOST_ID_1, row, k.host_id[1]);
    st<uint8_t>(p, L4C_L3_DEVICE_TYPE_0, row, k.l3_device_type[0]);
    st<uint8_t>(p, L4C_L3_DEVICE_TYPE_1, row, k.l3_device_type[1]);
    st<uint32_t>(p, L4C_L3_DEVICE_ID_0, row, k.l3_device_id[0]);
    st<uint32_t>(p, L4C_L3_DEVICE_ID_1, row, k.l3_device_id[1]);
    st<uint32_t>(p, L4C_POD_NODE_ID_0, row, k.pod_node_id[0]);
    st<uint32_t>(p, L4C_POD_NODE_ID_1, row, k.pod_node_id[1]);
    st<uint16_t>(p, L4C_POD_NS_ID_0, row, k.pod_ns_id[0]);
    st<uint16_t>(p, L4C_POD_NS_ID_1, row, k.pod_ns_id[1]);
    st<uint32_t>(p, L4C_POD_GROUP_ID_0, row, k.pod_group_id[0]);
    st<uint32_t>(p, L4C_POD_GROUP_ID_1, row, k.pod_group_id[1]);
    st<uint32_t>(p, L4C_POD_ID_0, row, k.pod_id[0]);
    st<uint32_t>(p, L4C_POD_ID_1, row, k.pod_id[1]);
    st<uint16_t>(p, L4C_POD_CLUSTER_ID_0, row, k.pod_cluster_id[0]);
    st<uint16_t>(p, L4C_POD_CLUSTER_ID_1, row, k.pod_cluster_id[1]);
    st<int32_t>(p, L4C_L3_EPC_ID_0, row, k.l3_epc_id[0]);
    st<int32_t>(p, L4C_L3_EPC_ID_1, row, k.l3_epc_id[1]);
    st<int32_t>(p, L4C_EPC_ID_0, row, k.epc_id[0]);
    st<int32_t>(p, L4C_EPC_ID_1, row, k.epc_id[1]);
    st<uint16_t>(p, L4C_SUBNET_ID_0, row, k.subnet_id[0]);
    st<uint16_t>(p, L4C_SUBNET_ID_1, row, k.subnet_id[1]);
    st<uint32_t>(p, L4C_SERVICE_ID_0, row, k.service_id[0]);
    st<uint32_t>(p, L4C_SERVICE_ID_1, row, k.service_id[1]);
    st<uint32_t>(p, L4C_AUTO_INSTANCE_ID_0, row, k.auto_instance_id[0]);
    st<uint8_t>(p, L4C_AUTO_INSTANCE_TYPE_0, row, k.auto_instance_type[0]);
    st<uint32_t>(p, L4C_AUTO_SERVICE_ID_0, row, k.auto_service_id[0]);
    st<uint8_t>(p, L4C_AUTO_SERVICE_TYPE_0, row, k.auto_service_type[0]);
    st<uint32_t>(p, L4C_AUTO_INSTANCE_ID_1, row, k.auto_instance_id[1]);
    st<uint8_t>(p, L4C_AUTO_INSTANCE_TYPE_1, row, k.auto_instance_type[1]);
    st<uint32_t>(p, L4C_AUTO_SERVICE_ID_1, row, k.auto_service_id[1]);
    st<uint8_t>(p, L4C_AUTO_SERVICE_TYPE_1, row, k.auto_service_type[1]);
    st<uint8_t>(p, L4C_TAG_SOURCE_0, row, k.tag_source[0]);
    st<uint8_t>(p, L4C_TAG_SOURCE_1, row, k.tag_source[1]);
    st<uint16_t>(p, L4C_TEAM_ID, row, (uint16_t)fr.team_id);
  }

  // FlowInfo.Fill :779-815
  st<uint32_t>(p, L4C_TIME, row, time_s);
  st<uint16_t>(p, L4C_CLOSE_TYPE, row, (uint16_t)r.close_type);
  st<uint16_t>(p, L4C_SIGNAL_SOURCE, row, (uint16_t)r.signal_source);
  st<uint64_t>(p, L4C_FLOW_ID, row, r.flow_id);
  st<uint32_t>(p, L4C_INIT_IPID, row, r.init_ipid);
  st<uint8_t>(p, L4C_CAPTURE_NETWORK_TYPE_ID, row, (uint8_t)r.tap_type);
  st<uint8_t>(p, L4C_NAT_SOURCE, row, (uint8_t)((r.tap_port >> 36) & 0xf));
  st<uint8_t>(p, L4C_CAPTURE_NIC_TYPE, row, (uint8_t)(r.tap_port >> 60));
  st<uint32_t>(p, L4C_CAPTURE_NIC, row, (uint32_t)r.tap_port);
  st<uint8_t>(p, L4C_OBSERVATION_POINT, row, dyn->lc_rank[LC_TAPSIDE][tside]);
  st<uint16_t>(p, L4C_AGENT_ID, row, (uint16_t)r.vtap_id);
  st<uint8_t>(p, L4C_L2_END_0, row, r.peer[0].is_l2_end == 1);
  st<uint8_t>(p, L4C_L2_END_1, row, r.peer[1].is_l2_end == 1);
  st<uint8_t>(p, L4C_L3_END_0, row, r.peer[0].is_l3_end == 1);
  st<uint8_t>(p, L4C_L3_END_1, row, r.peer[1].is_l3_end == 1);
  st<int64_t>(p, L4C_START_TIME, row, (int64_t)r.start_time / 1000);
  st<int64_t>(p, L4C_END_TIME, row, (int64_t)r.end_time / 1000);
  st<uint64_t>(p, L4C_DURATION, row, r.duration / 1000);
  st<uint8_t>(p, L4C_IS_NEW_FLOW, row, (uint8_t)r.is_new_flow);
  st<uint8_t>(p, L4C_STATUS, row, l4_status((uint8_t)(uint16_t)r.close_type, (uint8_t)r.proto));
  st<uint32_t>(p, L4C_GPROCESS_ID_0, row, r.peer[0].gpid);
  st<uint32_t>(p, L4C_GPROCESS_ID_1, row, r.peer[1].gpid);
  st<uint32_t>(p, L4C_NAT_REAL_IP4_0, row, r.peer[0].real_ip);
  st<uint32_t>(p, L4C_NAT_REAL_IP4_1, row, r.peer[1].real_ip);
  st<uint16_t>(p, L4C_NAT_REAL_PORT_0, row, (uint16_t)r.peer[0].real_port);
  st<uint16_t>(p, L4C_NAT_REAL_PORT_1, row, (uint16_t)r.peer[1].real_port);
  st<uint8_t>(p, L4C_DIRECTION_SCORE, row, (uint8_t)r.direction_score);

  // Metrics.Fill :817-884 + the float64 casts of l4_column_block.go:765-772
  st<uint64_t>(p, L4C_PACKET_TX, row, r.peer[0].packet_count);
  st<uint64_t>(p, L4C_PACKET_RX, row, r.peer[1].packet_count);
  st<uint64_t>(p, L4C_BYTE_TX, row, r.peer[0].byte_count);
  st<uint64_t>(p, L4C_BYTE_RX, row, r.peer[1].byte_count);
  st<uint64_t>(p, L4C_L3_BYTE_TX, row, r.peer[0].l3_byte_count);
  st<uint64_t>(p, L4C_L3_BYTE_RX, row, r.peer[1].l3_byte_count);
  st<uint64_t>(p, L4C_L4_BYTE_TX, row, r.peer[0].l4_byte_count);
  st<uint64_t>(p, L4C_L4_BYTE_RX, row, r.peer[1].l4_byte_count);
  st<uint64_t>(p, L4C_TOTAL_PACKET_TX, row, r.peer[0].total_packet_count);
  st<uint64_t>(p, L4C_TOTAL_PACKET_RX, row, r.peer[1].total_packet_count);
  st<uint64_t>(p, L4C_TOTAL_BYTE_TX, row, r.peer[0].total_byte_count);
  st<uint64_t>(p, L4C_TOTAL_BYTE_RX, row, r.peer[1].total_byte_count);
#define PF(x) (perf ? (x) : 0)
  st<uint32_t>(p, L4C_L7_REQUEST, row, PF(r.request_count));
  st<uint32_t>(p, L4C_L7_RESPONSE, row, PF(r.response_count));
  st<uint32_t>(p, L4C_L7_PARSE_FAILED, row, PF(r.l7_failed_count));
  st<double>(p, L4C_RTT, row, (double)PF(r.rtt));
  st<double>(p, L4C_RTT_CLIENT, row, (double)PF(r.rtt_client_max));
  st<double>(p, L4C_RTT_SERVER, row, (double)PF(r.rtt_server_max));
  st<double>(p, L4C_TLS_RTT, row, (double)PF(r.tls_rtt));
  st<double>(p, L4C_SRT_SUM, row, (double)PF(r.srt_sum));
  st<double>(p, L4C_ART_SUM, row, (double)PF(r.art_sum));
  st<double>(p, L4C_RRT_SUM, row, __ull2double_rn(perf ? r.rrt_sum : 0ull));
  st<double>(p, L4C_CIT_SUM, row, (double)PF(r.cit_sum));
  st<uint64_t>(p, L4C_SRT_COUNT, row, (uint64_t)PF(r.srt_count));
  st<uint64_t>(p, L4C_ART_COUNT, row, (uint64_t)PF(r.art_count));
  st<uint64_t>(p, L4C_RRT_COUNT, row, (uint64_t)PF(r.rrt_count));
  st<uint64_t>(p, L4C_CIT_COUNT, row, (uint64_t)PF(r.cit_count));
  st<uint32_t>(p, L4C_SRT_MAX, row, PF(r.srt_max));
  st<uint32_t>(p, L4C_ART_MAX, row, PF(r.art_max));
  st<uint32_t>(p, L4C_RRT_MAX, row, PF(r.rrt_max));
  st<uint32_t>(p, L4C_CIT_MAX, row, PF(r.cit_max));
  st<uint32_t>(p, L4C_RETRANS_TX, row, PF(r.tx_retrans));
  st<uint32_t>(p, L4C_RETRANS_RX, row, PF(r.rx_retrans));
  st<uint32_t>(p, L4C_ZERO_WIN_TX, row, PF(r.tx_zero_win));
  st<uint32_t>(p, L4C_ZERO_WIN_RX, row, PF(r.rx_zero_win));
  st<uint32_t>(p, L4C_SYN_COUNT, row, PF(r.syn_count));
  st<uint32_t>(p, L4C_SYNACK_COUNT, row, PF(r.synack_count));
  st<uint32_t>(p, L4C_RETRANS_SYN, row, (perf && r.syn_count > 0) ? r.syn_count - 1 : 0);
  st<uint32_t>(p, L4C_RETRANS_SYNACK, row, (perf && r.synack_count > 0) ? r.synack_count - 1 : 0);
  st<uint32_t>(p, L4C_L7_CLIENT_ERROR, row, PF(r.err_client_count));
  st<uint32_t>(p, L4C_L7_SERVER_ERROR, row, PF(r.err_server_count));
  st<uint32_t>(p, L4C_L7_SERVER_TIMEOUT, row, PF(r.err_timeout));
  st<uint32_t>(p, L4C_L7_ERROR, row, PF(r.err_client_count + r.err_server_count));
  st<uint32_t>(p, L4C_OOO_RX, row, PF(r.rx_ooo));
  st<uint32_t>(p, L4C_OOO_TX, row, PF(r.tx_ooo));
  st<uint32_t>(p, L4C_FIN_COUNT, row, PF(r.fin_count));
#undef PF

  // ---- variable-length columns ---------------------------------------------------------
  // aggregated_flow_ids (ColStr): Pos per row + characters in stream 0
  dfgpu_pos* agg_pos = reinterpret_cast<dfgpu_pos*>(p.arena + p.dyn->aux_off[L4C_AGGREGATED_FLOW_IDS]);
  agg_pos[row] = dfgpu_pos{(int64_t)agg_off, (int64_t)(agg_off + sz_agg)};
  // acl_gids (ColArr[uint16]): cumulative END offset per row + values in stream 1
  reinterpret_cast<uint64_t*>(p.arena + p.dyn->aux_off[L4C_ACL_GIDS])[row] = acl_off + sz_acl;
  // request_domain (ColStr)
  dfgpu_pos* dom_pos = reinterpret_cast<dfgpu_pos*>(p.arena + p.dyn->aux_off[L4C_REQUEST_DOMAIN]);
  dom_pos[row] = dfgpu_pos{(int64_t)dom_off, (int64_t)(dom_off + sz_dom)};
  if (sz_dom) {
    uint8_t* dst = p.arena + dyn->var_buf_off[2] + dom_off;
    const uint8_t* src = base + r.domain_off;
    for (uint32_t i = 0; i < sz_dom; i++) dst[i] = __ldg(src + i);
  }
  if (sz_agg | sz_acl) {  // rare: re-walk the record streaming the repeated scalars out
    EmitSink sink;
    sink.acl_out = reinterpret_cast<uint16_t*>(p.arena + dyn->var_buf_off[1]) + acl_off;
    sink.agg_out = p.arena + dyn->var_buf_off[0] + agg_off;
    sink.n_agg = 0;
    L4Rec scratch;
    memset(&scratch, 0, sizeof scratch);
    dec_tagged_flow(base, p.rec_len[rec], scratch, sink);
  }
}

void launch_l4_pass(const BatchParams& p, int emit, cudaStream_t s) {
  dim3 block(256), grid(rec_grid(p));
  if (emit) k_l4_pass<1><<<grid, block, 0, s>>>(p);
  else k_l4_pass<0><<<grid, block, 0, s>>>(p);
}

}  // namespace dfgpu
