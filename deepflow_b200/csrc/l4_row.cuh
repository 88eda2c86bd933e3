/*
 * l4_row.cuh — one decoded TaggedFlow record -> the values of its l4_flow_log row.  Pure per-record logic
 * (no warp collectives, no CUDA types): the decode kernel (l4.cu) runs it on the device with an emitter that
 * stores coalesced columns; tests/harness runs the same code on the host against the oracle.  Restates
 *   TaggedFlowToL4FlowLog  flow_log/log_data/l4_flow_log.go:955-971
 *   DataLinkLayer.Fill :467, NetworkLayer.Fill :484-512 (tunnel fields only if TunnelType != 0),
 *   TransportLayer.Fill :514, ApplicationLayer.Fill :525, Internet.Fill :531-534,
 *   FlowInfo.Fill :779-815 (ns -> us / s, TapPort.SplitToPortTypeTunnel tap_port.go:113-115, getStatus :745-760),
 *   Metrics.Fill :817-884 (only if has_perf_stats == 1; RetransSyn = SynCount - 1 :877-882; L7Error = client +
 *   server :839), genID :949-953
 *   L4FlowLog.AppendToColumnBlock l4_column_block.go:848-859 (u32 / u64 -> float64 casts :765-772)
 */
#pragma once
#include "l4_tok.h"
#include "schema_l4.h"

#include "../../include/dfgpu.h"

namespace dfgpu {

// getStatus l4_flow_log.go:745-760 (CloseType is a uint8, datatype/flow.go:33-84)
DF_HD uint8_t l4_status_hd(uint8_t t, uint8_t proto) {
  const bool tcp = proto == 6;
  if (t == 1 || t == 5 || t == 20 || (!tcp && t == 3) || t == 11 || t == 8) return 0;
  if (t == 10 || t == 9 || t == 13 || t == 18) return 4;
  if (tcp && (t == 2 || t == 3 || t == 7 || t == 15 || t == 17 || t == 19)) return 3;
  return 2;
}
// MarshalInt32WithSpecialID flow-metrics/tag.go:560-565
DF_HD int32_t marshal_epc_hd(int32_t v) {
  if (v > 0 || v == -1 || v == -2) return v;
  return (int32_t)((uint32_t)v & 0xffffu);
}
DF_HD uint32_t l4_tapside_code_hd(uint32_t v) {  // canonical ids of TAPSideEnum.String() (tag.go:211-233)
  switch (v & 0xff) {
    case 0: return 1;  case 1: return 2;  case 2: return 3;  case 4: return 4;
    case 9: return 5;  case 10: return 6; case 17: return 7; case 18: return 8;
    case 25: return 9; case 26: return 10; case 33: return 11; case 34: return 12;
    case 41: return 13; case 42: return 14; case 49: return 15; case 50: return 16;
    case 48: return 17;
  }
  return 0;
}

struct L4RowInfo {
  uint32_t tside;                 // canonical tap-side code
  uint32_t ip_src, ip_dst;        // Internet.Fill: the province look-up always takes the v4 fields
  uint32_t dom_off, dom_len;      // request_domain
  uint32_t n_acl, n_agg, agg_chars;
};

// S: uint32_t w(int word); uint64_t w64(int word)
// E: template <class T> void fix(int col, T v); void ip6(int col, uint32_t off, uint32_t len);
//    void kg(...)  SmartEncoding inputs (l4.cu)
template <class S, class E>
DF_HD L4RowInfo l4_build_row(const S& sl, E& e, const uint32_t analyzer_id, const uint32_t id_count /* id_counter + row + 1 */) {
  L4RowInfo o;
  const uint32_t eth_type = sl.w(W4_ETH_TYPE);
  const bool v6 = eth_type == 0x86DD;  // :956
  const uint64_t end_time = sl.w64(W4_END_TIME);
  const uint32_t time_s = (uint32_t)(end_time / 1000000000ull);
  const int P0 = W4_PEER0, P1 = W4_PEER1;
  // genID :949-953, counter pre-incremented per row in arrival order
  e.template fix<uint64_t>(L4C__ID, (uint64_t)time_s << 32 | (uint64_t)(analyzer_id & 0x3ff) << 22 | ((uint64_t)id_count & 0x3fffff));
  e.template fix<uint64_t>(L4C_MAC_0, sl.w64(W4_MAC_SRC));
  e.template fix<uint64_t>(L4C_MAC_1, sl.w64(W4_MAC_DST));
  e.template fix<uint16_t>(L4C_ETH_TYPE, (uint16_t)eth_type);
  e.template fix<uint16_t>(L4C_VLAN, (uint16_t)sl.w(W4_VLAN));
  const uint32_t ip_src = sl.w(W4_IP_SRC), ip_dst = sl.w(W4_IP_DST);
  e.template fix<uint32_t>(L4C_IP4_0, v6 ? 0u : ip_src);
  e.template fix<uint32_t>(L4C_IP4_1, v6 ? 0u : ip_dst);
  e.ip6(L4C_IP6_0, sl.w(W4_IP6_SRC_OFF), v6 ? sl.w(W4_IP6_SRC_LEN) : 0u);
  e.ip6(L4C_IP6_1, sl.w(W4_IP6_DST_OFF), v6 ? sl.w(W4_IP6_DST_LEN) : 0u);
  e.template fix<uint8_t>(L4C_IS_IPV4, v6 ? 0 : 1);
  const uint32_t proto = sl.w(W4_PROTO);
  e.template fix<uint8_t>(L4C_PROTOCOL, (uint8_t)proto);
  const uint32_t tunnel_type = sl.w(W4_TUNNEL_TYPE);
  const bool tun = tunnel_type != 0;  // :497-511
  e.template fix<uint8_t>(L4C_TUNNEL_TIER, tun ? (uint8_t)sl.w(W4_TIER) : 0);
  e.template fix<uint16_t>(L4C_TUNNEL_TYPE, tun ? (uint16_t)tunnel_type : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_TX_ID, tun ? sl.w(W4_TX_ID) : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_RX_ID, tun ? sl.w(W4_RX_ID) : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_TX_IP4_0, tun ? sl.w(W4_TX_IP0) : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_TX_IP4_1, tun ? sl.w(W4_TX_IP1) : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_RX_IP4_0, tun ? sl.w(W4_RX_IP0) : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_RX_IP4_1, tun ? sl.w(W4_RX_IP1) : 0);
  e.ip6(L4C_TUNNEL_TX_IP6_0, 0, 0);
  e.ip6(L4C_TUNNEL_TX_IP6_1, 0, 0);
  e.ip6(L4C_TUNNEL_RX_IP6_0, 0, 0);
  e.ip6(L4C_TUNNEL_RX_IP6_1, 0, 0);
  e.template fix<uint8_t>(L4C_TUNNEL_IS_IPV4, tun ? 1 : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_TX_MAC_0, tun ? sl.w(W4_TX_MAC0) : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_TX_MAC_1, tun ? sl.w(W4_TX_MAC1) : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_RX_MAC_0, tun ? sl.w(W4_RX_MAC0) : 0);
  e.template fix<uint32_t>(L4C_TUNNEL_RX_MAC_1, tun ? sl.w(W4_RX_MAC1) : 0);
  const uint32_t port_src = sl.w(W4_PORT_SRC), port_dst = sl.w(W4_PORT_DST);
  e.template fix<uint16_t>(L4C_CLIENT_PORT, (uint16_t)port_src);
  e.template fix<uint16_t>(L4C_SERVER_PORT, (uint16_t)port_dst);
  e.template fix<uint16_t>(L4C_TCP_FLAGS_BIT_0, (uint16_t)sl.w(P0 + P4_TCP_FLAGS));
  e.template fix<uint16_t>(L4C_TCP_FLAGS_BIT_1, (uint16_t)sl.w(P1 + P4_TCP_FLAGS));
  e.template fix<uint32_t>(L4C_SYN_SEQ, sl.w(W4_SYN_SEQ));
  e.template fix<uint32_t>(L4C_SYN_ACK_SEQ, sl.w(W4_SYNACK_SEQ));
  e.template fix<uint32_t>(L4C_LAST_KEEPALIVE_SEQ, sl.w(W4_KA_SEQ));
  e.template fix<uint32_t>(L4C_LAST_KEEPALIVE_ACK, sl.w(W4_KA_ACK));
  const bool perf = sl.w(W4_HAS_PERF) == 1;
  e.template fix<uint8_t>(L4C_L7_PROTOCOL, perf ? (uint8_t)sl.w(W4_L7_PROTOCOL) : 0);

  // FlowInfo.Fill :779-815
  const uint64_t tap_port = sl.w64(W4_TAP_PORT);
  const uint32_t vtap_id = sl.w(W4_VTAP_ID), close_type = sl.w(W4_CLOSE_TYPE), tap_side = sl.w(W4_TAP_SIDE);
  e.template fix<uint32_t>(L4C_TIME, time_s);
  e.template fix<uint16_t>(L4C_CLOSE_TYPE, (uint16_t)close_type);
  e.template fix<uint16_t>(L4C_SIGNAL_SOURCE, (uint16_t)sl.w(W4_SIGNAL_SOURCE));
  e.template fix<uint64_t>(L4C_FLOW_ID, sl.w64(W4_FLOW_ID));
  e.template fix<uint32_t>(L4C_INIT_IPID, sl.w(W4_INIT_IPID));
  e.template fix<uint8_t>(L4C_CAPTURE_NETWORK_TYPE_ID, (uint8_t)sl.w(W4_TAP_TYPE));
  e.template fix<uint8_t>(L4C_NAT_SOURCE, (uint8_t)((tap_port >> 36) & 0xf));
  e.template fix<uint8_t>(L4C_CAPTURE_NIC_TYPE, (uint8_t)(tap_port >> 60));
  e.template fix<uint32_t>(L4C_CAPTURE_NIC, (uint32_t)tap_port);
  e.template fix<uint16_t>(L4C_AGENT_ID, (uint16_t)vtap_id);
  e.template fix<uint8_t>(L4C_L2_END_0, sl.w(P0 + P4_L2_END) == 1);
  e.template fix<uint8_t>(L4C_L2_END_1, sl.w(P1 + P4_L2_END) == 1);
  e.template fix<uint8_t>(L4C_L3_END_0, sl.w(P0 + P4_L3_END) == 1);
  e.template fix<uint8_t>(L4C_L3_END_1, sl.w(P1 + P4_L3_END) == 1);
  e.template fix<int64_t>(L4C_START_TIME, (int64_t)sl.w64(W4_START_TIME) / 1000);
  e.template fix<int64_t>(L4C_END_TIME, (int64_t)end_time / 1000);
  e.template fix<uint64_t>(L4C_DURATION, sl.w64(W4_DURATION) / 1000);
  e.template fix<uint8_t>(L4C_IS_NEW_FLOW, (uint8_t)sl.w(W4_IS_NEW_FLOW));
  e.template fix<uint8_t>(L4C_STATUS, l4_status_hd((uint8_t)(uint16_t)close_type, (uint8_t)proto));
  const uint32_t gpid0 = sl.w(P0 + P4_GPID), gpid1 = sl.w(P1 + P4_GPID);
  e.template fix<uint32_t>(L4C_GPROCESS_ID_0, gpid0);
  e.template fix<uint32_t>(L4C_GPROCESS_ID_1, gpid1);
  e.template fix<uint32_t>(L4C_NAT_REAL_IP4_0, sl.w(P0 + P4_REAL_IP));
  e.template fix<uint32_t>(L4C_NAT_REAL_IP4_1, sl.w(P1 + P4_REAL_IP));
  e.template fix<uint16_t>(L4C_NAT_REAL_PORT_0, (uint16_t)sl.w(P0 + P4_REAL_PORT));
  e.template fix<uint16_t>(L4C_NAT_REAL_PORT_1, (uint16_t)sl.w(P1 + P4_REAL_PORT));
  e.template fix<uint8_t>(L4C_DIRECTION_SCORE, (uint8_t)sl.w(W4_DIRECTION_SCORE));

  // Metrics.Fill :817-884 + the float64 casts of l4_column_block.go:765-772
  e.template fix<uint64_t>(L4C_PACKET_TX, sl.w64(P0 + P4_PACKET));
  e.template fix<uint64_t>(L4C_PACKET_RX, sl.w64(P1 + P4_PACKET));
  e.template fix<uint64_t>(L4C_BYTE_TX, sl.w64(P0 + P4_BYTE));
  e.template fix<uint64_t>(L4C_BYTE_RX, sl.w64(P1 + P4_BYTE));
  e.template fix<uint64_t>(L4C_L3_BYTE_TX, sl.w64(P0 + P4_L3_BYTE));
  e.template fix<uint64_t>(L4C_L3_BYTE_RX, sl.w64(P1 + P4_L3_BYTE));
  e.template fix<uint64_t>(L4C_L4_BYTE_TX, sl.w64(P0 + P4_L4_BYTE));
  e.template fix<uint64_t>(L4C_L4_BYTE_RX, sl.w64(P1 + P4_L4_BYTE));
  e.template fix<uint64_t>(L4C_TOTAL_PACKET_TX, sl.w64(P0 + P4_TOTAL_PACKET));
  e.template fix<uint64_t>(L4C_TOTAL_PACKET_RX, sl.w64(P1 + P4_TOTAL_PACKET));
  e.template fix<uint64_t>(L4C_TOTAL_BYTE_TX, sl.w64(P0 + P4_TOTAL_BYTE));
  e.template fix<uint64_t>(L4C_TOTAL_BYTE_RX, sl.w64(P1 + P4_TOTAL_BYTE));
#define PF(wd) (perf ? sl.w(wd) : 0u)
  e.template fix<uint32_t>(L4C_L7_REQUEST, PF(W4_REQUEST_COUNT));
  e.template fix<uint32_t>(L4C_L7_RESPONSE, PF(W4_RESPONSE_COUNT));
  e.template fix<uint32_t>(L4C_L7_PARSE_FAILED, PF(W4_L7_FAILED));
  e.template fix<double>(L4C_RTT, (double)PF(W4_RTT));
  e.template fix<double>(L4C_RTT_CLIENT, (double)PF(W4_RTT_CLIENT_MAX));
  e.template fix<double>(L4C_RTT_SERVER, (double)PF(W4_RTT_SERVER_MAX));
  e.template fix<double>(L4C_TLS_RTT, (double)PF(W4_TLS_RTT));
  e.template fix<double>(L4C_SRT_SUM, (double)PF(W4_SRT_SUM));
  e.template fix<double>(L4C_ART_SUM, (double)PF(W4_ART_SUM));
  e.template fix<double>(L4C_RRT_SUM, (double)(perf ? sl.w64(W4_RRT_SUM) : 0ull));  // round to nearest even, like Go's float64(uint64)
  e.template fix<double>(L4C_CIT_SUM, (double)PF(W4_CIT_SUM));
  e.template fix<uint64_t>(L4C_SRT_COUNT, (uint64_t)PF(W4_SRT_COUNT));
  e.template fix<uint64_t>(L4C_ART_COUNT, (uint64_t)PF(W4_ART_COUNT));
  e.template fix<uint64_t>(L4C_RRT_COUNT, (uint64_t)PF(W4_RRT_COUNT));
  e.template fix<uint64_t>(L4C_CIT_COUNT, (uint64_t)PF(W4_CIT_COUNT));
  e.template fix<uint32_t>(L4C_SRT_MAX, PF(W4_SRT_MAX));
  e.template fix<uint32_t>(L4C_ART_MAX, PF(W4_ART_MAX));
  e.template fix<uint32_t>(L4C_RRT_MAX, PF(W4_RRT_MAX));
  e.template fix<uint32_t>(L4C_CIT_MAX, PF(W4_CIT_MAX));
  e.template fix<uint32_t>(L4C_RETRANS_TX, PF(W4_TX_RETRANS));
  e.template fix<uint32_t>(L4C_RETRANS_RX, PF(W4_RX_RETRANS));
  e.template fix<uint32_t>(L4C_ZERO_WIN_TX, PF(W4_TX_ZERO_WIN));
  e.template fix<uint32_t>(L4C_ZERO_WIN_RX, PF(W4_RX_ZERO_WIN));
  const uint32_t syn = PF(W4_SYN_COUNT), synack = PF(W4_SYNACK_COUNT);
  e.template fix<uint32_t>(L4C_SYN_COUNT, syn);
  e.template fix<uint32_t>(L4C_SYNACK_COUNT, synack);
  e.template fix<uint32_t>(L4C_RETRANS_SYN, syn > 0 ? syn - 1 : 0);
  e.template fix<uint32_t>(L4C_RETRANS_SYNACK, synack > 0 ? synack - 1 : 0);
  const uint32_t ec = PF(W4_ERR_CLIENT), es = PF(W4_ERR_SERVER);
  e.template fix<uint32_t>(L4C_L7_CLIENT_ERROR, ec);
  e.template fix<uint32_t>(L4C_L7_SERVER_ERROR, es);
  e.template fix<uint32_t>(L4C_L7_SERVER_TIMEOUT, PF(W4_ERR_TIMEOUT));
  e.template fix<uint32_t>(L4C_L7_ERROR, ec + es);
  e.template fix<uint32_t>(L4C_OOO_RX, PF(W4_RX_OOO));
  e.template fix<uint32_t>(L4C_OOO_TX, PF(W4_TX_OOO));
  e.template fix<uint32_t>(L4C_FIN_COUNT, PF(W4_FIN_COUNT));
#undef PF

  // SmartEncoding inputs (KnowledgeGraph.FillL4 :730-743 runs in the SmartEncoding kernel)
  e.kg(vtap_id, tap_side, proto, port_dst, v6, sl.w(P0 + P4_VIP_IF) == 1, sl.w(P1 + P4_VIP_IF) == 1, ip_src, ip_dst,
       (uint32_t)marshal_epc_hd((int32_t)sl.w(P0 + P4_L3_EPC)), (uint32_t)marshal_epc_hd((int32_t)sl.w(P1 + P4_L3_EPC)),
       sl.w64(W4_MAC_SRC), sl.w64(W4_MAC_DST), gpid0, gpid1, sl.w(W4_IP6_SRC_OFF), sl.w(W4_IP6_SRC_LEN), sl.w(W4_IP6_DST_OFF),
       sl.w(W4_IP6_DST_LEN));

  o.tside = l4_tapside_code_hd(tap_side);
  o.ip_src = ip_src;
  o.ip_dst = ip_dst;
  o.dom_off = sl.w(W4_DOMAIN_OFF);
  o.dom_len = sl.w(W4_DOMAIN_LEN);
  o.n_acl = sl.w(W4_ACL_N);
  o.n_agg = sl.w(W4_AGG_N);
  o.agg_chars = sl.w(W4_AGG_CHARS);
  return o;
}

}  // namespace dfgpu
