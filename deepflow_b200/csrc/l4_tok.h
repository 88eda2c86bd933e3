/*
 * l4_tok.h — word layout + tokenizer table of flow_log.TaggedFlow (message/flow_log.proto:14-177) for
 * pb_tok.cuh.  Words are 4 bytes (shared memory, one column per thread); a uint64 field takes two consecutive
 * words (lo, hi); the three bytes / string fields keep (offset, length) in two words each.
 * Replaces the gogo-generated (*TaggedFlow).Unmarshal + pb.TaggedFlow.ResetAll
 * (server/libs/datatype/pb/flow_log.go:122-165): every sub-message is treated as pre-allocated (never nil),
 * repeated occurrences merge, the last scalar wins.
 */
#pragma once
#include <cstdint>
#include <vector>

#include "pb_tok.cuh"

namespace dfgpu {

enum L4Msg : uint32_t {
  M4_TOP = 0, M4_FLOW, M4_KEY, M4_PEER0, M4_PEER1, M4_TUNNEL, M4_PERF, M4_TCP, M4_CNT_TX, M4_CNT_RX, M4_L7PERF, M4_COUNT
};
constexpr uint32_t L4_MAXF = 32;

// FlowMetricsPeer words, relative to W4_PEER0 / W4_PEER1
enum : int {
  P4_BYTE = 0, P4_L3_BYTE = 2, P4_L4_BYTE = 4, P4_PACKET = 6, P4_TOTAL_BYTE = 8, P4_TOTAL_PACKET = 10, P4_TCP_FLAGS = 12,
  P4_L3_EPC, P4_L2_END, P4_L3_END, P4_VIP_IF, P4_REAL_IP, P4_REAL_PORT, P4_GPID, P4_WORDS
};

enum L4Word : int {
  // FlowKey
  W4_VTAP_ID, W4_TAP_TYPE, W4_TAP_PORT, W4_TAP_PORT_HI, W4_MAC_SRC, W4_MAC_SRC_HI, W4_MAC_DST, W4_MAC_DST_HI, W4_IP_SRC,
  W4_IP_DST, W4_PORT_SRC, W4_PORT_DST, W4_PROTO,
  W4_IP6_SRC_OFF, W4_IP6_SRC_LEN, W4_IP6_DST_OFF, W4_IP6_DST_LEN,
  W4_PEER0, W4_PEER1 = W4_PEER0 + P4_WORDS,
  // TunnelField
  W4_TX_IP0 = W4_PEER1 + P4_WORDS, W4_TX_IP1, W4_RX_IP0, W4_RX_IP1, W4_TX_MAC0, W4_TX_MAC1, W4_RX_MAC0, W4_RX_MAC1, W4_TX_ID,
  W4_RX_ID, W4_TUNNEL_TYPE, W4_TIER,
  // Flow
  W4_FLOW_ID, W4_FLOW_ID_HI, W4_START_TIME, W4_START_TIME_HI, W4_END_TIME, W4_END_TIME_HI, W4_DURATION, W4_DURATION_HI,
  W4_VLAN, W4_ETH_TYPE, W4_HAS_PERF, W4_CLOSE_TYPE, W4_SIGNAL_SOURCE, W4_IS_NEW_FLOW, W4_TAP_SIDE, W4_SYN_SEQ, W4_SYNACK_SEQ,
  W4_KA_SEQ, W4_KA_ACK, W4_DIRECTION_SCORE, W4_INIT_IPID, W4_DOMAIN_OFF, W4_DOMAIN_LEN,
  W4_ACL_N, W4_AGG_N, W4_AGG_CHARS,
  // FlowPerfStats
  W4_L7_PROTOCOL, W4_L7_FAILED,
  // TCPPerfStats (+ TcpPerfCountsPeer tx / rx)
  W4_RTT_CLIENT_MAX, W4_RTT_SERVER_MAX, W4_SRT_MAX, W4_ART_MAX, W4_RTT, W4_SRT_SUM, W4_ART_SUM, W4_SRT_COUNT, W4_ART_COUNT,
  W4_SYN_COUNT, W4_SYNACK_COUNT, W4_CIT_MAX, W4_CIT_SUM, W4_CIT_COUNT, W4_FIN_COUNT,
  W4_TX_RETRANS, W4_TX_ZERO_WIN, W4_TX_OOO, W4_RX_RETRANS, W4_RX_ZERO_WIN, W4_RX_OOO,
  // L7PerfStats
  W4_REQUEST_COUNT, W4_RESPONSE_COUNT, W4_ERR_CLIENT, W4_ERR_SERVER, W4_ERR_TIMEOUT, W4_RRT_COUNT, W4_RRT_SUM, W4_RRT_SUM_HI,
  W4_RRT_MAX, W4_TLS_RTT,
  // sub-message occurrence counters nobody reads, and fields that are decoded (wire type checked) but unused
  W4_CNT, W4_DUMMY, W4_DUMMY_HI,
  W4_COUNT
};

// repeated scalars of a record, captured while it is decoded: [L4E_COUNT] u64 per record
enum : uint32_t { L4E_ACL = 0, L4E_AGG = 8, L4E_COUNT = 16 };
constexpr uint32_t L4E_CAP = 8;  // more elements than this: the pack kernel walks the record again

inline std::vector<uint32_t> l4_tok_table() {
  std::vector<uint32_t> t(M4_COUNT * L4_MAXF, 0u);
  auto set = [&](uint32_t m, uint32_t f, uint32_t ent) { t[m * L4_MAXF + f] = ent; };
  auto u32 = [&](uint32_t m, uint32_t f, int w) { set(m, f, TK_ENT(TK_U32, w, 0, 0)); };
  auto u64 = [&](uint32_t m, uint32_t f, int w) { set(m, f, TK_ENT(TK_U64, w, 0, 0)); };
  auto msg = [&](uint32_t m, uint32_t f, uint32_t sub) { set(m, f, TK_ENT(TK_MSG, W4_CNT, sub, 0)); };
  msg(M4_TOP, 1, M4_FLOW);
  // Flow
  msg(M4_FLOW, 1, M4_KEY); msg(M4_FLOW, 2, M4_PEER0); msg(M4_FLOW, 3, M4_PEER1); msg(M4_FLOW, 4, M4_TUNNEL);
  u64(M4_FLOW, 5, W4_FLOW_ID); u64(M4_FLOW, 6, W4_START_TIME); u64(M4_FLOW, 7, W4_END_TIME); u64(M4_FLOW, 8, W4_DURATION);
  u32(M4_FLOW, 10, W4_VLAN); u32(M4_FLOW, 11, W4_ETH_TYPE); u32(M4_FLOW, 12, W4_HAS_PERF); msg(M4_FLOW, 13, M4_PERF);
  u32(M4_FLOW, 14, W4_CLOSE_TYPE); u32(M4_FLOW, 15, W4_SIGNAL_SOURCE); u32(M4_FLOW, 16, W4_DUMMY); u32(M4_FLOW, 17, W4_DUMMY);
  u32(M4_FLOW, 18, W4_IS_NEW_FLOW); u32(M4_FLOW, 19, W4_TAP_SIDE); u32(M4_FLOW, 20, W4_SYN_SEQ); u32(M4_FLOW, 21, W4_SYNACK_SEQ);
  u32(M4_FLOW, 22, W4_KA_SEQ); u32(M4_FLOW, 23, W4_KA_ACK);
  set(M4_FLOW, 24, TK_ENT(TK_RVAR_CNT, W4_ACL_N, 0, L4E_ACL));
  u32(M4_FLOW, 25, W4_DIRECTION_SCORE);
  set(M4_FLOW, 26, TK_ENT(TK_STR2W, W4_DOMAIN_OFF, 0, 0));
  set(M4_FLOW, 27, TK_ENT(TK_RVAR_DIGITS, W4_AGG_N, 0, L4E_AGG));
  u32(M4_FLOW, 28, W4_INIT_IPID);
  // FlowKey
  u32(M4_KEY, 1, W4_VTAP_ID); u32(M4_KEY, 2, W4_TAP_TYPE); u64(M4_KEY, 3, W4_TAP_PORT); u64(M4_KEY, 4, W4_MAC_SRC);
  u64(M4_KEY, 5, W4_MAC_DST); u32(M4_KEY, 6, W4_IP_SRC); u32(M4_KEY, 7, W4_IP_DST);
  set(M4_KEY, 8, TK_ENT(TK_STR2W, W4_IP6_SRC_OFF, 0, 0)); set(M4_KEY, 9, TK_ENT(TK_STR2W, W4_IP6_DST_OFF, 0, 0));
  u32(M4_KEY, 10, W4_PORT_SRC); u32(M4_KEY, 11, W4_PORT_DST); u32(M4_KEY, 12, W4_PROTO);
  // FlowMetricsPeer x 2
  for (int s = 0; s < 2; s++) {
    const uint32_t m = s ? M4_PEER1 : M4_PEER0;
    const int b = s ? W4_PEER1 : W4_PEER0;
    u64(m, 1, b + P4_BYTE); u64(m, 2, b + P4_L3_BYTE); u64(m, 3, b + P4_L4_BYTE); u64(m, 4, b + P4_PACKET);
    u64(m, 5, b + P4_TOTAL_BYTE); u64(m, 6, b + P4_TOTAL_PACKET); u64(m, 7, W4_DUMMY); u64(m, 8, W4_DUMMY);
    u32(m, 9, b + P4_TCP_FLAGS); u32(m, 10, b + P4_L3_EPC); u32(m, 11, b + P4_L2_END); u32(m, 12, b + P4_L3_END);
    u32(m, 13, W4_DUMMY); u32(m, 14, W4_DUMMY); u32(m, 15, b + P4_VIP_IF); u32(m, 16, W4_DUMMY);
    u32(m, 20, b + P4_REAL_IP); u32(m, 21, b + P4_REAL_PORT); u32(m, 22, b + P4_GPID);
  }
  // TunnelField
  for (uint32_t f = 1; f <= 12; f++) u32(M4_TUNNEL, f, W4_TX_IP0 + (int)f - 1);
  u32(M4_TUNNEL, 13, W4_DUMMY);
  // FlowPerfStats
  msg(M4_PERF, 1, M4_TCP); msg(M4_PERF, 2, M4_L7PERF); u32(M4_PERF, 3, W4_DUMMY); u32(M4_PERF, 4, W4_L7_PROTOCOL);
  u32(M4_PERF, 5, W4_L7_FAILED);
  // TCPPerfStats
  u32(M4_TCP, 1, W4_RTT_CLIENT_MAX); u32(M4_TCP, 2, W4_RTT_SERVER_MAX); u32(M4_TCP, 3, W4_SRT_MAX); u32(M4_TCP, 4, W4_ART_MAX);
  u32(M4_TCP, 5, W4_RTT); u32(M4_TCP, 8, W4_SRT_SUM); u32(M4_TCP, 9, W4_ART_SUM); u32(M4_TCP, 12, W4_SRT_COUNT);
  u32(M4_TCP, 13, W4_ART_COUNT); msg(M4_TCP, 14, M4_CNT_TX); msg(M4_TCP, 15, M4_CNT_RX); u32(M4_TCP, 16, W4_DUMMY);
  u32(M4_TCP, 17, W4_SYN_COUNT); u32(M4_TCP, 18, W4_SYNACK_COUNT); u32(M4_TCP, 19, W4_CIT_MAX); u32(M4_TCP, 20, W4_CIT_SUM);
  u32(M4_TCP, 21, W4_CIT_COUNT); u32(M4_TCP, 22, W4_FIN_COUNT);
  u32(M4_CNT_TX, 1, W4_TX_RETRANS); u32(M4_CNT_TX, 2, W4_TX_ZERO_WIN); u32(M4_CNT_TX, 3, W4_TX_OOO);
  u32(M4_CNT_RX, 1, W4_RX_RETRANS); u32(M4_CNT_RX, 2, W4_RX_ZERO_WIN); u32(M4_CNT_RX, 3, W4_RX_OOO);
  // L7PerfStats
  u32(M4_L7PERF, 1, W4_REQUEST_COUNT); u32(M4_L7PERF, 2, W4_RESPONSE_COUNT); u32(M4_L7PERF, 3, W4_ERR_CLIENT);
  u32(M4_L7PERF, 4, W4_ERR_SERVER); u32(M4_L7PERF, 5, W4_ERR_TIMEOUT); u32(M4_L7PERF, 6, W4_RRT_COUNT);
  u64(M4_L7PERF, 7, W4_RRT_SUM); u32(M4_L7PERF, 8, W4_RRT_MAX); u32(M4_L7PERF, 9, W4_TLS_RTT);
  return t;
}

}  // namespace dfgpu
