/*
 * l7_tok.h — word / reference layout + tokenizer table of flow_log.AppProtoLogsData
 * (message/flow_log.proto:179-310) for pb_tok.cuh.  Words are 4 bytes (shared memory, one column per
 * thread); a uint64 field takes two consecutive words (lo, hi).
 */
#pragma once
#include <cstdint>
#include <vector>

#include "pb_tok.cuh"

namespace dfgpu {

enum L7Msg : uint32_t { M7_TOP = 0, M7_BASE, M7_HEAD, M7_REQ, M7_RESP, M7_TRACE, M7_EXT, M7_COUNT };
constexpr uint32_t L7_MAXF = 48;  // field numbers < 48 (AppProtoLogsBaseInfo goes up to 43)

enum L7Word : int {
  // ---- uint32 ----
  W7_VTAP_ID, W7_TAP_TYPE, W7_IS_IPV6, W7_TAP_SIDE, W7_IP_SRC, W7_IP_DST, W7_EPC_SRC, W7_EPC_DST, W7_PORT_SRC,
  W7_PORT_DST, W7_PROTOCOL, W7_VIP_SRC, W7_VIP_DST, W7_REQ_TCP_SEQ, W7_RESP_TCP_SEQ, W7_PID_0, W7_PID_1,
  W7_THREAD_0, W7_THREAD_1, W7_CAP_SEQ_0, W7_CAP_SEQ_1, W7_GPID_0, W7_GPID_1, W7_POD_0, W7_POD_1, W7_BIZ_TYPE,
  W7_HEAD_PROTO, W7_HEAD_MSG_TYPE, W7_REQ_LEN, W7_RESP_LEN, W7_ROW_EFFECT, W7_DIRECTION_SCORE, W7_FLAGS,
  W7_CAP_REQ_BYTE, W7_CAP_RESP_BYTE, W7_RESP_STATUS, W7_RESP_CODE, W7_REQUEST_ID,
  // ---- occurrence / element counters ----
  W7_BASE_CNT, W7_HEAD_CNT, W7_REQ_CNT, W7_RESP_CNT, W7_TRACE_CNT, W7_EXT_CNT, W7_TRACE_IDS_N,
  W7_ATTR_NAMES, W7_ATTR_NAMES_B, W7_ATTR_VALUES, W7_ATTR_VALUES_B, W7_METRICS_NAMES, W7_METRICS_NAMES_B,
  W7_METRICS_VALUES, W7_METRICS_VALUES_OCC,
  // ---- uint64 (lo, hi) ----
  W7_START_TIME, W7_START_TIME_HI, W7_END_TIME, W7_END_TIME_HI, W7_FLOW_ID, W7_FLOW_ID_HI, W7_TAP_PORT,
  W7_TAP_PORT_HI, W7_MAC_SRC, W7_MAC_SRC_HI, W7_MAC_DST, W7_MAC_DST_HI, W7_SYS_REQ, W7_SYS_REQ_HI, W7_SYS_RESP,
  W7_SYS_RESP_HI, W7_CORO_0, W7_CORO_0_HI, W7_CORO_1, W7_CORO_1_HI, W7_HEAD_RRT, W7_HEAD_RRT_HI,
  W7_COUNT
};

// string / bytes fields as they come off the wire ("raw" references)
enum L7Raw : int {
  R7_IP6_SRC, R7_IP6_DST, R7_KNAME_0, R7_KNAME_1, R7_VERSION, R7_BIZ_CODE, R7_BIZ_SCENARIO, R7_BIZ_RESPONSE_CODE,
  R7_REQ_TYPE, R7_DOMAIN, R7_RESOURCE, R7_ENDPOINT, R7_EXCEPTION, R7_RESULT, R7_TRACE_ID, R7_SPAN_ID,
  R7_PARENT_SPAN_ID, R7_TRACE_IDS_0, R7_TRACE_IDS_1, R7_SERVICE_NAME, R7_CLIENT_IP, R7_X_REQUEST_ID_0,
  R7_USER_AGENT, R7_REFERER, R7_RPC_SERVICE, R7_PROTOCOL_STR, R7_X_REQUEST_ID_1, R7_EXT_RANGE,
  R7_COUNT
};
static_assert(R7_COUNT <= 32, "presence mask of the raw references is one 32-bit word");

// elements of the repeated ExtendedInfo fields, per record: [L7E_COUNT] references.  A record with more
// elements than a block holds (or with its repeated fields split over several ext_info occurrences) takes the
// slow path of the arrays kernel (a walk over the record).
enum : uint32_t { L7E_ATTR_NAMES = 0, L7E_ATTR_VALUES = 8, L7E_METRICS_NAMES = 16, L7E_METRICS_VALUES = 20, L7E_COUNT = 21 };
constexpr uint32_t L7E_CAP_ATTR = 8, L7E_CAP_METRICS = 4;

inline std::vector<uint32_t> l7_tok_table() {
  std::vector<uint32_t> t(M7_COUNT * L7_MAXF, 0u);
  auto set = [&](uint32_t m, uint32_t f, uint32_t ent) { t[m * L7_MAXF + f] = ent; };
  auto u32 = [&](uint32_t m, uint32_t f, int w) { set(m, f, TK_ENT(TK_U32, w, 0, 0)); };
  auto u64 = [&](uint32_t m, uint32_t f, int w) { set(m, f, TK_ENT(TK_U64, w, 0, 0)); };
  auto str = [&](uint32_t m, uint32_t f, int r) { set(m, f, TK_ENT(TK_STR, 0, 0, r)); };
  // AppProtoLogsData
  set(M7_TOP, 1, TK_ENT(TK_MSG, W7_BASE_CNT, M7_BASE, 0));
  u32(M7_TOP, 9, W7_REQ_LEN);
  u32(M7_TOP, 10, W7_RESP_LEN);
  set(M7_TOP, 11, TK_ENT(TK_MSG, W7_REQ_CNT, M7_REQ, 0));
  set(M7_TOP, 12, TK_ENT(TK_MSG, W7_RESP_CNT, M7_RESP, 0));
  str(M7_TOP, 13, R7_VERSION);
  set(M7_TOP, 14, TK_ENT(TK_MSG, W7_TRACE_CNT, M7_TRACE, 0));
  set(M7_TOP, 15, TK_ENT(TK_MSG, W7_EXT_CNT, M7_EXT, R7_EXT_RANGE) | TKF_RANGE);
  u32(M7_TOP, 16, W7_ROW_EFFECT);
  u32(M7_TOP, 17, W7_DIRECTION_SCORE);
  u32(M7_TOP, 18, W7_FLAGS);
  u32(M7_TOP, 19, W7_CAP_REQ_BYTE);
  u32(M7_TOP, 20, W7_CAP_RESP_BYTE);
  str(M7_TOP, 21, R7_BIZ_CODE);
  str(M7_TOP, 22, R7_BIZ_SCENARIO);
  str(M7_TOP, 23, R7_BIZ_RESPONSE_CODE);
  // AppProtoLogsBaseInfo
  u64(M7_BASE, 1, W7_START_TIME); u64(M7_BASE, 2, W7_END_TIME); u64(M7_BASE, 3, W7_FLOW_ID); u64(M7_BASE, 4, W7_TAP_PORT);
  u32(M7_BASE, 5, W7_VTAP_ID); u32(M7_BASE, 6, W7_TAP_TYPE); u32(M7_BASE, 7, W7_IS_IPV6); u32(M7_BASE, 8, W7_TAP_SIDE);
  set(M7_BASE, 9, TK_ENT(TK_MSG, W7_HEAD_CNT, M7_HEAD, 0));
  u64(M7_BASE, 10, W7_MAC_SRC); u64(M7_BASE, 11, W7_MAC_DST); u32(M7_BASE, 12, W7_IP_SRC); u32(M7_BASE, 13, W7_IP_DST);
  str(M7_BASE, 14, R7_IP6_SRC); str(M7_BASE, 15, R7_IP6_DST); u32(M7_BASE, 16, W7_EPC_SRC); u32(M7_BASE, 17, W7_EPC_DST);
  u32(M7_BASE, 18, W7_PORT_SRC); u32(M7_BASE, 19, W7_PORT_DST); u32(M7_BASE, 20, W7_PROTOCOL);
  u32(M7_BASE, 21, W7_VIP_SRC); u32(M7_BASE, 22, W7_VIP_DST); u32(M7_BASE, 23, W7_REQ_TCP_SEQ);
  u32(M7_BASE, 24, W7_RESP_TCP_SEQ); u32(M7_BASE, 25, W7_PID_0); u32(M7_BASE, 26, W7_PID_1);
  str(M7_BASE, 27, R7_KNAME_0); str(M7_BASE, 28, R7_KNAME_1); u64(M7_BASE, 29, W7_SYS_REQ); u64(M7_BASE, 30, W7_SYS_RESP);
  u32(M7_BASE, 31, W7_THREAD_0); u32(M7_BASE, 32, W7_THREAD_1); u32(M7_BASE, 33, W7_CAP_SEQ_0);
  u32(M7_BASE, 34, W7_CAP_SEQ_1); u32(M7_BASE, 35, W7_GPID_0); u32(M7_BASE, 36, W7_GPID_1);
  u64(M7_BASE, 39, W7_CORO_0); u64(M7_BASE, 40, W7_CORO_1); u32(M7_BASE, 41, W7_POD_0); u32(M7_BASE, 42, W7_POD_1);
  u32(M7_BASE, 43, W7_BIZ_TYPE);
  u32(M7_HEAD, 1, W7_HEAD_PROTO); u32(M7_HEAD, 2, W7_HEAD_MSG_TYPE); u64(M7_HEAD, 5, W7_HEAD_RRT);
  str(M7_REQ, 1, R7_REQ_TYPE); str(M7_REQ, 2, R7_DOMAIN); str(M7_REQ, 3, R7_RESOURCE); str(M7_REQ, 4, R7_ENDPOINT);
  u32(M7_RESP, 1, W7_RESP_STATUS); u32(M7_RESP, 2, W7_RESP_CODE); str(M7_RESP, 3, R7_EXCEPTION); str(M7_RESP, 4, R7_RESULT);
  str(M7_TRACE, 1, R7_TRACE_ID); str(M7_TRACE, 2, R7_SPAN_ID); str(M7_TRACE, 3, R7_PARENT_SPAN_ID);
  set(M7_TRACE, 4, TK_ENT(TK_RSTR_FIRST2, W7_TRACE_IDS_N, 0, R7_TRACE_IDS_0));
  str(M7_EXT, 1, R7_SERVICE_NAME); str(M7_EXT, 2, R7_CLIENT_IP); u32(M7_EXT, 3, W7_REQUEST_ID);
  str(M7_EXT, 4, R7_X_REQUEST_ID_0); str(M7_EXT, 6, R7_USER_AGENT); str(M7_EXT, 7, R7_REFERER);
  str(M7_EXT, 8, R7_RPC_SERVICE); str(M7_EXT, 9, R7_PROTOCOL_STR); str(M7_EXT, 10, R7_X_REQUEST_ID_1);
  set(M7_EXT, 16, TK_ENT(TK_RSTR_CNT, W7_ATTR_NAMES, 0, L7E_ATTR_NAMES));
  set(M7_EXT, 17, TK_ENT(TK_RSTR_CNT, W7_ATTR_VALUES, 0, L7E_ATTR_VALUES));
  set(M7_EXT, 18, TK_ENT(TK_RSTR_CNT, W7_METRICS_NAMES, 0, L7E_METRICS_NAMES));
  set(M7_EXT, 19, TK_ENT(TK_RDBL_CNT, W7_METRICS_VALUES, 0, L7E_METRICS_VALUES));
  return t;
}

}  // namespace dfgpu
