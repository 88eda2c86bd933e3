/*
 * l7_dev.h — device-side state shared by the L7 kernels: decode (l7_pass.cu), strings
 * (l7_strings.cu), SmartEncoding (l7_kg.cu) and the flow_tag stage (l7.cu).
 */
#pragma once
#include "ctx.h"
#include "l7_decode.cuh"
#include "l7_row.cuh"
#include "tagset.cuh"

namespace dfgpu {

// final references of a record, written by the decode kernel's tail: [L7REF_COUNT][rec_cap] u64
// (offset | length << 32; REF_CS in the offset: constant blob).  0..21 = string streams L7V_*.
enum : int {
  L7REF_EXT_RANGE = 22,  // the one ext_info occurrence carrying the repeated fields (len 0: none / slow)
  L7REF_UA = 23, L7REF_REFERER = 24, L7REF_RPC = 25,  // pseudo attributes (:418-429)
  L7REF_IP6_SRC = 26, L7REF_IP6_DST = 27,
  L7REF_COUNT = 28
};
// per-record array sizes: [L7ARR_COUNT][rec_cap] u32
enum : int { L7ARR_AN_E = 0, L7ARR_AN_B, L7ARR_AV_E, L7ARR_AV_B, L7ARR_MN_E, L7ARR_MN_B, L7ARR_MV_E, L7ARR_COUNT };
// meta word of a record
constexpr uint32_t L7META_OK = 1u, L7META_HAS_EXT = 2u, L7META_SLOW_ARRAYS = 4u;  // bits 8..15: tap-side code
// SmartEncoding inputs: [L7KG_WORDS][rec_cap] u32
enum : int {
  L7KG_VTAP_SIDE_PROTO = 0,  // vtap_id | tap_side << 16 | protocol << 24
  L7KG_PORT_FLAGS,           // port_dst | v6 << 16 | vip_src << 17 | vip_dst << 18
  L7KG_IP_SRC, L7KG_IP_DST, L7KG_EPC_SRC, L7KG_EPC_DST, L7KG_MAC_SRC_LO, L7KG_MAC_SRC_HI, L7KG_MAC_DST_LO,
  L7KG_MAC_DST_HI, L7KG_GPID_0, L7KG_GPID_1, L7KG_POD_0, L7KG_POD_1, L7KG_PID_0, L7KG_PID_1,
  L7KG_WORDS
};

struct L7Dev {
  L7Tables tb;
  uint32_t tok_tbl[M7_COUNT * L7_MAXF];  // pb_tok table (l7_tok.h)
  FtEntry* ft;
  TagSetState* fts;
  uint32_t ft_mask;
  uint32_t ft_timeout;
  uint32_t epoch;
  uint16_t org_id;
  uint8_t* ft_arena;
  uint64_t ft_arena_bytes;
  uint8_t* row_flag;  // [rows] set by flow_tag phase 0: the row has a new/expired key
  // ---- decode kernel -> strings / SmartEncoding kernels (all indexed by record, stride rec_cap) ----
  uint64_t* raw_scratch;  // [resident warps][R7_COUNT][32]: wire references while a record is being decoded
  uint64_t* refs;         // [L7REF_COUNT][rec_cap]
  uint64_t* elems;        // [rec_cap][L7E_COUNT] elements of the repeated ext_info fields (l7_tok.h)
  uint32_t* arr;          // [L7ARR_COUNT][rec_cap]
  uint32_t* meta;         // [rec_cap]
  uint64_t* thash;        // [rec_cap] DJB hash of the row's trace id
  uint32_t* kgin;         // [L7KG_WORDS][rec_cap]
  uint32_t* rec_row;      // [rec_cap] row of the record in l7_flow_log (0xffffffff: none)
  uint32_t rec_cap;
};

unsigned l7_decode_warps_per_cta();  // l7_pass.cu

__device__ __forceinline__ SRef cs_ref(const L7Dev* d, uint32_t id) {
  uint32_t a = d->tb.cs_off[id], b = d->tb.cs_off[id + 1];
  return SRef{d->tb.cs_blob + a, b - a};
}

}  // namespace dfgpu
