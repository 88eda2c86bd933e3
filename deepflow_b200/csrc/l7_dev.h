/*
 * l7_dev.h — device-side state shared by the L7 pass kernels (l7_pass.cu) and the
 * flow_tag stage (l7.cu).
 */
#pragma once
#include "ctx.h"
#include "l7_decode.cuh"
#include "l7_slots.h"
#include "schema_l7.h"
#include "tagset.cuh"

#include "../../include/dfgpu_exception_tables.h"

namespace dfgpu {

// ---- constant strings the path can emit (protocol names, pseudo attribute names,
// exception descriptions): one device blob, addressed by id -------------------------
enum : uint32_t {
  CS_EMPTY = 0, CS_ATTR_UA, CS_ATTR_REFERER, CS_ATTR_RPC, CS_FT_APP_SERVICE, CS_FT_ENDPOINT,
  CS_FT_APP_INSTANCE, CS_PROTO_NA, CS_PROTO_TLS, CS_DNS_AVAILABLE, CS_DNS_PRIVATE, CS_KAFKA_UNKNOWN,
  CS_FIXED_COUNT
};


struct L7Dev {
  const uint8_t* cs_blob;
  const uint32_t* cs_off;  // [n+1]
  uint32_t cs_http, cs_dns, cs_dubbo, cs_kafka, cs_mqtt3, cs_mqtt5;  // base ids of the tables
  uint16_t proto_cs[256][2];  // L7Protocol.String(isTLS) -> cs id
  uint32_t pb_tbl[M7_COUNT * L7_MAXF];  // pb_interp table (l7_slots.h)
  FtEntry* ft;
  uint32_t ft_mask;
  uint32_t ft_timeout;
  uint32_t epoch;
  uint16_t org_id;
  uint8_t* ft_arena;
  uint64_t ft_arena_bytes;
  uint8_t* row_flag;  // [nrows] set by flow_tag phase 0: the row has a new/expired key
  // decoded slots of every record, slot-major ([S7_COUNT + 3][slot_stride], coalesced): written by the
  // SIZE pass so that the EMIT pass does not decode the record a second time; the two extra rows are
  // the presence bitmap of the slots (bit 63 of the second word: the record decoded), the third the DJB
  // hash of the record's trace id
  uint64_t* slotbuf;
  uint32_t slot_stride;
  uint32_t* rec_row;  // [records] row of the record in l7_flow_log (0xffffffff: none), EMIT pass -> l7_kg.cu
};


__device__ __forceinline__ SRef cs_ref(const L7Dev* d, uint32_t id) {
  uint32_t a = d->cs_off[id], b = d->cs_off[id + 1];
  return SRef{d->cs_blob + a, b - a};
}

}  // namespace dfgpu
