/*
 * pb_interp.cuh — table-driven protobuf decoder for the device.
 *
 * The first version of the kernels decoded with one hand-written `switch` per
 * message (l4_decode.cuh / l7_decode.cuh).  ncu showed that shape to be the problem,
 * not the memory system: every lane of a warp sits at a different field of a
 * different record, so the switch bodies serialise (9 of 32 lanes active on average,
 * 255 registers, 10 % occupancy — profiles/ncu_l7_r01.md).  Here every lane runs the
 * SAME short loop — read tag, read varint or length, look the (message, field) pair
 * up in a small table, store into a per-thread slot array — so the warp stays
 * converged and the decoded record lives in local memory slots instead of registers.
 *
 * Acceptance rules are unchanged (pb.cuh): they are the gogo/protobuf v1.3.2
 * generated-Unmarshal rules the reference relies on (simple_codec.go:241).
 */
#pragma once
#include "pb.cuh"

namespace dfgpu {

enum PbKind : uint32_t {
  PK_SKIP = 0,       // unknown field: skipped
  PK_U32 = 1,        // uint32 / int32 varint: low 32 bits kept
  PK_U64 = 2,        // uint64 varint
  PK_STR = 3,        // string / bytes: slot = offset | len << 32 (offset from the record start)
  PK_MSG = 4,        // nested message: slot = occurrence count, slot+1 = last (offset | len << 32)
  PK_RSTR_CNT = 5,   // repeated string: slot = count | bytes << 32
  PK_RSTR_FIRST2 = 6,  // repeated string: slot, slot+1 = first two refs, slot+2 = count
  PK_RDBL_CNT = 7,   // repeated double (packed or not): slot = count
  PK_RVAR_CNT = 8,   // repeated varint (packed or not): slot = count
  PK_RVAR_DIGITS = 9,  // repeated uint64: slot = count | decimal chars incl. ',' separators << 32
};
#define PB_ENT(kind, slot, sub) ((uint32_t)(kind) | (uint32_t)(slot) << 8 | (uint32_t)(sub) << 16)

constexpr int PB_MAX_DEPTH = 6;

__device__ __forceinline__ uint32_t pb_digits(uint64_t v) {
  uint32_t n = 1;
  while (v >= 10) {
    v /= 10;
    n++;
  }
  return n;
}

// Decodes rec[0..len) (top-level message id 0) into slots[].  tbl is [nmsg][maxf].
// Returns false on any error the generated Unmarshal would return.
// SW: the slot store — a private uint64_t array, or a view that writes straight into a slot-major
// global buffer (anything whose operator[] yields a uint64_t lvalue).
// uvarint size of a length prefix (ClickHouse native String encoding)
__device__ __forceinline__ uint32_t pb_vsize(uint32_t n) {
  return 1u + (n >= 0x80u) + (n >= 0x4000u) + (n >= 0x200000u) + (n >= 0x10000000u);
}
// len_prefix_mask: ~0u when repeated strings are to be SIZED in native encoding (their byte count then
// includes one uvarint length prefix per element, dfgpu_cfg.native_strings), else 0.
template <class SW>
__device__ __forceinline__ bool pb_interp(const uint8_t* rec, uint32_t len, const uint32_t* __restrict__ tbl,
                                          uint32_t maxf, SW& slots, const uint32_t len_prefix_mask = 0u) {
  uint32_t stack_end[PB_MAX_DEPTH];
  uint32_t stack_msg[PB_MAX_DEPTH];
  uint32_t depth = 0, msg = 0, end = len;
  Rd r{rec, 0, len};
  for (;;) {
    while (r.i >= end) {  // end of the current (sub)message
      if (depth == 0) return true;
      depth--;
      end = stack_end[depth];
      msg = stack_msg[depth];
    }
    r.l = end;  // every read is bounded by the enclosing message, like dAtA[iNdEx:postIndex]
    uint64_t w;
    if (!rd_varint(r, w)) return false;
    const uint32_t wt = (uint32_t)(w & 7);
    const int32_t f = (int32_t)(w >> 3);
    if (wt == 4 || f <= 0) return false;
    const uint32_t ent = (uint32_t)f < maxf ? tbl[msg * maxf + (uint32_t)f] : 0u;
    const uint32_t kind = ent & 0xff, slot = (ent >> 8) & 0xff;
    if (wt == 0) {
      uint64_t v;
      if (!rd_varint(r, v)) return false;
      if (kind == PK_U32) slots[slot] = (uint32_t)v;
      else if (kind == PK_U64) slots[slot] = v;
      else if (kind == PK_RVAR_CNT) slots[slot] += 1;
      else if (kind == PK_RVAR_DIGITS) {
        uint64_t s = slots[slot];
        slots[slot] = s + 1 + ((uint64_t)(pb_digits(v) + ((uint32_t)s ? 1u : 0u)) << 32);
      } else if (kind != PK_SKIP) return false;  // known field, wrong wire type
    } else if (wt == 2) {
      uint32_t n;
      if (!rd_len(r, n)) return false;
      const uint32_t off = r.i;
      if (kind == PK_MSG) {
        if (depth >= PB_MAX_DEPTH) return false;
        slots[slot] += 1;
        slots[slot + 1] = (uint64_t)off | (uint64_t)n << 32;
        stack_end[depth] = end;
        stack_msg[depth] = msg;
        depth++;
        end = off + n;
        msg = ent >> 16;
        continue;
      }
      if (kind == PK_STR) slots[slot] = (uint64_t)off | (uint64_t)n << 32;
      else if (kind == PK_RSTR_CNT) slots[slot] += 1 + ((uint64_t)(n + (pb_vsize(n) & len_prefix_mask)) << 32);
      else if (kind == PK_RSTR_FIRST2) {
        uint64_t c = slots[slot + 2];
        if (c < 2) slots[slot + c] = (uint64_t)off | (uint64_t)n << 32;
        slots[slot + 2] = c + 1;
      } else if (kind == PK_RVAR_CNT || kind == PK_RVAR_DIGITS) {
        const uint32_t pend = off + n;
        while (r.i < pend) {  // packed: varints bounded by the message, loop while before postIndex
          uint64_t v;
          if (!rd_varint(r, v)) return false;
          if (kind == PK_RVAR_CNT) slots[slot] += 1;
          else {
            uint64_t s = slots[slot];
            slots[slot] = s + 1 + ((uint64_t)(pb_digits(v) + ((uint32_t)s ? 1u : 0u)) << 32);
          }
        }
        continue;
      } else if (kind == PK_RDBL_CNT) {
        const uint32_t pend = off + n;
        while (r.i < pend) {
          if (r.l - r.i < 8) return false;
          r.i += 8;
          slots[slot] += 1;
        }
        continue;
      } else if (kind != PK_SKIP) return false;
      r.i = off + n;
    } else if (wt == 1) {
      if (kind == PK_RDBL_CNT) {
        if (r.l - r.i < 8) return false;
        r.i += 8;
        slots[slot] += 1;
      } else if (kind == PK_SKIP) {
        if (r.l - r.i < 8) return false;
        r.i += 8;
      } else return false;
    } else {
      if (kind != PK_SKIP) return false;
      if (!rd_skip(r, wt)) return false;  // fixed32, groups, illegal wire types
    }
  }
}

// S: anything indexable by slot (the private array, or a view of the slot-major buffer)
template <class S>
__device__ __forceinline__ uint32_t sl_u32(const S& s, int i) { return (uint32_t)s[i]; }
template <class S>
__device__ __forceinline__ int32_t sl_i32(const S& s, int i) { return (int32_t)(uint32_t)s[i]; }
template <class S>
__device__ __forceinline__ uint32_t sl_off(const S& s, int i) { return (uint32_t)s[i]; }
template <class S>
__device__ __forceinline__ uint32_t sl_len(const S& s, int i) { return (uint32_t)(s[i] >> 32); }

}  // namespace dfgpu
