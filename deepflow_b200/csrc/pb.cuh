/*
 * pb.cuh — device-side protobuf wire primitives.  One thread walks one record.
 * Acceptance rules are those of the gogo/protobuf v1.3.2 generated Unmarshal the
 * reference calls (server/libs/codec/simple_codec.go:241): varints longer than 10
 * bytes, truncated fields, a known field with the wrong wire type, field number
 * <= 0, a stray end-group tag and wire types 6/7 are errors; unknown fields are
 * skipped (groups by nesting depth); uint32/int32 fields keep the low 32 bits.
 */
#pragma once
#include <cstdint>

namespace dfgpu {

struct Rd {
  const uint8_t* p;  // start of the current (sub)message
  uint32_t i;        // cursor
  uint32_t l;        // length of the current (sub)message
  // 8-byte register window over the input.  ncu (profiles/r01_*): the decode passes were bound by
  // the long-scoreboard wait behind every LDG.U8 of the varint loops (one dependent L1 round trip
  // per wire byte).  Reading the aligned 8-byte block once and shifting bytes out of a register
  // cuts the dependent loads by ~8x.  The block never crosses the 8-byte granule that holds a
  // valid byte, so it stays inside the (>= 16-byte granular) device allocation.
  uint64_t w;        // bytes of the aligned block at address wblk
  uint64_t wblk;     // 0 = no block loaded (aggregate initialisation `Rd{p, 0, n}` leaves it 0)
};

__device__ __forceinline__ uint8_t ldb(const uint8_t* p) { return __ldg(p); }

// the byte at the cursor (no bounds check), cursor advanced
__device__ __forceinline__ uint8_t rd_byte(Rd& r) {
  const uint64_t addr = (uint64_t)(r.p + r.i);
  const uint64_t blk = addr & ~7ull;
  if (blk != r.wblk) {
    r.w = __ldg(reinterpret_cast<const unsigned long long*>(blk));
    r.wblk = blk;
  }
  r.i++;
  return (uint8_t)(r.w >> ((uint32_t)(addr & 7) * 8));
}

__device__ __forceinline__ bool rd_varint(Rd& r, uint64_t& out) {
  uint64_t v = 0;
#pragma unroll 1
  for (uint32_t shift = 0;; shift += 7) {
    if (shift >= 64) return false;
    if (r.i >= r.l) return false;
    uint8_t b = rd_byte(r);
    v |= (uint64_t)(b & 0x7f) << shift;
    if (b < 0x80) break;
  }
  out = v;
  return true;
}

// tag -> field number (>0) and wire type; false on error
__device__ __forceinline__ bool rd_tag(Rd& r, uint32_t& field, uint32_t& wt) {
  uint64_t w;
  if (!rd_varint(r, w)) return false;
  wt = (uint32_t)(w & 7);
  int32_t f = (int32_t)(w >> 3);
  if (wt == 4) return false;
  if (f <= 0) return false;
  field = (uint32_t)f;
  return true;
}

// length prefix of a wire-type-2 field; leaves the cursor at the payload
__device__ __forceinline__ bool rd_len(Rd& r, uint32_t& len) {
  uint64_t n;
  if (!rd_varint(r, n)) return false;
  if ((int64_t)n < 0) return false;
  if (n > (uint64_t)(r.l - r.i)) return false;
  len = (uint32_t)n;
  return true;
}

// skipFlowLog(): skip one unknown field whose tag (wire type wt) was just read
__device__ inline bool rd_skip(Rd& r, uint32_t wt) {
  int depth = 0;
#pragma unroll 1
  for (;;) {
    switch (wt) {
      case 0: {
        uint32_t shift = 0;
        for (;; shift += 7) {
          if (shift >= 64) return false;
          if (r.i >= r.l) return false;
          if (rd_byte(r) < 0x80) break;
        }
        break;
      }
      case 1:
        if (r.l - r.i < 8) return false;
        r.i += 8;
        break;
      case 2: {
        uint64_t n;
        if (!rd_varint(r, n)) return false;
        if ((int64_t)n < 0) return false;
        if (n > (uint64_t)(r.l - r.i)) return false;
        r.i += (uint32_t)n;
        break;
      }
      case 3: depth++; break;
      case 4:
        if (depth == 0) return false;
        depth--;
        break;
      case 5:
        if (r.l - r.i < 4) return false;
        r.i += 4;
        break;
      default: return false;
    }
    if (depth == 0) return true;
    // inside a group: next tag (any field number)
    uint64_t w;
    if (!rd_varint(r, w)) return false;
    wt = (uint32_t)(w & 7);
  }
}

#define PB_U32(dst)                        \
  {                                        \
    uint64_t _v;                           \
    if (wt != 0 || !rd_varint(r, _v)) return false; \
    (dst) = (uint32_t)_v;                  \
    break;                                 \
  }
#define PB_U64(dst)                        \
  {                                        \
    uint64_t _v;                           \
    if (wt != 0 || !rd_varint(r, _v)) return false; \
    (dst) = _v;                            \
    break;                                 \
  }
#define PB_I32(dst)                        \
  {                                        \
    uint64_t _v;                           \
    if (wt != 0 || !rd_varint(r, _v)) return false; \
    (dst) = (int32_t)(uint32_t)_v;         \
    break;                                 \
  }
// string/bytes: record absolute offset (relative to `base`) and length
#define PB_STR(dst_off, dst_len)           \
  {                                        \
    uint32_t _n;                           \
    if (wt != 2 || !rd_len(r, _n)) return false; \
    (dst_off) = (uint32_t)(r.p + r.i - base); \
    (dst_len) = _n;                        \
    r.i += _n;                             \
    break;                                 \
  }
#define PB_SKIP()                          \
  {                                        \
    if (!rd_skip(r, wt)) return false;     \
    break;                                 \
  }

}  // namespace dfgpu
