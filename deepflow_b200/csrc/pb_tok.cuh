/*
 * pb_tok.cuh — the lean table-driven protobuf tokenizer of the decode kernels.
 *
 * Round 1's interpreter (pb_interp.cuh) wrote 8-byte slots into a slot-major GLOBAL buffer
 * that a second kernel read back: 0.8 GB of round trip per 625 k records and, because the
 * in-flight slot columns of all resident threads (113 MB) did not fit L2, partial-sector
 * evictions that tripled the DRAM traffic (ncu, profiles/r01_*).  Here the decoded record never
 * leaves the SM:
 *   - scalar slots are 4-byte words in SHARED memory, word-major ([word][thread]: bank = lane,
 *     conflict free whatever slot each lane is writing);
 *   - string fields become (offset, length) references in a small per-warp scratch that the same
 *     warp reads back a few microseconds later (L2 resident: 7 KB per resident warp);
 *   - the wire bytes come through an 8-byte register window with the next block prefetched, and
 *     the common case (1-2 byte tag + 1-2 byte varint) is decoded from one 4-byte peek.
 *
 * Acceptance rules are unchanged: those of the gogo/protobuf v1.3.2 generated Unmarshal the
 * reference calls (server/libs/codec/simple_codec.go:241) as restated in pb.cuh / pb_interp.cuh:
 * varints longer than 10 bytes, truncated fields, a known field with the wrong wire type, field
 * number <= 0, a stray end-group tag and wire types 6/7 are errors; unknown fields are skipped
 * (groups by nesting depth); uint32/int32 fields keep the low 32 bits; the last occurrence of a
 * scalar wins; repeated sub-messages merge.
 */
#pragma once
#include <cstdint>

// The decode core (tokenizer + row builders) compiles for the host too: tests/harness builds it with
// plain g++ to check the per-record logic against the oracle without a GPU.  The product library only
// ever runs it on the device.
#if defined(__CUDACC__)
#define DF_HD __host__ __device__ __forceinline__
#define DF_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define DF_HD inline
#define DF_HD_NOINLINE inline
#endif
#if defined(__CUDA_ARCH__)
#define DF_LDG(p) __ldg(p)
#define DF_FUNNEL_R(lo, hi, s) __funnelshift_r(lo, hi, s)
#define DF_BYTE_PERM(a, b, k) __byte_perm(a, b, k)
#else
#define DF_LDG(p) (*(p))
#define DF_FUNNEL_R(lo, hi, s) ((uint32_t)((((uint64_t)(hi) << 32) | (uint64_t)(lo)) >> ((s) & 31u)))
#define DF_BYTE_PERM(a, b, k) ((uint32_t)(((((uint64_t)(b) << 32) | (uint64_t)(a)) >> (((k) & 7u) * 8u)) & 0xffu))
#endif

namespace dfgpu {
struct TokU2 {  // uint2 without the CUDA headers
  uint32_t x, y;
};

enum TokKind : uint32_t {
  TK_SKIP = 0,       // unknown field
  TK_U32 = 1,        // uint32 / int32 varint -> word[slot]
  TK_U64 = 2,        // uint64 varint -> word[slot], word[slot + 1]
  TK_STR = 3,        // string / bytes -> ref[ref]
  TK_MSG = 4,        // nested message: word[slot] += 1 (occurrences); TKF_RANGE: ref[ref] = its byte range
  TK_RSTR_CNT = 5,   // repeated string: word[slot] += 1, word[slot + 1] += bytes (+ length prefixes)
  TK_RSTR_FIRST2 = 6,  // repeated string: ref[ref], ref[ref + 1] = first two, word[slot] = count
  TK_RDBL_CNT = 7,   // repeated double (packed or not): word[slot] += 1 per value
  TK_RVAR_CNT = 8,   // repeated varint (packed or not): word[slot] += 1 per value
  TK_RVAR_DIGITS = 9,  // repeated uint64: word[slot] += 1, word[slot + 1] += decimal chars + ',' separators
};
constexpr uint32_t TKF_RANGE = 1u << 4;
// entry: kind[0:4) | flags[4:8) | slot[8:16) | sub-message[16:20) | ref[20:26)
#define TK_ENT(kind, slot, sub, ref) ((uint32_t)(kind) | (uint32_t)(slot) << 8 | (uint32_t)(sub) << 16 | (uint32_t)(ref) << 20)

constexpr uint32_t TK_MAX_DEPTH = 5;  // nesting below the top-level message (register stack)

// uvarint size of a length prefix (ClickHouse native String encoding)
DF_HD uint32_t tk_vsize(uint32_t n) {
  return 1u + (n >= 0x80u) + (n >= 0x4000u) + (n >= 0x200000u) + (n >= 0x10000000u);
}
DF_HD uint32_t tk_digits(uint64_t v) {
  uint32_t n = 1;
  while (v >= 10) {
    v /= 10;
    n++;
  }
  return n;
}

// ---- byte window over the input arena ---------------------------------------------------------------
// `in` is the staging buffer (device), positions are absolute byte offsets into it.  The window holds the
// aligned 8-byte block of the cursor and the block after it, so that a 4-byte peek never needs a load of
// its own.  last_blk bounds the prefetch inside the buffer.
struct TokWin {
  const TokU2* in8;
  TokU2 cur, nxt;
  uint32_t blk;       // index of `cur`
  uint32_t last_blk;  // last block holding a valid byte of the buffer
  DF_HD void init(const uint8_t* in, uint64_t nbytes) {
    in8 = reinterpret_cast<const TokU2*>(in);
    blk = 0xfffffff0u;
    last_blk = nbytes ? (uint32_t)((nbytes - 1) >> 3) : 0u;
    cur = nxt = TokU2{0, 0};
  }
  static DF_HD TokU2 ld8(const TokU2* p) {
#if defined(__CUDA_ARCH__)
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    return TokU2{v.x, v.y};
#else
    return *p;
#endif
  }
  DF_HD void seek(uint32_t pos) {
    const uint32_t b = pos >> 3;
    if (b != blk) {
      const uint32_t nb = b < last_blk ? b + 1 : last_blk;
      if (b == blk + 1) cur = nxt; else cur = ld8(in8 + b);
      nxt = ld8(in8 + nb);
      blk = b;
    }
  }
  // bytes pos .. pos + 3, little endian (caller guarantees they are inside the buffer or unused)
  DF_HD uint32_t peek4(uint32_t pos) {
    seek(pos);
    const uint32_t k = pos & 7u;
    const uint32_t a = k < 4 ? cur.x : cur.y, b = k < 4 ? cur.y : nxt.x;
    return DF_FUNNEL_R(a, b, (k & 3u) * 8u);
  }
  DF_HD uint32_t byte(uint32_t pos) {
    seek(pos);
    return DF_BYTE_PERM(cur.x, cur.y, pos & 7u) & 0xffu;
  }
};

// varint at pos (< lim), slow path: any length.  false: truncated or longer than 10 bytes.
DF_HD bool tk_varint(TokWin& w, uint32_t& pos, const uint32_t lim, uint64_t& out) {
  uint32_t lo = 0, hi = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (uint32_t shift = 0;; shift += 7) {
    if (shift >= 64 || pos >= lim) return false;
    const uint32_t b = w.byte(pos++);
    const uint32_t v = b & 0x7fu;
    if (shift < 32) {
      lo |= v << shift;
      if (shift > 25) hi |= v >> (32 - shift);
    } else {
      hi |= v << (shift - 32);
    }
    if (b < 0x80u) break;
  }
  out = (uint64_t)hi << 32 | lo;
  return true;
}

// skipFlowLog(): skip one unknown field of wire type wt (3: group start, 5: fixed32, 6/7: illegal) whose
// tag was just read.  Out of line: rare.
constexpr uint32_t TK_FAIL = 0xffffffffu;
// returns the position behind the field, or TK_FAIL.  Uses a window of its own so that the caller's stays
// in registers (nothing of the hot loop has its address taken).
DF_HD_NOINLINE uint32_t tk_skip_slow(const TokU2* in8, uint32_t last_blk, uint32_t pos, const uint32_t lim, uint32_t wt) {
  TokWin w;
  w.in8 = in8;
  w.last_blk = last_blk;
  w.blk = 0xfffffff0u;
  w.cur = w.nxt = TokU2{0, 0};
  int depth = 0;
  for (;;) {
    switch (wt) {
      case 0: {
        uint64_t v;
        if (!tk_varint(w, pos, lim, v)) return TK_FAIL;
        break;
      }
      case 1:
        if (lim - pos < 8) return TK_FAIL;
        pos += 8;
        break;
      case 2: {
        uint64_t n;
        if (!tk_varint(w, pos, lim, n)) return TK_FAIL;
        if ((int64_t)n < 0 || n > (uint64_t)(lim - pos)) return TK_FAIL;
        pos += (uint32_t)n;
        break;
      }
      case 3: depth++; break;
      case 4:
        if (depth == 0) return TK_FAIL;
        depth--;
        break;
      case 5:
        if (lim - pos < 4) return TK_FAIL;
        pos += 4;
        break;
      default: return TK_FAIL;
    }
    if (depth == 0) return pos;
    uint64_t t;  // inside a group: next tag (any field number)
    if (!tk_varint(w, pos, lim, t)) return TK_FAIL;
    wt = (uint32_t)(t & 7);
  }
}

// Sink interface (all __device__ __forceinline__):
//   void u32(slot, v); void u64(slot, lo, hi); void add(slot, v)           scalar words
//   uint32_t word(slot)                                                     read a word back
//   void ref(r, off, len)                                                   string reference (absolute offset)
// Decodes the record at [rec_off, rec_off + len) of the buffer.  tbl: [nmsg][maxf] entries in shared memory.
template <uint32_t MAXF, class Sink>
DF_HD bool tk_decode(TokWin& w, const uint32_t rec_off, const uint32_t len,
                                          const uint32_t* __restrict__ tbl, Sink& sink,
                                          const uint32_t len_prefix_mask) {
  constexpr uint32_t maxf = MAXF;
  uint32_t pos = rec_off, end = rec_off + len;
  // register stack of the enclosing messages (shift register: no dynamic indexing)
  uint32_t e1 = 0, e2 = 0, e3 = 0, e4 = 0, e5 = 0, msgs = 0, depth = 0, msg = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (;;) {
    while (pos >= end) {  // end of the current (sub)message
      if (depth == 0) return true;
      depth--;
      end = e1; e1 = e2; e2 = e3; e3 = e4; e4 = e5;
      msg = msgs & 15u;
      msgs >>= 4;
    }
    // ---- tag + first varint: one 4-byte peek covers the common case ----
    uint32_t wt, f;
    uint32_t vlo = 0, vhi = 0;  // value varint (wire type 0) or length (wire type 2)
    bool have_v = false;
    if (end - pos >= 4) {
      uint32_t x = w.peek4(pos);
      uint32_t t = x & 0x7fu, nb = 1;
      if (x & 0x80u) {
        t |= (x >> 1) & 0x3f80u;
        nb = 2;
      }
      if (nb == 2 && (x & 0x8000u)) {
        // tag of 3+ bytes (field number >= 2048): generic path
        uint64_t tv;
        if (!tk_varint(w, pos, end, tv)) return false;
        wt = (uint32_t)(tv & 7);
        const int32_t fi = (int32_t)(tv >> 3);
        if (wt == 4 || fi <= 0) return false;
        f = (uint32_t)fi;
      } else {
        wt = t & 7u;
        f = t >> 3;
        if (wt == 4 || f == 0) return false;
        pos += nb;
        x >>= 8 * nb;  // 2 or 3 value bytes left in x
        if ((wt == 0 || wt == 2)) {
          if (!(x & 0x80u)) {
            vlo = x & 0x7fu;
            pos += 1;
            have_v = true;
          } else if (!(x & 0x8000u)) {
            vlo = (x & 0x7fu) | ((x >> 1) & 0x3f80u);
            pos += 2;
            have_v = true;
          }
        }
      }
    } else {
      uint64_t tv;
      if (!tk_varint(w, pos, end, tv)) return false;
      wt = (uint32_t)(tv & 7);
      const int32_t fi = (int32_t)(tv >> 3);
      if (wt == 4 || fi <= 0) return false;
      f = (uint32_t)fi;
    }
    const uint32_t ent = f < maxf ? tbl[msg * maxf + f] : 0u;
    const uint32_t kind = ent & 15u, slot = (ent >> 8) & 0xffu;
    if (wt == 0) {
      if (!have_v) {
        uint64_t v;
        if (!tk_varint(w, pos, end, v)) return false;
        vlo = (uint32_t)v;
        vhi = (uint32_t)(v >> 32);
      }
      if (kind == TK_U32) sink.u32(slot, vlo);
      else if (kind == TK_U64) sink.u64(slot, vlo, vhi);
      else if (kind == TK_RVAR_CNT) sink.add(slot, 1);
      else if (kind == TK_RVAR_DIGITS) {
        const uint32_t c = sink.word(slot);
        sink.u32(slot, c + 1);
        sink.add(slot + 1, tk_digits((uint64_t)vhi << 32 | vlo) + (c ? 1u : 0u));
      } else if (kind != TK_SKIP) return false;  // known field, wrong wire type
    } else if (wt == 2) {
      if (!have_v) {
        uint64_t n;
        if (!tk_varint(w, pos, end, n)) return false;
        if ((int64_t)n < 0) return false;
        vlo = (uint32_t)n;
        vhi = (uint32_t)(n >> 32);
      }
      if (vhi != 0 || vlo > end - pos) return false;
      const uint32_t n = vlo, off = pos;
      if (kind == TK_MSG) {
        if (depth >= TK_MAX_DEPTH) return false;
        sink.add(slot, 1);
        if (ent & TKF_RANGE) sink.ref((ent >> 20) & 63u, off, n);
        e5 = e4; e4 = e3; e3 = e2; e2 = e1; e1 = end;
        msgs = msgs << 4 | msg;
        depth++;
        end = off + n;
        msg = (ent >> 16) & 15u;
        continue;
      }
      if (kind == TK_STR) sink.ref((ent >> 20) & 63u, off, n);
      else if (kind == TK_RSTR_CNT) {
        sink.add(slot, 1);
        sink.add(slot + 1, n + (tk_vsize(n) & len_prefix_mask));
      } else if (kind == TK_RSTR_FIRST2) {
        const uint32_t c = sink.word(slot);
        if (c < 2) sink.ref(((ent >> 20) & 63u) + c, off, n);
        sink.u32(slot, c + 1);
      } else if (kind == TK_RVAR_CNT || kind == TK_RVAR_DIGITS) {
        const uint32_t pend = off + n;
        while (pos < pend) {  // packed: varints bounded by the enclosing message, loop while before postIndex
          uint64_t v;
          if (!tk_varint(w, pos, end, v)) return false;
          if (kind == TK_RVAR_CNT) sink.add(slot, 1);
          else {
            const uint32_t c = sink.word(slot);
            sink.u32(slot, c + 1);
            sink.add(slot + 1, tk_digits(v) + (c ? 1u : 0u));
          }
        }
        continue;
      } else if (kind == TK_RDBL_CNT) {
        const uint32_t pend = off + n;
        while (pos < pend) {
          if (end - pos < 8) return false;
          pos += 8;
          sink.add(slot, 1);
        }
        continue;
      } else if (kind != TK_SKIP) return false;
      pos = off + n;
    } else if (wt == 1) {
      if (kind != TK_RDBL_CNT && kind != TK_SKIP) return false;
      if (end - pos < 8) return false;
      pos += 8;
      if (kind == TK_RDBL_CNT) sink.add(slot, 1);
    } else {
      if (kind != TK_SKIP) return false;
      pos = tk_skip_slow(w.in8, w.last_blk, pos, end, wt);  // fixed32, groups, illegal wire types
      if (pos == TK_FAIL) return false;
    }
  }
}

}  // namespace dfgpu
