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
#define DF_FFS(x) __ffs((int)(x))
#define DF_POPC(x) ((uint32_t)__popc(x))
#else
#define DF_FFS(x) __builtin_ffs((int)(x))
#define DF_POPC(x) ((uint32_t)__builtin_popcount(x))
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
  TK_RSTR_CNT = 5,   // repeated string: word[slot] += 1, word[slot + 1] += bytes (+ length prefixes);
                     // element i also goes to sink.elem(ref, i, off, len)
  TK_RSTR_FIRST2 = 6,  // repeated string: ref[ref], ref[ref + 1] = first two, word[slot] = count
  TK_RDBL_CNT = 7,   // repeated double (packed or not): word[slot] += 1 per value; word[slot + 1] += 1 per packed
                     // occurrence (its range -> sink.elem(ref, 0, off, len)), += 2 per unpacked value
  TK_RVAR_CNT = 8,   // repeated varint (packed or not): word[slot] += 1 per value
  TK_RVAR_DIGITS = 9,  // repeated uint64: word[slot] += 1, word[slot + 1] += decimal chars + ',' separators
                       // (both RVAR kinds: value i also goes to sink.velem(ref, i, lo, hi))
  TK_STR2W = 10,     // string / bytes -> word[slot] = offset, word[slot + 1] = length
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
// `in` is the staging buffer (device, 8-byte aligned), positions are absolute byte offsets into it.  The
// window holds THREE aligned 8-byte blocks starting at the cursor's block, so that one seek per field makes
// the 12 bytes behind the cursor available (a tag of <= 2 bytes + a varint of <= 10): tag and value are
// decoded from registers, with no load of their own.  Sequential walks reload one block per 8 bytes
// consumed (the third block doubles as the prefetch); last_blk bounds the loads inside the buffer.
struct TokWin {
  const TokU2* in8;
  TokU2 b0, b1, b2, pf;  // pf: block blk + 3, prefetched
  uint32_t blk;       // index of b0
  uint32_t last_blk;  // last block holding a valid byte of the buffer
  DF_HD void init(const uint8_t* in, uint64_t nbytes) {
    in8 = reinterpret_cast<const TokU2*>(in);
    blk = 0xfffffff0u;
    last_blk = nbytes ? (uint32_t)((nbytes - 1) >> 3) : 0u;
    b0 = b1 = b2 = pf = TokU2{0, 0};
  }
  static DF_HD TokU2 ld8(const TokU2* p) {
#if defined(__CUDA_ARCH__)
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    return TokU2{v.x, v.y};
#else
    return *p;
#endif
  }
  DF_HD TokU2 ldc(uint32_t b) const { return ld8(in8 + (b < last_blk ? b : last_blk)); }
  DF_HD void seek(uint32_t pos) {
    const uint32_t nb = pos >> 3;
    if (nb != blk) {
      if (nb == blk + 1) {  // the walk moved into the next block: the prefetched one becomes the third
        b0 = b1;
        b1 = b2;
        b2 = pf;
      } else {
        b0 = ldc(nb);
        b1 = ldc(nb + 1);
        b2 = ldc(nb + 2);
      }
      // not touched before the next block change: its latency hides behind a field or two.  The asm ties the
      // load's destination to pf's own registers — left to itself ptxas loads into a temporary and copies it
      // at once, which turns the prefetch into a blocking load (ncu, round 2: 23 % of the kernel's stall samples)
#if defined(__CUDA_ARCH__)
      {
        const TokU2* q = in8 + (nb + 3 < last_blk ? nb + 3 : last_blk);
        asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "+r"(pf.x), "+r"(pf.y) : "l"(q));
      }
#else
      pf = ldc(nb + 3);
#endif
      blk = nb;
    }
  }
  // bytes pos .. pos + 11 as three little-endian words (after seek(pos)); bytes past the buffer are garbage
  // the caller never uses (it bounds every field by the enclosing message)
  DF_HD void get12(uint32_t pos, uint32_t& x0, uint32_t& x1, uint32_t& x2) const {
    const uint32_t k = pos & 7u, s = (k & 3u) * 8u;
    const bool up = k >= 4;
    const uint32_t w0 = up ? b0.y : b0.x, w1 = up ? b1.x : b0.y, w2 = up ? b1.y : b1.x, w3 = up ? b2.x : b1.y;
    x0 = DF_FUNNEL_R(w0, w1, s);
    x1 = DF_FUNNEL_R(w1, w2, s);
    x2 = DF_FUNNEL_R(w2, w3, s);
  }
  DF_HD uint32_t byte(uint32_t pos) {
    seek(pos);
    return DF_BYTE_PERM(b0.x, b0.y, pos & 7u) & 0xffu;
  }
};

// Varint of 1..10 bytes held in registers (v0 = bytes 0..3, v1 = 4..7, v2 = 8..9), branch-free on its value:
// the terminator is found with one bit scan, the 7-bit groups are compacted with masks and shifts.  (ncu,
// round 2: the byte-at-a-time loop was 40 % of the decode kernel's instructions at 9 of 32 lanes — in a warp
// of 32 records some lane is always inside a 9-byte timestamp or a 10-byte negative int32.)
// Returns the length in bytes, 0 when the varint has no terminator within 10 bytes (ErrIntOverflow).
DF_HD uint32_t tk_varint_regs(const uint32_t v0, const uint32_t v1, const uint32_t v2, uint32_t& vlo, uint32_t& vhi) {
  const uint32_t tl = ~v0 & 0x80808080u, th = ~v1 & 0x80808080u;  // bit 7 of the bytes without a continuation bit
  // t ^ (t - 1): every bit up to and including the lowest terminator bit (all ones when there is none)
  const uint32_t ml = tl ^ (tl - 1u), mh = tl ? 0u : th ^ (th - 1u);
  const bool wide = (tl | th) == 0u;  // 9 or 10 bytes
  if (wide && (v2 & 0x8080u) == 0x8080u) return 0u;
  const uint32_t len = ((DF_POPC(ml) + DF_POPC(mh)) >> 3) + (wide ? 1u + ((v2 >> 7) & 1u) : 0u);
  const uint32_t xl = v0 & ml & 0x7f7f7f7fu, xh = v1 & mh & 0x7f7f7f7fu;
  const uint32_t tl2 = (xl & 0x007f007fu) | ((xl >> 1) & 0x3f803f80u), th2 = (xh & 0x007f007fu) | ((xh >> 1) & 0x3f803f80u);
  const uint32_t l28 = (tl2 & 0x3fffu) | ((tl2 >> 2) & 0x0fffc000u), h28 = (th2 & 0x3fffu) | ((th2 >> 2) & 0x0fffc000u);
  vlo = l28 | (h28 << 28);
  vhi = (h28 >> 4) | (wide ? (v2 & 0x7fu) << 24 | ((v2 & 0x80u) ? (v2 & 0x100u) << 23 : 0u) : 0u);
  return len;
}

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
  w.b0 = w.b1 = w.b2 = w.pf = TokU2{0, 0};
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
//   void elem(base, i, off, len)                                            i-th element of a repeated string field
//   void velem(base, i, lo, hi)                                             i-th value of a repeated varint field
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
    // ---- tag (1-2 bytes) and the varint behind it (value of wire type 0, length of wire type 2), both out
    // of the 12 bytes the window holds behind the cursor ----
    uint32_t wt, f;
    uint32_t vlo = 0, vhi = 0;
    bool have_v = false;
    {
      w.seek(pos);
      uint32_t x0, x1, x2;
      w.get12(pos, x0, x1, x2);
      const uint32_t two = (x0 >> 7) & 1u;
      if ((two && (x0 & 0x8000u)) || (two && end - pos < 2)) {
        // tag of 3+ bytes (field number >= 2048) or cut by the message end: generic path
        uint64_t tv;
        if (!tk_varint(w, pos, end, tv)) return false;
        wt = (uint32_t)(tv & 7);
        const int32_t fi = (int32_t)(tv >> 3);
        if (wt == 4 || fi <= 0) return false;
        f = (uint32_t)fi;
      } else {
        const uint32_t t = (x0 & 0x7fu) | (two ? (x0 >> 1) & 0x3f80u : 0u);
        wt = t & 7u;
        f = t >> 3;
        if (wt == 4 || f == 0) return false;
        const uint32_t nb = 1u + two, sh = nb * 8u;
        pos += nb;
        if (wt == 0 || wt == 2) {
          if (pos >= end) return false;
          const uint32_t n = tk_varint_regs(DF_FUNNEL_R(x0, x1, sh), DF_FUNNEL_R(x1, x2, sh), x2 >> sh, vlo, vhi);
          if (n == 0 || n > end - pos) return false;  // longer than 10 bytes / truncated by the enclosing message
          pos += n;
          have_v = true;
        }
      }
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
      else if (kind == TK_RVAR_CNT || kind == TK_RVAR_DIGITS) {
        const uint32_t c = sink.word(slot);
        sink.u32(slot, c + 1);
        if (kind == TK_RVAR_DIGITS) sink.add(slot + 1, tk_digits((uint64_t)vhi << 32 | vlo) + (c ? 1u : 0u));
        sink.velem((ent >> 20) & 63u, c, vlo, vhi);
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
      else if (kind == TK_STR2W) {
        sink.u32(slot, off);
        sink.u32(slot + 1, n);
      }
      else if (kind == TK_RSTR_CNT) {
        const uint32_t c = sink.word(slot);
        sink.u32(slot, c + 1);
        sink.add(slot + 1, n + (tk_vsize(n) & len_prefix_mask));
        sink.elem((ent >> 20) & 63u, c, off, n);
      } else if (kind == TK_RSTR_FIRST2) {
        const uint32_t c = sink.word(slot);
        if (c < 2) sink.ref(((ent >> 20) & 63u) + c, off, n);
        sink.u32(slot, c + 1);
      } else if (kind == TK_RVAR_CNT || kind == TK_RVAR_DIGITS) {
        const uint32_t pend = off + n;
        while (pos < pend) {  // packed: varints bounded by the enclosing message, loop while before postIndex
          uint64_t v;
          if (!tk_varint(w, pos, end, v)) return false;
          const uint32_t c = sink.word(slot);
          sink.u32(slot, c + 1);
          if (kind == TK_RVAR_DIGITS) sink.add(slot + 1, tk_digits(v) + (c ? 1u : 0u));
          sink.velem((ent >> 20) & 63u, c, (uint32_t)v, (uint32_t)(v >> 32));
        }
        continue;
      } else if (kind == TK_RDBL_CNT) {
        sink.add(slot + 1, 1);
        sink.elem((ent >> 20) & 63u, 0, off, n);
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
      if (kind == TK_RDBL_CNT) {
        sink.add(slot, 1);
        sink.add(slot + 1, 2);
      }
    } else {
      if (kind != TK_SKIP) return false;
      pos = tk_skip_slow(w.in8, w.last_blk, pos, end, wt);  // fixed32, groups, illegal wire types
      if (pos == TK_FAIL) return false;
    }
  }
}

}  // namespace dfgpu
