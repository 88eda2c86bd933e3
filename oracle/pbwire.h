/*
 * ORACLE — test infrastructure only (see oracle/README.md).  Never linked into
 * or called from the product path (libdfgpu.so).
 *
 * pbwire.h: table-driven protobuf wire decoder restating the acceptance rules of
 * the gogo/protobuf v1.3.2 generated `Unmarshal` methods the reference calls at
 * server/libs/codec/simple_codec.go:241 (the generated flow_log.pb.go /
 * metric.pb.go are NOT in the reference tree; gogo's generator emits the same
 * loop for every message, restated here from its published template):
 *   - tag = varint; shift >= 64 -> ErrIntOverflow; running past the end ->
 *     io.ErrUnexpectedEOF; fieldNum <= 0 -> "illegal tag"; wireType 4 at message
 *     level -> "wiretype end group for non-group".
 *   - known field with the wrong wire type -> "wrong wireType" error; repeated
 *     scalar fields accept both the packed (2) and the unpacked encoding.
 *   - scalars: last one wins; uint32/int32 keep the low 32 bits of the varint.
 *   - nested message on wire type 2: length < 0 -> ErrInvalidLength, past end ->
 *     EOF; decoded INTO the existing sub-message (merge, no reset).
 *   - strings/bytes are sliced without UTF-8 validation.
 *   - unknown fields skipped by skipXxx(): wire types 0,1,2,5 and groups (3/4 with
 *     a depth counter); 6,7 -> "illegal wireType".
 * Schema = message/flow_log.proto:14-310 and message/metric.proto:14-195.
 */
#pragma once
#include <cstdint>
#include <cstring>
#include <string_view>
#include <vector>

namespace oracle {

enum Kind : uint8_t {
  K_NONE = 0,
  K_U32,   // uint32 varint
  K_U64,   // uint64 varint
  K_I32,   // int32 varint (sign-extended 10-byte form on the wire)
  K_STR,   // string / bytes
  K_MSG,   // nested message
  K_RU32,  // repeated uint32 (packed or not)
  K_RU64,  // repeated uint64
  K_RSTR,  // repeated string
  K_RDBL,  // repeated double
};

struct MsgDesc;
struct FieldDesc {
  uint32_t num;
  Kind kind;
  const MsgDesc* sub;
};
struct MsgDesc {
  const char* name;
  std::vector<FieldDesc> fields;
  uint32_t max_num;
  std::vector<int> by_num;  // field number -> index in fields, -1 unknown
  MsgDesc(const char* n, std::vector<FieldDesc> f) : name(n), fields(std::move(f)) {
    max_num = 0;
    for (auto& x : fields) max_num = x.num > max_num ? x.num : max_num;
    by_num.assign(max_num + 1, -1);
    for (size_t i = 0; i < fields.size(); i++) by_num[fields[i].num] = (int)i;
  }
};

// A decoded message.  Indexed by FIELD NUMBER.  Sub-messages are allocated once
// and reused (like the reference's pooled/pre-allocated pb objects,
// server/libs/datatype/pb/flow_log.go:23-165); `has[num]` says whether the
// sub-message pointer would be non-nil in Go for THIS record.
struct Msg {
  const MsgDesc* d = nullptr;
  std::vector<uint64_t> u;
  std::vector<std::string_view> s;
  std::vector<Msg*> m;
  std::vector<uint8_t> has;
  std::vector<std::vector<uint64_t>> ru;
  std::vector<std::vector<std::string_view>> rs;
  std::vector<std::vector<double>> rd;

  explicit Msg(const MsgDesc* desc) : d(desc) {
    size_t n = desc->max_num + 1;
    u.assign(n, 0);
    s.assign(n, {});
    m.assign(n, nullptr);
    has.assign(n, 0);
    ru.resize(n);
    rs.resize(n);
    rd.resize(n);
    for (auto& f : desc->fields)
      if (f.kind == K_MSG) m[f.num] = new Msg(f.sub);
  }
  ~Msg() {
    for (auto* p : m) delete p;
  }
  Msg(const Msg&) = delete;
  Msg& operator=(const Msg&) = delete;

  // proto Reset() of this message and everything below it.
  void reset() {
    for (auto& f : d->fields) {
      switch (f.kind) {
        case K_U32: case K_U64: case K_I32: u[f.num] = 0; break;
        case K_STR: s[f.num] = {}; break;
        case K_MSG: has[f.num] = 0; m[f.num]->reset(); break;
        case K_RU32: case K_RU64: ru[f.num].clear(); break;
        case K_RSTR: rs[f.num].clear(); break;
        case K_RDBL: rd[f.num].clear(); break;
        default: break;
      }
    }
  }
  uint32_t u32(uint32_t n) const { return (uint32_t)u[n]; }
  uint64_t u64(uint32_t n) const { return u[n]; }
  int32_t i32(uint32_t n) const { return (int32_t)(uint32_t)u[n]; }
  std::string_view str(uint32_t n) const { return s[n]; }
  const Msg& sub(uint32_t n) const { return *m[n]; }
  Msg& sub(uint32_t n) { return *m[n]; }
  bool present(uint32_t n) const { return has[n] != 0; }
};

// Returns false on any error the gogo Unmarshal would return.
inline bool read_varint(const uint8_t* p, size_t l, size_t& i, uint64_t& out) {
  uint64_t v = 0;
  for (unsigned shift = 0;; shift += 7) {
    if (shift >= 64) return false;  // ErrIntOverflow
    if (i >= l) return false;       // io.ErrUnexpectedEOF
    uint8_t b = p[i++];
    v |= (uint64_t)(b & 0x7f) << shift;
    if (b < 0x80) break;
  }
  out = v;
  return true;
}

// skipFlowLog()/skipMetric(): returns bytes consumed or -1.
inline long skip_field(const uint8_t* p, size_t l) {
  size_t i = 0;
  int depth = 0;
  while (i < l) {
    uint64_t wire;
    if (!read_varint(p, l, i, wire)) return -1;
    int wt = (int)(wire & 7);
    switch (wt) {
      case 0: {
        // gogo walks bytes until one without the continuation bit
        for (unsigned shift = 0;; shift += 7) {
          if (shift >= 64) return -1;
          if (i >= l) return -1;
          i++;
          if (p[i - 1] < 0x80) break;
        }
        break;
      }
      case 1: i += 8; break;
      case 2: {
        uint64_t len;
        if (!read_varint(p, l, i, len)) return -1;
        if ((int64_t)len < 0) return -1;  // ErrInvalidLength
        // gogo computes iNdEx += length in `int`; a huge value overflows to
        // negative -> ErrInvalidLength.  Either way the caller fails.
        if (len > (uint64_t)(l - i) + (1ull << 40)) return -1;
        i += (size_t)len;
        break;
      }
      case 3: depth++; break;
      case 4:
        if (depth == 0) return -1;  // ErrUnexpectedEndOfGroup
        depth--;
        break;
      case 5: i += 4; break;
      default: return -1;  // illegal wireType
    }
    if (depth == 0) return (long)i;
  }
  return -1;  // io.ErrUnexpectedEOF
}

inline bool decode_msg(Msg& msg, const uint8_t* p, size_t l) {
  const MsgDesc* d = msg.d;
  size_t i = 0;
  while (i < l) {
    size_t pre = i;
    uint64_t wire;
    if (!read_varint(p, l, i, wire)) return false;
    int32_t field = (int32_t)(wire >> 3);
    int wt = (int)(wire & 7);
    if (wt == 4) return false;    // end group for non-group
    if (field <= 0) return false; // illegal tag
    int idx = ((uint32_t)field <= d->max_num) ? d->by_num[field] : -1;
    if (idx < 0) {
      long sk = skip_field(p + pre, l - pre);
      if (sk < 0) return false;
      if (pre + (size_t)sk > l) return false;
      i = pre + (size_t)sk;
      continue;
    }
    const FieldDesc& f = d->fields[idx];
    switch (f.kind) {
      case K_U32:
      case K_I32: {
        if (wt != 0) return false;
        uint64_t v;
        if (!read_varint(p, l, i, v)) return false;
        msg.u[field] = (uint32_t)v;  // uint32(b&0x7F)<<shift: bits >=32 fall off
        break;
      }
      case K_U64: {
        if (wt != 0) return false;
        uint64_t v;
        if (!read_varint(p, l, i, v)) return false;
        msg.u[field] = v;
        break;
      }
      case K_STR: {
        if (wt != 2) return false;
        uint64_t len;
        if (!read_varint(p, l, i, len)) return false;
        if ((int64_t)len < 0) return false;
        if (len > l - i) return false;
        msg.s[field] = std::string_view((const char*)p + i, (size_t)len);
        i += (size_t)len;
        break;
      }
      case K_MSG: {
        if (wt != 2) return false;
        uint64_t len;
        if (!read_varint(p, l, i, len)) return false;
        if ((int64_t)len < 0) return false;
        if (len > l - i) return false;
        msg.has[field] = 1;
        if (!decode_msg(*msg.m[field], p + i, (size_t)len)) return false;
        i += (size_t)len;
        break;
      }
      case K_RU32:
      case K_RU64: {
        if (wt == 0) {
          uint64_t v;
          if (!read_varint(p, l, i, v)) return false;
          msg.ru[field].push_back(f.kind == K_RU32 ? (uint32_t)v : v);
        } else if (wt == 2) {
          uint64_t len;
          if (!read_varint(p, l, i, len)) return false;
          if ((int64_t)len < 0) return false;
          if (len > l - i) return false;
          size_t end = i + (size_t)len;
          while (i < end) {
            uint64_t v;
            // gogo bounds the packed varints by the whole message (l), then
            // keeps looping while iNdEx < postIndex
            if (!read_varint(p, l, i, v)) return false;
            msg.ru[field].push_back(f.kind == K_RU32 ? (uint32_t)v : v);
          }
        } else {
          return false;
        }
        break;
      }
      case K_RSTR: {
        if (wt != 2) return false;
        uint64_t len;
        if (!read_varint(p, l, i, len)) return false;
        if ((int64_t)len < 0) return false;
        if (len > l - i) return false;
        msg.rs[field].push_back(std::string_view((const char*)p + i, (size_t)len));
        i += (size_t)len;
        break;
      }
      case K_RDBL: {
        if (wt == 1) {
          if (i + 8 > l) return false;
          double v;
          memcpy(&v, p + i, 8);
          i += 8;
          msg.rd[field].push_back(v);
        } else if (wt == 2) {
          uint64_t len;
          if (!read_varint(p, l, i, len)) return false;
          if ((int64_t)len < 0) return false;
          if (len > l - i) return false;
          size_t end = i + (size_t)len;
          while (i < end) {
            if (i + 8 > l) return false;
            double v;
            memcpy(&v, p + i, 8);
            i += 8;
            msg.rd[field].push_back(v);
          }
        } else {
          return false;
        }
        break;
      }
      default:
        return false;
    }
  }
  return i <= l;
}

}  // namespace oracle
