/*
 * ORACLE — test infrastructure only.
 * columns.h: host containers shaped like the ch-go v0.69 proto.Col* types the
 * reference appends to (server/libs/ckdb/ckdb.go:33-67 and every
 * *_column_block.go).  ch-go itself is not in the reference tree; the layouts are
 * the ones its public API exposes:
 *   ColUInt8/16/32/64, ColInt32/64, ColFloat64, ColIPv4, ColDateTime  = []T
 *   ColIPv6 = [][16]byte, ColDateTime64 = []int64 (+precision)
 *   ColStr  = {Buf []byte; Pos []Position{Start,End}}
 *   ColArr[T] = {Offsets ColUInt64 (cumulative ends); Data Col[T]}
 *   ColNullable[T] = {Nulls ColUInt8; Values Col[T]} (NULL rows hold T's zero)
 *   ColLowCardinality[string]: Append(v) records v; the dictionary (first
 *     appearance order) + per-row keys are what Prepare() derives.  We export
 *     both forms (see DFGPU_CH_LC_STR / DFGPU_CH_LC_STR_ROWS).
 */
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "../include/dfgpu.h"

namespace oracle {

struct Column {
  std::string name;
  int ch_type = 0;
  int elem_size = 0;
  uint64_t nrows = 0;
  std::vector<uint8_t> data;        // fixed values | Buf | element values
  std::vector<dfgpu_pos> pos;       // strings
  std::vector<uint64_t> arr_off;    // arrays
  uint64_t nelems = 0;
  std::vector<uint8_t> nulls;       // nullable
  // LowCardinality dictionary (first-appearance order)
  std::vector<uint8_t> dict_buf;
  std::vector<dfgpu_pos> dict_pos;
  std::unordered_map<std::string, uint32_t> dict_idx;

  Column(const char* n, int t, int es = 0) : name(n), ch_type(t), elem_size(es) {}

  template <class T>
  void put(T v) {
    size_t o = data.size();
    data.resize(o + sizeof(T));
    memcpy(data.data() + o, &v, sizeof(T));
  }
  template <class T>
  void append(T v) {  // fixed width
    put<T>(v);
    nrows++;
  }
  void append_ip6(const uint8_t* ip, size_t len) {  // ckdb.AppendIPv6 (ckdb.go:46-56)
    uint8_t b[16] = {0};
    if (len) memcpy(b, ip, len > 16 ? 16 : len);
    data.insert(data.end(), b, b + 16);
    nrows++;
  }
  void append_str(std::string_view s) {  // ColStr.Append
    dfgpu_pos p{(int64_t)data.size(), (int64_t)(data.size() + s.size())};
    data.insert(data.end(), s.begin(), s.end());
    pos.push_back(p);
    nrows++;
  }
  void append_lc(std::string_view s) {  // ColLowCardinality[string].Append, dictionary coded
    std::string k(s);
    auto it = dict_idx.find(k);
    uint32_t id;
    if (it == dict_idx.end()) {
      id = (uint32_t)dict_pos.size();
      dict_idx.emplace(std::move(k), id);
      dfgpu_pos p{(int64_t)dict_buf.size(), (int64_t)(dict_buf.size() + s.size())};
      dict_buf.insert(dict_buf.end(), s.begin(), s.end());
      dict_pos.push_back(p);
    } else {
      id = it->second;
    }
    put<uint32_t>(id);
    nrows++;
  }
  template <class T>
  void append_nullable(const T* v) {  // ckdb.AppendColNullable (ckdb.go:38-44)
    nulls.push_back(v ? 0 : 1);
    put<T>(v ? *v : T());
    nrows++;
  }
  template <class T>
  void append_arr(const std::vector<T>& v) {  // ColArr[T].Append
    for (auto& x : v) put<T>(x);
    nelems += v.size();
    arr_off.push_back(nelems);
    nrows++;
  }
  void append_arr_str(const std::vector<std::string_view>& v) {
    for (auto& s : v) {
      dfgpu_pos p{(int64_t)data.size(), (int64_t)(data.size() + s.size())};
      data.insert(data.end(), s.begin(), s.end());
      pos.push_back(p);
    }
    nelems += v.size();
    arr_off.push_back(nelems);
    nrows++;
  }

  void clear() {  // Reset(): keep the capacity, like the reference's pooled column blocks
    nrows = 0;
    nelems = 0;
    data.clear();
    pos.clear();
    arr_off.clear();
    nulls.clear();
    dict_buf.clear();
    dict_pos.clear();
    dict_idx.clear();
  }

  dfgpu_col view() const {
    dfgpu_col c;
    memset(&c, 0, sizeof(c));
    c.name = name.c_str();
    c.ch_type = ch_type;
    c.elem_size = elem_size;
    c.nrows = nrows;
    c.data = data.data();
    c.data_bytes = data.size();
    c.pos = pos.empty() ? nullptr : pos.data();
    c.arr_offsets = arr_off.empty() ? nullptr : arr_off.data();
    c.nelems = nelems;
    c.nulls = nulls.empty() ? nullptr : nulls.data();
    c.dict_data = dict_buf.data();
    c.dict_pos = dict_pos.empty() ? nullptr : dict_pos.data();
    c.dict_n = dict_pos.size();
    return c;
  }
};

struct Block {
  std::vector<Column> cols;
  uint64_t nrows = 0;
  Column& add(const char* n, int t, int es = 0) {
    cols.emplace_back(n, t, es);
    return cols.back();
  }
  Column& operator[](size_t i) { return cols[i]; }
  void clear() {
    nrows = 0;
    for (auto& c : cols) c.clear();
  }
};

}  // namespace oracle
