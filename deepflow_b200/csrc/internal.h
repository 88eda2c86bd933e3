/*
 * internal.h — host-side structures of libdfgpu.so (context, per-batch arena,
 * dictionary handles).  Not part of the ABI.
 */
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/dfgpu.h"
#include "kg.cuh"
#include "tables.cuh"

namespace dfgpu {

// ---- host image of a dictionary, then its device copy ------------------------------
struct HostDict {
  std::vector<T64Entry> t64;
  uint32_t t64_mask = 0;
  std::vector<T6Entry> t6;
  uint32_t t6_mask = 0;
  std::vector<DInfo> infos;
  std::vector<uint64_t> gp_vals;
  std::vector<DVtap> vtaps;
  std::vector<DCidr4> cidr4;
  std::vector<DRange> cidr4_ranges;
  std::vector<DCidr6> cidr6;
  std::vector<DRange> cidr6_ranges;
  uint32_t counts[TID_COUNT] = {0};
  uint32_t counts6[T6_COUNT] = {0};
  uint32_t region_id = 0, analyzer_id = 0;
  uint64_t version = 0;
};
int build_dict(const dfgpu_dict_snapshot& s, HostDict& out);

struct DeviceDict {
  void* blob = nullptr;  // one cudaMalloc holding every table
  size_t blob_bytes = 0;
  DevDict view{};        // device pointers into blob
  uint64_t version = 0;
  ~DeviceDict() {
    if (blob) cudaFree(blob);
  }
};

// geo: sorted disjoint ranges (libs/geo GEO_ENTRIES) + canonical name ids
struct DevGeo {
  const uint32_t* start = nullptr;  // n, ascending
  const uint32_t* end = nullptr;
  const uint8_t* code = nullptr;    // canonical string id per range
  uint32_t n = 0;
};

// ---- column schema -------------------------------------------------------------------
struct ColSpec {
  const char* name;
  int ch_type;
  int elem_size;  // fixed part per row (keys for LC_STR, values for nullable, 0 for pure var)
  int var;        // index of the variable-length stream(s) it owns, -1 none
};

constexpr int MAX_COLS = 160;
constexpr int MAX_VAR = 40;

// Where every column of one table lives inside the batch arena.
struct TableLayout {
  uint64_t fix_off[MAX_COLS];   // fixed-width part (values / keys / nullable values)
  uint64_t aux_off[MAX_COLS];   // nulls (nullable), arr_offsets (arrays), pos (strings)
  uint64_t var_pos_off[MAX_VAR];   // dfgpu_pos entries of a string stream
  uint64_t var_buf_off[MAX_VAR];   // bytes / element values of a stream
};

// Written by the device, read back by fetch.
struct BatchStats {
  uint32_t nrec;          // records found by framing
  uint32_t nrows;         // rows emitted into the main table
  uint32_t need_repair;   // a record failed to decode: validity must be re-derived
  uint32_t error_count;   // ErrorCount (decoder.go:74)
  uint32_t frame_errors;  // frames whose length chain broke (subset of error_count)
  uint32_t l7_counts[6];  // http,dns,sql,nosql,rpc,mq (decoder.go:498-525)
  uint32_t pad;
  uint64_t var_total[MAX_VAR];  // elements of stream v
  uint64_t var_bytes[MAX_VAR];  // bytes of stream v
  uint64_t ft_tot[8];           // flow_tag stage totals (see l7.cu FT_* slots)
  uint64_t ft_off[32];          // flow_tag table column offsets inside the ft arena
  uint64_t ft_end;              // bytes used in the ft arena
  uint32_t ft_overflow;         // ft arena too small: host grows it and re-runs the writer
  uint32_t ft_count;            // entries held by the flow_tag set (TagSetState::count): the host schedules rebuilds
  uint32_t ft_unrecorded;       // cumulative: emissions of keys the saturated set could not store
  uint32_t ft_evicted;          // cumulative: live entries dropped oldest-first by rebuilds
  uint32_t ft_any;              // phase 0 found a row with a new / expired key: phases 1-3 have work to do
  uint32_t ft_pad;
  uint64_t last_trace_id_index;
  uint32_t lc_first[8][256];    // first record index each LowCardinality code appears at
};

}  // namespace dfgpu

struct dfgpu_batch {
  dfgpu_ctx* ctx = nullptr;
  int msg_type = 0;
  uint8_t* host = nullptr;  // pinned copy of the arena (dfgpu_cfg.block_lz4: of the frame streams)
  uint64_t lz_frame_bytes = 0;
  dfgpu::BatchStats stats{};
  dfgpu_counters counters{};
  struct Table {
    int id = -1;
    uint64_t nrows = 0;
    std::vector<dfgpu_col> cols;
  };
  std::vector<Table> tables;
  struct TagFp {  // 128-bit lru-key fingerprints of a flow_tag table's rows, row order
    int table;
    const uint64_t* fps;
    uint64_t n;
  };
  std::vector<TagFp> tag_fps;
  std::vector<std::vector<uint8_t>> owned;   // small host-built buffers (LC dictionaries)
  std::vector<std::vector<dfgpu_pos>> owned_pos;
};
