/*
 * pipeline.cuh — kernels and parameter blocks shared by the L4 / L7 / metrics
 * pipelines: record framing (SimpleDecoder.ReadPB, server/libs/codec/
 * simple_codec.go:229-248), the per-warp slot scans that turn per-record sizes
 * into output offsets, and LowCardinality first-appearance ranking.
 *
 * Pipeline shape (one CUDA stream; the record count and everything sized by it stay on the device):
 *   frame_walk -> frame_scan -> prepare (plan: warps, column offsets) -> frame_fill
 *   decode      : 1 thread / record, the record decoded once into shared-memory words, the row built with the
 *                 warp converged (coalesced column stores), per-warp sizes of the variable-length outputs
 *                 -> warp_tot[slot][warp]
 *   slot_scan   : exclusive scan over warps, one CTA per slot
 *   finish      : stream offsets inside the arena, LC code ranks
 *   pack / SmartEncoding / flow_tag kernels of the message type (l7_strings.cu, kg_kernel.cuh, l7.cu, l4.cu,
 *                 metrics.cu), commit
 */
#pragma once
#include "internal.h"

namespace dfgpu {

constexpr uint32_t NO_BAD = 0xffffffffu;
constexpr int MAX_SLOTS = 40;
constexpr int LC_SPACES = 8;

struct StreamSpec {
  int8_t slot_elems;  // slot holding the element count, -1: one element per row
  int8_t slot_bytes;  // slot holding the byte count, -1: bytes = elems * elem_size
  uint8_t elem_size;
  uint8_t has_pos;    // stream carries a dfgpu_pos per element (Array(String))
};

struct PersistState {      // survives across batches (device resident)
  uint32_t id_counter;     // L4FlowCounter / L7FlowLogCounter (l4_flow_log.go:947)
  uint32_t pad;
  uint64_t last_trace_id_index;  // lastTraceIdIndex (l7_flow_log.go:294)
};

struct DynState {  // per batch, device resident, read back by fetch
  BatchStats stats;
  uint64_t var_pos_off[MAX_VAR];
  uint64_t var_buf_off[MAX_VAR];
  uint64_t var_end;       // first free byte after the last stream that goes back to the host
  uint32_t overflow;      // a device buffer is too small (arena, record index ...): fetch grows it and re-runs
  uint32_t rec_overflow;  // more records than the record index holds (dfgpu_cfg.max_batch_records)
  uint64_t dev_end;       // first free byte of the arena (>= var_end: device-only regions follow var_end)
  uint8_t lc_rank[LC_SPACES][256];
  uint64_t trace_carry_hash;  // scratch
  // The record count is only known on the device (framing): everything sized by it is planned there
  // too (k_prepare), so that dfgpu_run never waits for the GPU.
  uint32_t nwarps;        // ceil(nrec / 32)
  uint32_t plan_valid;    // plan_rows holds the row count of the main table (L7 repair pass), else rows = nrec
  uint32_t plan_rows;
  uint32_t pad2;
  uint64_t var_base;      // end of the fixed-width region = start of the variable-length streams
  uint64_t fix_off[MAX_COLS];  // fixed-width part of column c (values / keys / nullable values)
  uint64_t aux_off[MAX_COLS];  // nulls (nullable), arr_offsets (arrays), pos (strings)
};

struct BatchParams {
  const uint8_t* in;
  uint64_t in_bytes;
  const dfgpu_frame* frames;
  uint32_t nframes;
  uint32_t rec_cap;   // capacity of the record index (rec_off / rec_len / rec_frame)
  uint32_t warp_cap;  // stride of warp_tot / warp_base: [slot][warp_cap]
  int32_t repair;
  uint32_t* frame_nrec;
  uint32_t* frame_base;
  uint32_t* frame_chain_err;
  uint32_t* frame_bad;
  uint32_t* frame_drop;  // L7: rows dropped in earlier frames (0 unless a record failed to decode)
  uint32_t* rec_off;
  uint32_t* rec_len;
  uint32_t* rec_frame;
  uint32_t* rec_scratch;  // [nbytes/4 + 1] payload offsets, indexed by frame byte range
  uint32_t* warp_tot;   // [slot][nwarps]
  uint32_t* warp_base;  // [slot][nwarps]
  DynState* dyn;
  PersistState* persist;
  uint8_t* arena;
  uint64_t arena_bytes;
  uint32_t ncols;       // main table (L4 / L7): columns planned by k_prepare
  uint32_t nslots;
  uint32_t nstreams;
  uint32_t n_lc;
  uint32_t cfg_trace_disabled;
  uint32_t native_str;  // dfgpu_cfg.native_strings (l7_flow_log only): strings leave as uvarint(len) + bytes
  const void* ext;  // message-type specific device state (L7Dev)
  StreamSpec streams[MAX_VAR];
  DevDict dict;
  DevGeo geo;
  uint8_t col_es[MAX_COLS];   // bytes per row of the fixed-width part of column c (0: none)
  uint8_t col_aux[MAX_COLS];  // bytes per row of its aux part (nulls / arr_offsets / pos)
};

// records found by framing (device side); 0 when the record index overflowed
__device__ __forceinline__ uint32_t batch_nrec(const BatchParams& p) { return p.dyn->rec_overflow ? 0u : p.dyn->stats.nrec; }

extern const char* const kTapSideNames[18];  // TAPSideEnum.String() texts by canonical code (framing.cu)
// host launchers of the shared kernels (framing.cu)
void launch_frame_count(const BatchParams& p, cudaStream_t s);
void launch_prepare(const BatchParams& p, cudaStream_t s);  // record-count dependent planning (device side)
void launch_frame_fill(const BatchParams& p, cudaStream_t s);
// grid of a thread-per-record kernel: covers the record index capacity, CTAs past the batch's records exit
inline unsigned rec_grid(const BatchParams& p, unsigned block = 256) { return (p.rec_cap + block - 1) / block; }
void launch_slot_scan(const BatchParams& p, cudaStream_t s);
// generic: scan `nslots` slot rows of p.warp_tot into p.warp_base, totals to totals_out (device)
void launch_slot_scan_to(const BatchParams& p, uint32_t nslots, uint64_t* totals_out, cudaStream_t s);
void launch_finish(const BatchParams& p, cudaStream_t s);
void launch_commit(const BatchParams& p, int is_l7, cudaStream_t s);

// ---- warp helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v) {
  uint32_t lane = threadIdx.x & 31, x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += t;
  }
  return x - v;
}

// first-appearance tracking of a LowCardinality code (one atomic per distinct code per warp)
__device__ __forceinline__ void lc_note(uint32_t* first /*[256]*/, uint32_t code, uint32_t rec, bool active) {
  uint32_t m = __match_any_sync(0xffffffffu, active ? code : 0xffffffffu);
  if (active && (int)(threadIdx.x & 31) == __ffs(m) - 1) atomicMin(first + code, rec);
}

// sorted-range lookup replacing the netmask tree (libs/geo/netmask_tree.go:97-104)
__device__ __forceinline__ uint32_t geo_code(const DevGeo& g, uint32_t ip) {
  uint32_t lo = 0, hi = g.n;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(g.start + mid) <= ip) lo = mid + 1; else hi = mid;
  }
  if (lo == 0) return 0;
  return ip <= __ldg(g.end + lo - 1) ? __ldg(g.code + lo - 1) : 0;
}

// canonical string ids of flow_metrics.TAPSideEnum.String() (tag.go:211-233);
// 0 = "" (slice slot without a name)
__device__ __forceinline__ uint32_t tapside_code(uint32_t v) {
  switch (v & 0xff) {
    case 0: return 1;  case 1: return 2;  case 2: return 3;  case 4: return 4;
    case 9: return 5;  case 10: return 6; case 17: return 7; case 18: return 8;
    case 25: return 9; case 26: return 10; case 33: return 11; case 34: return 12;
    case 41: return 13; case 42: return 14; case 49: return 15; case 50: return 16;
    case 48: return 17;
  }
  return 0;
}

// getStatus l4_flow_log.go:745-760 (CloseType is a uint8, datatype/flow.go:33-84)
__device__ __forceinline__ uint8_t l4_status(uint8_t t, uint8_t proto) {
  bool tcp = proto == 6;
  if (t == 1 || t == 5 || t == 20 || (!tcp && t == 3) || t == 11 || t == 8) return 0;
  if (t == 10 || t == 9 || t == 13 || t == 18) return 4;
  if (tcp && (t == 2 || t == 3 || t == 7 || t == 15 || t == 17 || t == 19)) return 3;
  return 2;
}

// MarshalInt32WithSpecialID flow-metrics/tag.go:560-565
__device__ __forceinline__ int32_t marshal_epc(int32_t v) {
  if (v > 0 || v == -1 || v == -2) return v;
  return (int32_t)((uint32_t)v & 0xffffu);
}

}  // namespace dfgpu
