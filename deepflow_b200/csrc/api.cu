/*
 * api.cu — the C ABI of include/dfgpu.h: context, dictionary / geo upload, the
 * stage -> run -> fetch pipeline and the column-block views handed back to the
 * host shim.  See DESIGN.md for the data layout.
 */
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "ctx.h"
#include "lz4_dev.h"
#include "schema_l4.h"
#include "schema_l7.h"

using namespace dfgpu;

namespace dfgpu {
int l4_prepare(dfgpu_ctx* ctx, cudaStream_t s);
void l4_free_state(dfgpu_ctx* ctx);
void launch_l4_decode(dfgpu_ctx* ctx, cudaStream_t s);
void launch_l4_pack(dfgpu_ctx* ctx, cudaStream_t s);
void launch_l4_kg(dfgpu_ctx* ctx, cudaStream_t s);
void launch_l7_repair_plan(const BatchParams& p, cudaStream_t s);
void launch_l7_decode(const BatchParams& p, unsigned grid, cudaStream_t s);
void launch_l7_strings(const BatchParams& p, unsigned grid, cudaStream_t s);
void launch_l7_kg(const BatchParams& p, unsigned grid, cudaStream_t s);
void l7_grids(dfgpu_ctx* ctx, unsigned* decode, unsigned* strings, unsigned* kg);
int l7_flow_tags(dfgpu_ctx* ctx, cudaStream_t s);
int l7_prepare(dfgpu_ctx* ctx, cudaStream_t s);
int l7_tags_import(dfgpu_ctx* ctx, const uint64_t* fps, const uint32_t* times, uint64_t n);
int l7_tags_reset(dfgpu_ctx* ctx);
// metrics.cu
int metrics_passes(dfgpu_ctx* ctx, int repair);
int metrics_fetch(dfgpu_ctx* ctx, dfgpu_batch* b, uint8_t* (*acquire)(dfgpu_ctx*, size_t));
int rollup_add_resident(dfgpu_ctx* ctx, int table, uint32_t bucket);
int rollup_add_rows(dfgpu_ctx* ctx, int table, uint32_t bucket, const dfgpu_rollup_rows* rows);
int rollup_flush(dfgpu_ctx* ctx, int table, dfgpu_batch* b, uint8_t* (*acquire)(dfgpu_ctx*, size_t));
void rollup_free_state(dfgpu_ctx* ctx);
int rollup_describe(dfgpu_ctx* ctx, int table, int* ntag, int* nfields, uint8_t* elem, uint8_t* kind);
void metrics_free_state(dfgpu_ctx* ctx);
}  // namespace dfgpu

#define X(id, name, t, es) {name, t, es, -1},
static const ColSpec kL4Cols[] = {L4_COLUMNS(X)};
static const ColSpec kL7Cols[] = {L7_COLUMNS(X)};
#undef X

struct TableSchema {
  int table_id;
  const ColSpec* cols;
  int ncols;
};

// Arena bytes that hold any batch of `recs` records / `bytes` wire bytes: every row of the widest table
// (fixed-width columns + 16-byte Positions: l7 ~750 B, l4 ~900 B, flow_metrics ~1100 B per row) + the
// variable-length streams (<= the wire bytes, twice with their native length prefixes / device-only
// Positions) + slack.  Larger batches are caught on the device (DynState::overflow) and grown by fetch.
static uint64_t arena_need(uint64_t recs, uint64_t bytes) { return recs * 1200 + 3 * bytes + (8u << 20); }

// (Re)allocates every device buffer whose size follows the three capacities.  Called by dfgpu_init with
// the capacities of dfgpu_cfg, and by dfgpu_fetch when the device reported an overflow — never between
// dfgpu_stage and the final D2H of a batch that fits.
static int alloc_buffers(dfgpu_ctx* ctx, size_t in_cap, uint32_t rec_cap, size_t arena_cap) {
  if (in_cap > ctx->d_in_cap) {
    if (ctx->d_in) CU(cudaFree(ctx->d_in));
    if (ctx->d_rec_scratch) CU(cudaFree(ctx->d_rec_scratch));
    ctx->d_in = nullptr;
    ctx->d_rec_scratch = nullptr;
    CU(cudaMalloc(&ctx->d_in, in_cap + 64));
    CU(cudaMalloc(&ctx->d_rec_scratch, (in_cap / 4 + 2) * sizeof(uint32_t)));
    ctx->d_in_cap = in_cap;
    ctx->rec_scratch_cap = in_cap / 4 + 2;
  }
  if (rec_cap > ctx->rec_cap) {
    if (ctx->d_rec_u32) CU(cudaFree(ctx->d_rec_u32));
    if (ctx->d_warp) CU(cudaFree(ctx->d_warp));
    ctx->d_rec_u32 = nullptr;
    ctx->d_warp = nullptr;
    CU(cudaMalloc(&ctx->d_rec_u32, (size_t)rec_cap * 3 * sizeof(uint32_t)));
    ctx->rec_cap = rec_cap;
    ctx->warp_cap = (size_t)rec_cap / 32 + 2;
    CU(cudaMalloc(&ctx->d_warp, ctx->warp_cap * 2 * MAX_SLOTS * sizeof(uint32_t)));
  }
  if (arena_cap > ctx->arena_cap) {
    if (ctx->d_arena) CU(cudaFree(ctx->d_arena));
    ctx->d_arena = nullptr;
    CU(cudaMalloc(&ctx->d_arena, arena_cap));
    ctx->arena_cap = arena_cap;
  }
  return 0;
}
static int alloc_frames(dfgpu_ctx* ctx, uint32_t nframes) {
  if (nframes <= ctx->frames_cap) return 0;
  if (ctx->d_frames) CU(cudaFree(ctx->d_frames));
  if (ctx->d_frame_u32) CU(cudaFree(ctx->d_frame_u32));
  if (ctx->h_frame_u32) CU(cudaFreeHost(ctx->h_frame_u32));
  ctx->d_frames = nullptr;
  ctx->d_frame_u32 = nullptr;
  ctx->h_frame_u32 = nullptr;
  const size_t cap = std::max<size_t>(nframes + nframes / 2, 4096);
  CU(cudaMalloc(&ctx->d_frames, cap * sizeof(dfgpu_frame)));
  CU(cudaMalloc(&ctx->d_frame_u32, cap * 5 * sizeof(uint32_t)));  // nrec | base | chain_err | bad | drop
  CU(cudaHostAlloc(&ctx->h_frame_u32, cap * 2 * sizeof(uint32_t), cudaHostAllocDefault));
  ctx->frames_cap = (uint32_t)cap;
  ctx->h_frames_cap = (uint32_t)cap;
  return 0;
}

extern "C" {

int dfgpu_close(dfgpu_ctx* ctx);

int dfgpu_init(int device, const dfgpu_cfg* cfg, dfgpu_ctx** out) {
  if (!out) return DFGPU_EINVAL;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device >= n) return DFGPU_ENODEV;
  dfgpu_ctx* ctx = new dfgpu_ctx();
  ctx->device = device;
  if (cfg) ctx->cfg = *cfg;
  if (!ctx->cfg.max_batch_bytes) ctx->cfg.max_batch_bytes = 256ull << 20;
  if (!ctx->cfg.max_batch_records) ctx->cfg.max_batch_records = (uint32_t)(ctx->cfg.max_batch_bytes / 128);
  if (!ctx->cfg.flow_tag_cache_flush_timeout) ctx->cfg.flow_tag_cache_flush_timeout = 1800;
  if (!ctx->cfg.flow_tag_cache_max_size) ctx->cfg.flow_tag_cache_max_size = 1u << 18;
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
  if (const char* v = getenv("DFGPU_KG_OVERLAP")) ctx->kg_overlap = atoi(v);
  if (const char* v = getenv("DFGPU_KG_CTAS")) ctx->kg_ctas = atoi(v);
  if (const char* v = getenv("DFGPU_STR_CTAS")) ctx->str_ctas = atoi(v);
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_dyn, sizeof(DynState));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_persist, sizeof(PersistState));
  if (e == cudaSuccess) e = cudaMemset(ctx->d_persist, 0, sizeof(PersistState));
  if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_dyn, sizeof(DynState), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_persist, sizeof(PersistState), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
  if (e != cudaSuccess) {
    delete ctx;
    return DFGPU_ENODEV;
  }
  memset(ctx->h_persist, 0, sizeof(PersistState));
  // every device buffer is sized here, from the configured batch capacities: dfgpu_run neither
  // allocates nor waits for the GPU
  int rc = alloc_frames(ctx, 4096);
  if (!rc) rc = alloc_buffers(ctx, (size_t)ctx->cfg.max_batch_bytes, ctx->cfg.max_batch_records,
                              (size_t)arena_need(ctx->cfg.max_batch_records, ctx->cfg.max_batch_bytes));
  if (rc) {
    dfgpu_close(ctx);
    return rc;
  }
  *out = ctx;
  return DFGPU_OK;
}

void dfgpu_l7_free_state(dfgpu_ctx* ctx);

int dfgpu_close(dfgpu_ctx* ctx) {
  if (!ctx) return DFGPU_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  dfgpu_l7_free_state(ctx);
  l4_free_state(ctx);
  lz4_free_state(ctx);
  rollup_free_state(ctx);
  metrics_free_state(ctx);
  cudaFree(ctx->d_in);
  cudaFree(ctx->d_frames);
  cudaFree(ctx->d_frame_u32);
  cudaFree(ctx->d_rec_u32);
  cudaFree(ctx->d_rec_scratch);
  cudaFree(ctx->d_warp);
  cudaFree(ctx->d_dyn);
  cudaFree(ctx->d_persist);
  cudaFree(ctx->d_arena);
  cudaFreeHost(ctx->h_dyn);
  cudaFreeHost(ctx->h_persist);
  cudaFreeHost(ctx->h_frame_u32);
  for (auto& b : ctx->host_pool) cudaFreeHost(b.p);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  cudaStreamDestroy(ctx->stream);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  delete ctx;
  return DFGPU_OK;
}

const char* dfgpu_last_error(const dfgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
void* dfgpu_stream(dfgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t dfgpu_launch_count(const dfgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dfgpu_dict_load(dfgpu_ctx* ctx, uint16_t org_id, const dfgpu_dict_snapshot* snap) {
  if (!ctx || !snap) return DFGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  HostDict h;
  int rc = build_dict(*snap, h);
  if (rc) return fail(ctx, rc, "build_dict failed");
  auto dd = std::make_shared<DeviceDict>();
  struct Part {
    const void* src;
    size_t bytes;
    size_t off;
  };
  std::vector<Part> parts = {
      {h.t64.data(), h.t64.size() * sizeof(T64Entry), 0},
      {h.t6.data(), h.t6.size() * sizeof(T6Entry), 0},
      {h.infos.data(), h.infos.size() * sizeof(DInfo), 0},
      {h.gp_vals.data(), h.gp_vals.size() * sizeof(uint64_t), 0},
      {h.vtaps.data(), h.vtaps.size() * sizeof(DVtap), 0},
      {h.cidr4.data(), h.cidr4.size() * sizeof(DCidr4), 0},
      {h.cidr4_ranges.data(), h.cidr4_ranges.size() * sizeof(DRange), 0},
      {h.cidr6.data(), h.cidr6.size() * sizeof(DCidr6), 0},
      {h.cidr6_ranges.data(), h.cidr6_ranges.size() * sizeof(DRange), 0},
  };
  size_t total = 0;
  for (auto& p : parts) {
    p.off = total;
    total += align_up(p.bytes + 16, 256);
  }
  CU(cudaMalloc(&dd->blob, total));
  dd->blob_bytes = total;
  // plain synchronous copies: the upload happens off the hot path (60 s cadence)
  for (auto& p : parts)
    if (p.bytes) CU(cudaMemcpy((uint8_t*)dd->blob + p.off, p.src, p.bytes, cudaMemcpyHostToDevice));
  uint8_t* b = (uint8_t*)dd->blob;
  DevDict& v = dd->view;
  memset(&v, 0, sizeof v);
  v.t64 = (const T64Entry*)(b + parts[0].off);
  v.t64_mask = h.t64_mask;
  v.t6 = (const T6Entry*)(b + parts[1].off);
  v.t6_mask = h.t6_mask;
  v.infos = (const DInfo*)(b + parts[2].off);
  v.gp_vals = (const uint64_t*)(b + parts[3].off);
  v.vtaps = (const DVtap*)(b + parts[4].off);
  v.cidr4 = (const DCidr4*)(b + parts[5].off);
  v.cidr4_ranges = (const DRange*)(b + parts[6].off);
  v.cidr6 = (const DCidr6*)(b + parts[7].off);
  v.cidr6_ranges = (const DRange*)(b + parts[8].off);
  memcpy(v.counts, h.counts, sizeof v.counts);
  memcpy(v.counts6, h.counts6, sizeof v.counts6);
  v.region_id = h.region_id;
  v.analyzer_id = h.analyzer_id;
  dd->version = h.version;
  std::lock_guard<std::mutex> g(ctx->dict_mu);
  ctx->dicts[org_id] = dd;  // the previous snapshot is freed when the last batch using it ends
  return DFGPU_OK;
}

int dfgpu_geo_load(dfgpu_ctx* ctx, const dfgpu_geo_entry* entries, uint64_t n, const char* const* region_names,
                   uint32_t n_names) {
  if (!ctx || (n && !entries) || !region_names || n_names == 0) return DFGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  // canonical code per name: DecodeRegion (libs/geo/utils.go:27-40) maps out-of-range
  // codes to names[0]'s text; identical names share one LowCardinality entry.
  std::vector<std::string> names(region_names, region_names + n_names);
  std::vector<uint8_t> canon(256, 0);
  for (uint32_t c = 0; c < 256; c++) {
    uint32_t idx = c < n_names ? c : 0;
    const std::string& s = c < n_names ? names[c] : std::string("\xe6\x9c\xaa\xe7\x9f\xa5");
    uint32_t first = idx;
    for (uint32_t k = 0; k < n_names && k < 256; k++)
      if (names[k] == s) {
        first = k;
        break;
      }
    canon[c] = (uint8_t)first;
  }
  auto g = std::make_shared<GeoTables>();
  g->names.assign(256, std::string());
  for (uint32_t c = 0; c < n_names && c < 256; c++) g->names[c] = names[c];
  std::vector<dfgpu_geo_entry> e(entries, entries + n);
  std::stable_sort(e.begin(), e.end(), [](const dfgpu_geo_entry& a, const dfgpu_geo_entry& b) { return a.ip_start < b.ip_start; });
  // one buffer: start[n] | end[n] | code[n]
  std::vector<uint32_t> img(2 * n + (n + 3) / 4 + 1);
  uint8_t* code = reinterpret_cast<uint8_t*>(img.data() + 2 * n);
  for (uint64_t i = 0; i < n; i++) {
    img[i] = e[i].ip_start;
    img[n + i] = e[i].ip_end;
    code[i] = canon[e[i].region];
  }
  g->n = (uint32_t)n;
  if (n) {
    CU(cudaMalloc(&g->d_start, img.size() * 4));
    CU(cudaMemcpy(g->d_start, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
    g->d_end = g->d_start + n;
    g->d_code = reinterpret_cast<uint8_t*>(g->d_start + 2 * n);
  }
  // built aside, swapped in under the lock: a batch staged before this point keeps the old snapshot
  std::lock_guard<std::mutex> lk(ctx->dict_mu);
  ctx->geo = g;
  return DFGPU_OK;
}

}  // extern "C"

// ---- layout ---------------------------------------------------------------------------
static const TableSchema* main_schema(int msg) {
  static const TableSchema l4{DFGPU_TABLE_L4_FLOW_LOG, kL4Cols, (int)(sizeof kL4Cols / sizeof kL4Cols[0])};
  static const TableSchema l7{DFGPU_TABLE_L7_FLOW_LOG, kL7Cols, (int)(sizeof kL7Cols / sizeof kL7Cols[0])};
  if (msg == DFGPU_MSG_TAGGEDFLOW) return &l4;
  if (msg == DFGPU_MSG_PROTOCOLLOG) return &l7;
  return nullptr;
}

static size_t aux_bytes_per_row(int ch) {
  switch (ch) {
    case DFGPU_CH_STR:
    case DFGPU_CH_LC_STR_ROWS: return 16;  // dfgpu_pos
    case DFGPU_CH_ARR_UINT16:
    case DFGPU_CH_ARR_STR:
    case DFGPU_CH_ARR_LC_STR:
    case DFGPU_CH_ARR_FLOAT64: return 8;   // cumulative end offset
    case DFGPU_CH_NULLABLE_UINT8:
    case DFGPU_CH_NULLABLE_INT32:
    case DFGPU_CH_NULLABLE_INT64:
    case DFGPU_CH_NULLABLE_UINT64: return 1;  // null flag
  }
  return 0;
}
static bool is_array(int ch) {
  return ch == DFGPU_CH_ARR_UINT16 || ch == DFGPU_CH_ARR_STR || ch == DFGPU_CH_ARR_LC_STR || ch == DFGPU_CH_ARR_FLOAT64;
}

namespace dfgpu {
void l7_streams(BatchParams& bp);
}

static void setup_streams(BatchParams& bp, int msg) {
  memset(bp.streams, 0, sizeof bp.streams);
  if (msg == DFGPU_MSG_TAGGEDFLOW) {
    bp.nslots = L4S_COUNT;
    bp.nstreams = 3;
    bp.n_lc = 3;
    bp.streams[0] = StreamSpec{-1, L4S_AGG_BYTES, 1, 0};   // aggregated_flow_ids bytes
    bp.streams[1] = StreamSpec{L4S_ACL_ELEMS, -1, 2, 0};   // acl_gids values
    bp.streams[2] = StreamSpec{-1, L4S_DOM_BYTES, 1, 0};   // request_domain bytes
  } else {
    l7_streams(bp);
  }
}

// per-column bytes per row of the main table, for the device-side layout (k_prepare)
static void setup_columns(BatchParams& bp, const TableSchema& s) {
  bp.ncols = (uint32_t)s.ncols;
  for (int c = 0; c < s.ncols; c++) {
    const ColSpec& cs = s.cols[c];
    bp.col_es[c] = (uint8_t)(is_array(cs.ch_type) ? 0 : cs.elem_size);
    const bool str_pos = cs.ch_type == DFGPU_CH_STR || cs.ch_type == DFGPU_CH_LC_STR_ROWS;
    // native strings: String columns carry no Position array in the host-bound part of the arena
    // (k_finish puts the device-only ones behind var_end)
    bp.col_aux[c] = (uint8_t)((bp.native_str && str_pos) ? 0 : aux_bytes_per_row(cs.ch_type));
  }
}

// Everything after framing: planning on the device, the passes of the message type, the commit.
// repair = 1 when re-running with the frames' cut points known (a record failed to decode).
static int launch_passes(dfgpu_ctx* ctx, int repair) {
  BatchParams& bp = ctx->bp;
  bp.repair = repair;
  cudaStream_t s = ctx->stream;
  if (ctx->cur_msg == DFGPU_MSG_METRICS) {
    launch_prepare(bp, s);
    ctx->launches++;
    if (bp.nframes) {
      launch_frame_fill(bp, s);
      ctx->launches++;
    }
    return metrics_passes(ctx, repair);
  }
  // reset first-appearance tables
  CU(cudaMemsetAsync(&ctx->d_dyn->stats.lc_first, 0xff, sizeof(ctx->d_dyn->stats.lc_first), s));
  const bool l7 = ctx->cur_msg == DFGPU_MSG_PROTOCOLLOG;
  {
    if (repair) {
      launch_l7_repair_plan(bp, s);  // rows dropped before every frame, rows the layout is planned for
      ctx->launches++;
    } else if (bp.nframes) {
      CU(cudaMemsetAsync(bp.frame_drop, 0, (size_t)bp.nframes * 4, s));
    }
  }
  launch_prepare(bp, s);
  ctx->launches++;
  if (bp.nframes) {
    launch_frame_fill(bp, s);
    ctx->launches++;
  }
  ctx->mark(2);
  unsigned g_dec = 0, g_str = 0, g_kg = 0;
  const bool kg_side = l7 && ctx->kg_overlap;
  if (l7) {
    l7_grids(ctx, &g_dec, &g_str, &g_kg);
    if (ctx->kg_ctas) g_kg = g_kg / 4 * (unsigned)ctx->kg_ctas;
    if (ctx->str_ctas) g_str = g_str / 6 * (unsigned)ctx->str_ctas;
    launch_l7_decode(bp, g_dec, s);  // decode + fixed-width columns + per-warp sizes (+ trace carry scan)
    if (kg_side) {  // SmartEncoding needs nothing but the decode kernel's hand-over: fork
      CU(cudaEventRecord(ctx->ev_fork, s));
      CU(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
      launch_l7_kg(bp, g_kg, ctx->stream2);
      ctx->launches++;
      CU(cudaEventRecord(ctx->ev_join, ctx->stream2));
    }
  } else {
    launch_l4_decode(ctx, s);
  }
  ctx->mark(3);
  launch_slot_scan(bp, s);
  launch_finish(bp, s);
  ctx->launches += 4;
  ctx->mark(4);
  cudaEvent_t pe0 = nullptr, pe1 = nullptr;
  if (ctx->prof_on && ctx->prof_used < dfgpu_ctx::PROF_SLOTS) {
    pe0 = ctx->prof_ev[2 * ctx->prof_used];
    pe1 = ctx->prof_ev[2 * ctx->prof_used + 1];
    ctx->prof_bytes[ctx->prof_used] = ctx->cur_nbytes;
    ctx->prof_used++;
    CU(cudaEventRecord(pe0, s));
  }
  if (l7) launch_l7_strings(bp, g_str, s);
  else launch_l4_pack(ctx, s);
  if (pe1) CU(cudaEventRecord(pe1, s));  // the profile events bracket the pack kernel alone
  ctx->launches++;
  if (l7) {
    ctx->mark(5);
    if (kg_side) {
      CU(cudaStreamWaitEvent(s, ctx->ev_join, 0));  // the flow_tag rows read SmartEncoding columns
    } else {
      launch_l7_kg(bp, g_kg, s);  // SmartEncoding columns, one lane per (record, side)
      ctx->launches++;
    }
    ctx->mark(6);
    int rc = l7_flow_tags(ctx, s);  // commits the counters at its end
    if (rc) return rc;
  } else {
    ctx->mark(5);
    launch_l4_kg(ctx, s);
    ctx->mark(6);
    launch_commit(bp, 0, s);
    ctx->launches++;
  }
  CU(cudaGetLastError());
  return 0;
}

// framing + passes of the staged batch
static int run_pipeline(dfgpu_ctx* ctx) {
  cudaStream_t s = ctx->stream;
  const TableSchema* sch = main_schema(ctx->cur_msg);
  BatchParams& bp = ctx->bp;
  memset(&bp, 0, sizeof bp);
  CU(cudaMemsetAsync(ctx->d_dyn, 0, sizeof(DynState), s));
  const uint32_t nf = ctx->cur_nframes;
  bp.in = ctx->cur_in;
  bp.in_bytes = ctx->cur_nbytes;
  bp.frames = ctx->cur_frames;
  bp.nframes = nf;
  bp.frame_nrec = ctx->d_frame_u32;
  bp.frame_base = ctx->d_frame_u32 + ctx->frames_cap;
  bp.frame_chain_err = ctx->d_frame_u32 + 2 * (size_t)ctx->frames_cap;
  bp.frame_bad = ctx->d_frame_u32 + 3 * (size_t)ctx->frames_cap;
  bp.frame_drop = ctx->d_frame_u32 + 4 * (size_t)ctx->frames_cap;
  bp.dyn = ctx->d_dyn;
  bp.persist = ctx->d_persist;
  bp.dict = ctx->cur_dict->view;
  if (ctx->cur_geo) {
    bp.geo.start = ctx->cur_geo->d_start;
    bp.geo.end = ctx->cur_geo->d_end;
    bp.geo.code = ctx->cur_geo->d_code;
    bp.geo.n = ctx->cur_geo->n;
  }
  bp.cfg_trace_disabled = ctx->cfg.trace_id_index_disabled;
  bp.rec_scratch = ctx->d_rec_scratch;
  bp.rec_cap = ctx->rec_cap;
  bp.warp_cap = (uint32_t)ctx->warp_cap;
  bp.rec_off = ctx->d_rec_u32;
  bp.rec_len = ctx->d_rec_u32 + ctx->rec_cap;
  bp.rec_frame = ctx->d_rec_u32 + 2 * (size_t)ctx->rec_cap;
  bp.warp_tot = ctx->d_warp;
  bp.warp_base = ctx->d_warp + (size_t)MAX_SLOTS * ctx->warp_cap;
  bp.arena = ctx->d_arena;
  bp.arena_bytes = ctx->arena_cap;
  bp.native_str = (ctx->cfg.native_strings && ctx->cur_msg == DFGPU_MSG_PROTOCOLLOG) ? 1u : 0u;
  if (sch) {
    setup_columns(bp, *sch);
    setup_streams(bp, ctx->cur_msg);
  }
  if (ctx->cur_msg == DFGPU_MSG_PROTOCOLLOG) {
    int rc = l7_prepare(ctx, s);
    if (rc) return rc;
  } else if (ctx->cur_msg == DFGPU_MSG_TAGGEDFLOW) {
    int rc = l4_prepare(ctx, s);
    if (rc) return rc;
  }
  ctx->mark(0);
  if (nf) {
    CU(cudaMemsetAsync(bp.frame_bad, 0xff, (size_t)nf * 4, s));
    launch_frame_count(bp, s);  // walks the length chains, leaves the record count on the device
    ctx->launches += 2;
  }
  ctx->mark(1);
  return launch_passes(ctx, 0);
}

extern "C" {

int dfgpu_stage(dfgpu_ctx* ctx, int msg_type, const uint8_t* frames, size_t nbytes, const dfgpu_frame* tbl,
                int nframes) {
  if (!ctx || nframes < 0 || (nframes && !tbl) || (nbytes && !frames)) return DFGPU_EINVAL;
  if (!main_schema(msg_type) && msg_type != DFGPU_MSG_METRICS)
    return fail(ctx, DFGPU_EINVAL, "unsupported msg_type %d", msg_type);
  if (nbytes >= (1ull << 31)) return fail(ctx, DFGPU_EINVAL, "batch too large (%zu bytes)", nbytes);
  CU(cudaSetDevice(ctx->device));
  cudaPointerAttributes attr;
  if (tbl && cudaPointerGetAttributes(&attr, tbl) == cudaSuccess && attr.type == cudaMemoryTypeDevice)
    return fail(ctx, DFGPU_EINVAL, "the frame table must be host memory (only `frames` may be device resident)");
  cudaGetLastError();
  uint16_t org = nframes ? tbl[0].org_id : 1;
  for (int i = 1; i < nframes; i++)
    if (tbl[i].org_id != org) return fail(ctx, DFGPU_EMIXEDORG, "frames of one submit must share org_id");
  bool sorted = true;
  for (int i = 0; i < nframes; i++) {
    // offset + len may wrap in uint64: compare without the addition
    if (tbl[i].offset > nbytes || tbl[i].len > nbytes - tbl[i].offset)
      return fail(ctx, DFGPU_EINVAL, "frame %d outside the staging buffer", i);
    if (i && tbl[i].offset < tbl[i - 1].offset + tbl[i - 1].len) sorted = false;
  }
  if (!sorted) {
    // frames may come in any order (they are separate RecvBuffers), but they share framing scratch by byte
    // range: two frames must not overlap
    std::vector<std::pair<uint64_t, uint32_t>> v((size_t)nframes);
    for (int i = 0; i < nframes; i++) v[(size_t)i] = {tbl[i].offset, tbl[i].len};
    std::sort(v.begin(), v.end());
    for (int i = 1; i < nframes; i++)
      if (v[(size_t)i].second && v[(size_t)i].first < v[(size_t)i - 1].first + v[(size_t)i - 1].second)
        return fail(ctx, DFGPU_EINVAL, "frames overlap inside the staging buffer (at offset %llu)", (unsigned long long)v[(size_t)i].first);
  }
  {
    std::lock_guard<std::mutex> g(ctx->dict_mu);
    auto it = ctx->dicts.find(org);
    if (it == ctx->dicts.end()) return fail(ctx, DFGPU_ENOENT, "no dictionary loaded for org %u", org);
    ctx->cur_dict = it->second;
    ctx->cur_geo = ctx->geo;
  }
  bool on_device = false;
  if (frames && cudaPointerGetAttributes(&attr, frames) == cudaSuccess && attr.type == cudaMemoryTypeDevice) on_device = true;
  cudaGetLastError();
  if (on_device && ((uintptr_t)frames & 7))
    return fail(ctx, DFGPU_EINVAL, "device-resident frames must be 8-byte aligned");
  // a batch beyond the configured capacities grows the buffers here, before anything is queued
  if ((uint32_t)nframes > ctx->frames_cap && alloc_frames(ctx, (uint32_t)nframes)) return DFGPU_ENODEV;
  if (nbytes > ctx->d_in_cap &&
      alloc_buffers(ctx, nbytes + nbytes / 4, ctx->rec_cap, (size_t)arena_need(ctx->rec_cap, nbytes + nbytes / 4)))
    return DFGPU_ENODEV;
  if (on_device) {
    ctx->cur_in = frames;  // inputs already resident in HBM
  } else {
    if (nbytes) CU(cudaMemcpyAsync(ctx->d_in, frames, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->cur_in = ctx->d_in;
  }
  if (nframes) CU(cudaMemcpyAsync(ctx->d_frames, tbl, (size_t)nframes * sizeof(dfgpu_frame), cudaMemcpyHostToDevice, ctx->stream));
  ctx->cur_frames = ctx->d_frames;
  ctx->cur_msg = msg_type;
  ctx->cur_nbytes = nbytes;
  ctx->cur_nframes = (uint32_t)nframes;
  ctx->cur_org = org;
  ctx->staged = true;
  ctx->ran = false;
  ctx->metrics_resident = false;
  return DFGPU_OK;
}

// Queues every kernel of the staged batch on the context's stream and returns: no allocation, no
// synchronisation, no device->host copy (the record count and everything sized by it stay on the device).
int dfgpu_run(dfgpu_ctx* ctx, uint32_t* id_counter, uint64_t* trace_id_index) {
  if (!ctx || !ctx->staged) return fail(ctx, DFGPU_EINVAL, "dfgpu_run without dfgpu_stage");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  ctx->id_counter_out = id_counter;
  ctx->trace_idx_out = trace_id_index;
  if (id_counter || trace_id_index) {
    // caller-provided counters win over the device-resident state
    if (id_counter) ctx->h_persist->id_counter = *id_counter;
    if (trace_id_index) ctx->h_persist->last_trace_id_index = *trace_id_index;
    CU(cudaMemcpyAsync(ctx->d_persist, ctx->h_persist, sizeof(PersistState), cudaMemcpyHostToDevice, s));
  }
  CU(cudaEventRecord(ctx->ev0, s));
  ctx->stage_open = ctx->prof_on && ctx->stage_used < dfgpu_ctx::PROF_SLOTS;
  int rc = run_pipeline(ctx);
  if (rc) return rc;
  ctx->mark(7);
  if (ctx->stage_open) {
    ctx->stage_used++;
    ctx->stage_open = false;
  }
  CU(cudaEventRecord(ctx->ev1, s));
  ctx->ran = true;
  return DFGPU_OK;
}

int dfgpu_sync(dfgpu_ctx* ctx) {
  if (!ctx) return DFGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  return DFGPU_OK;
}

}  // extern "C"

static uint8_t* host_acquire(dfgpu_ctx* ctx, size_t bytes) {
  std::unique_lock<std::mutex> lk(ctx->pool_mu);
  size_t best = (size_t)-1;
  for (size_t i = 0; i < ctx->host_pool.size(); i++)
    if (ctx->host_pool[i].cap >= bytes && (best == (size_t)-1 || ctx->host_pool[i].cap < ctx->host_pool[best].cap)) best = i;
  if (best != (size_t)-1) {
    uint8_t* p = ctx->host_pool[best].p;
    // keep the capacity in a small header in front of the user pointer
    ctx->host_pool.erase(ctx->host_pool.begin() + best);
    return p;
  }
  lk.unlock();
  size_t cap = align_up(bytes + bytes / 8 + 4096, 4096);
  uint8_t* p = nullptr;
  if (cudaHostAlloc(&p, cap + 64, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  *reinterpret_cast<size_t*>(p) = cap;
  return p;
}
static void host_release(dfgpu_ctx* ctx, uint8_t* p) {
  if (!p) return;
  HostBuf b;
  b.p = p;
  b.cap = *reinterpret_cast<size_t*>(p);
  std::lock_guard<std::mutex> lk(ctx->pool_mu);
  ctx->host_pool.push_back(b);
}

namespace dfgpu {
// filled in by l7.cu: appends the flow_tag / app_service tables of a fetched batch
int l7_fetch_tables(dfgpu_ctx* ctx, dfgpu_batch* b);
}

// dictionary of one LowCardinality code space, in first-appearance order
static void lc_dictionary(dfgpu_batch* b, const uint32_t* first, const std::vector<std::string>& names_by_code,
                          dfgpu_col& col) {
  std::vector<std::pair<uint32_t, uint32_t>> present;
  for (uint32_t c = 0; c < 256; c++)
    if (first[c] != NO_BAD) present.push_back({first[c], c});
  std::sort(present.begin(), present.end());
  b->owned.emplace_back();
  b->owned_pos.emplace_back();
  auto& buf = b->owned.back();
  auto& pos = b->owned_pos.back();
  for (auto& pr : present) {
    const std::string& s = pr.second < names_by_code.size() ? names_by_code[pr.second] : std::string();
    dfgpu_pos p{(int64_t)buf.size(), (int64_t)(buf.size() + s.size())};
    buf.insert(buf.end(), s.begin(), s.end());
    pos.push_back(p);
  }
  col.dict_data = buf.data();
  col.dict_pos = pos.data();
  col.dict_n = pos.size();
}

namespace dfgpu {
int l7_lc_space(int col);
const std::vector<std::string>& l7_lc_names(dfgpu_ctx* ctx, int space);
}

extern "C" {

int dfgpu_fetch(dfgpu_ctx* ctx, dfgpu_batch** out) {
  if (!ctx || !out || !ctx->ran) return fail(ctx, DFGPU_EINVAL, "dfgpu_fetch without dfgpu_run");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  BatchParams& bp = ctx->bp;
  const TableSchema* sch = main_schema(ctx->cur_msg);
  const uint32_t nf = ctx->cur_nframes;
  // The first (and, for a batch within the configured capacities, only) wait for the GPU: sizes, offsets,
  // error flags.  A batch that outgrew a device buffer, or in which a record failed to decode, is re-run.
  for (int attempt = 0;; attempt++) {
    CU(cudaMemcpyAsync(ctx->h_dyn, ctx->d_dyn, sizeof(DynState), cudaMemcpyDeviceToHost, s));
    if (nf) {
      CU(cudaMemcpyAsync(ctx->h_frame_u32, bp.frame_chain_err, (size_t)nf * 4, cudaMemcpyDeviceToHost, s));
      CU(cudaMemcpyAsync(ctx->h_frame_u32 + ctx->h_frames_cap, bp.frame_bad, (size_t)nf * 4, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    DynState* d = ctx->h_dyn;
    const bool repair = d->stats.need_repair && !bp.repair;
    const bool grow = d->overflow != 0;
    if (!repair && !grow) break;
    if (attempt >= 4) return fail(ctx, DFGPU_ENOMEM, "could not size the device buffers for this batch");
    if (grow) {
      // the batch exceeds dfgpu_cfg.max_batch_records / the arena planned from it: grow once, re-run all
      uint32_t rec_cap = ctx->rec_cap;
      if (d->rec_overflow || d->stats.nrec > rec_cap) rec_cap = d->stats.nrec + d->stats.nrec / 4 + 1024;
      uint64_t arena = std::max<uint64_t>(std::max(d->dev_end, d->var_end) + (16u << 20), arena_need(rec_cap, ctx->cur_nbytes));
      if (arena <= ctx->arena_cap) arena = ctx->arena_cap + ctx->arena_cap / 2;
      int rc = alloc_buffers(ctx, ctx->d_in_cap, rec_cap, (size_t)arena);
      if (rc) return rc;
      rc = run_pipeline(ctx);
      if (rc) return rc;
    } else {
      // a record failed to decode: everything behind it in its frame is dropped
      // (decoder.go:206-209); re-derive rows and sizes with the per-frame cut known.
      const uint32_t keep_nrec = d->stats.nrec;
      CU(cudaMemsetAsync(ctx->d_dyn, 0, sizeof(DynState), s));
      CU(cudaMemcpyAsync(&ctx->d_dyn->stats.nrec, &keep_nrec, 4, cudaMemcpyHostToDevice, s));
      int rc = launch_passes(ctx, 1);
      if (rc) return rc;
    }
    CU(cudaEventRecord(ctx->ev1, s));
  }
  if (ctx->cur_msg == DFGPU_MSG_METRICS) {
    dfgpu_batch* b = new dfgpu_batch();
    b->ctx = ctx;
    b->msg_type = ctx->cur_msg;
    memset(&b->counters, 0, sizeof b->counters);
    int rc = metrics_fetch(ctx, b, host_acquire);
    if (rc) {
      dfgpu_batch_release(b);
      return rc;
    }
    b->counters.raw_count = nf;
    for (uint32_t f = 0; f < nf; f++)  // ErrDocCount: one per abandoned frame (unmarshaller.go:236-240)
      if (ctx->h_frame_u32[f] || ctx->h_frame_u32[ctx->h_frames_cap + f] != NO_BAD) b->counters.error_count++;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) b->counters.device_us = ms * 1000.0;
    ctx->ran = false;
    ctx->metrics_resident = true;
    *out = b;
    return DFGPU_OK;
  }
  DynState* d = ctx->h_dyn;
  uint32_t nrows = d->stats.nrows;
  uint64_t total = d->var_end ? d->var_end : d->var_base;
  dfgpu_batch* b = new dfgpu_batch();
  b->ctx = ctx;
  b->msg_type = ctx->cur_msg;
  b->stats = d->stats;
  // dfgpu_cfg.block_lz4: the arena stays on the device; the column views are first laid out against a virtual
  // base, then every buffer is compressed into frames and the views are re-pointed at the frame streams
  const bool lz = ctx->cfg.block_lz4 != 0 && nrows > 0;
  uint8_t* const vbase = reinterpret_cast<uint8_t*>(uintptr_t(1) << 44);
  uint8_t* hb = vbase;
  if (!lz) {
    b->host = host_acquire(ctx, (size_t)total + 64);
    if (!b->host) {
      delete b;
      return fail(ctx, DFGPU_ENOMEM, "pinned host allocation of %llu bytes failed", (unsigned long long)total);
    }
    hb = b->host + 64;
  }
  // D2H.  Common case (every record produced a row): the fixed region was planned for exactly these
  // rows, so the whole used arena is dense -> ONE copy instead of ~150 small ones.  Otherwise: per
  // column only the rows that exist; streams only their used bytes.
  const bool dense = nrows > 0 && (nrows == d->stats.nrec || (d->plan_valid && nrows == d->plan_rows));
  if (dense && !lz) CU(cudaMemcpyAsync(hb, bp.arena, (size_t)total, cudaMemcpyDeviceToHost, s));
  for (int c = 0; c < sch->ncols && !dense && !lz; c++) {
    const ColSpec& cs = sch->cols[c];
    size_t es = is_array(cs.ch_type) ? 0 : (size_t)cs.elem_size;
    if (es && nrows) CU(cudaMemcpyAsync(hb + d->fix_off[c], bp.arena + d->fix_off[c], (size_t)nrows * es, cudaMemcpyDeviceToHost, s));
    size_t ab = aux_bytes_per_row(cs.ch_type);
    if (bp.native_str && (cs.ch_type == DFGPU_CH_STR || cs.ch_type == DFGPU_CH_LC_STR_ROWS)) ab = 0;  // no Positions
    if (ab && nrows) CU(cudaMemcpyAsync(hb + d->aux_off[c], bp.arena + d->aux_off[c], (size_t)nrows * ab, cudaMemcpyDeviceToHost, s));
  }
  if (!dense && !lz && total > d->var_base)
    CU(cudaMemcpyAsync(hb + d->var_base, bp.arena + d->var_base, (size_t)(total - d->var_base), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(ctx->h_persist, ctx->d_persist, sizeof(PersistState), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (ctx->id_counter_out) *ctx->id_counter_out = ctx->h_persist->id_counter;
  if (ctx->trace_idx_out) *ctx->trace_idx_out = ctx->h_persist->last_trace_id_index;

  // ---- column views ------------------------------------------------------------------
  b->tables.emplace_back();
  dfgpu_batch::Table& t = b->tables.back();
  t.id = sch->table_id;
  t.nrows = nrows;
  t.cols.resize(sch->ncols);
  b->owned.reserve(64);
  b->owned_pos.reserve(64);
  int stream_of_col = 0;
  for (int c = 0; c < sch->ncols; c++) {
    const ColSpec& cs = sch->cols[c];
    dfgpu_col& col = t.cols[c];
    memset(&col, 0, sizeof col);
    col.name = cs.name;
    col.ch_type = cs.ch_type;
    col.elem_size = cs.elem_size;
    col.nrows = nrows;
    switch (cs.ch_type) {
      case DFGPU_CH_STR:
      case DFGPU_CH_LC_STR_ROWS: {
        int v = stream_of_col++;
        col.pos = bp.native_str ? nullptr : (const dfgpu_pos*)(hb + d->aux_off[c]);
        col.data = hb + d->var_buf_off[v];
        col.data_bytes = d->stats.var_bytes[v];
        break;
      }
      case DFGPU_CH_ARR_UINT16:
      case DFGPU_CH_ARR_FLOAT64: {
        int v = stream_of_col++;
        col.arr_offsets = (const uint64_t*)(hb + d->aux_off[c]);
        col.data = hb + d->var_buf_off[v];
        col.data_bytes = d->stats.var_bytes[v];
        col.nelems = d->stats.var_total[v];
        break;
      }
      case DFGPU_CH_ARR_STR:
      case DFGPU_CH_ARR_LC_STR: {
        int v = stream_of_col++;
        col.arr_offsets = (const uint64_t*)(hb + d->aux_off[c]);
        col.pos = bp.native_str ? nullptr : (const dfgpu_pos*)(hb + d->var_pos_off[v]);
        col.data = hb + d->var_buf_off[v];
        col.data_bytes = d->stats.var_bytes[v];
        col.nelems = d->stats.var_total[v];
        break;
      }
      case DFGPU_CH_LC_STR: {
        col.data = hb + d->fix_off[c];
        col.data_bytes = (uint64_t)nrows * cs.elem_size;
        if (ctx->cur_msg == DFGPU_MSG_TAGGEDFLOW) {
          static const std::vector<std::string> no_names;
          const std::vector<std::string>& geo_names = ctx->cur_geo ? ctx->cur_geo->names : no_names;
          if (c == L4C_PROVINCE_0) lc_dictionary(b, d->stats.lc_first[LC_PROVINCE_0], geo_names, col);
          else if (c == L4C_PROVINCE_1) lc_dictionary(b, d->stats.lc_first[LC_PROVINCE_1], geo_names, col);
          else {
            static const std::vector<std::string> tap(kTapSideNames, kTapSideNames + 18);
            lc_dictionary(b, d->stats.lc_first[LC_TAPSIDE], tap, col);
          }
        } else {
          int sp = l7_lc_space(c);
          lc_dictionary(b, d->stats.lc_first[sp], l7_lc_names(ctx, sp), col);
        }
        break;
      }
      default: {
        col.data = hb + d->fix_off[c];
        col.data_bytes = (uint64_t)nrows * cs.elem_size;
        if (aux_bytes_per_row(cs.ch_type) == 1) col.nulls = hb + d->aux_off[c];
        break;
      }
    }
  }
  if (lz) {
    // one region per buffer of every column, in column order
    std::vector<LzRegion> regs;
    struct Ref { int col, which; };
    std::vector<Ref> refs;
    auto add = [&](int c, int which, const void* ptr, uint64_t len, uint32_t stride) {
      if (!ptr || !len) return;
      regs.push_back(LzRegion{(uint64_t)(reinterpret_cast<const uint8_t*>(ptr) - vbase), len, 0u, stride});
      refs.push_back(Ref{c, which});
    };
    for (int c = 0; c < sch->ncols; c++) {
      const dfgpu_col& col = t.cols[c];
      const bool arr = is_array(col.ch_type);
      const bool str = col.ch_type == DFGPU_CH_STR || col.ch_type == DFGPU_CH_LC_STR_ROWS || col.ch_type == DFGPU_CH_ARR_STR ||
                       col.ch_type == DFGPU_CH_ARR_LC_STR;
      const uint32_t es = col.ch_type == DFGPU_CH_ARR_UINT16 ? 2u : col.ch_type == DFGPU_CH_ARR_FLOAT64 ? 8u : (uint32_t)col.elem_size;
      add(c, 0, col.data, col.data_bytes, str ? 0u : (es > 16 ? 16u : es));
      add(c, 1, col.pos, (arr ? col.nelems : col.nrows) * 16, 16u);
      add(c, 2, col.arr_offsets, col.nrows * 8, 8u);
      add(c, 3, col.nulls, col.nrows, 1u);
    }
    const uint8_t* d_frames = nullptr;
    std::vector<uint64_t> fpos;
    int rc = lz4_compress_regions(ctx, bp.arena, regs, &d_frames, fpos);
    if (rc) {
      delete b;
      return rc;
    }
    const uint64_t ftotal = fpos.empty() ? 0 : fpos.back();
    b->host = host_acquire(ctx, (size_t)ftotal + 64);
    if (!b->host) {
      delete b;
      return fail(ctx, DFGPU_ENOMEM, "pinned host allocation of %llu bytes failed", (unsigned long long)ftotal);
    }
    uint8_t* fb = b->host + 64;
    if (ftotal) CU(cudaMemcpyAsync(fb, d_frames, (size_t)ftotal, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (auto& col : t.cols) {
      col.compressed = 1;
      col.data_raw_bytes = col.data_bytes;
      // buffers without bytes (an empty stream) keep no frames
      col.data = nullptr;
      col.pos = nullptr;
      col.arr_offsets = nullptr;
      col.nulls = nullptr;
      col.data_bytes = 0;
    }
    for (size_t r = 0; r < regs.size(); r++) {
      dfgpu_col& col = t.cols[refs[r].col];
      const uint8_t* p = fb + fpos[r];
      const uint64_t n = fpos[r + 1] - fpos[r];
      switch (refs[r].which) {
        case 0: col.data = p; col.data_bytes = n; break;
        case 1: col.pos = reinterpret_cast<const dfgpu_pos*>(p); col.pos_bytes = n; break;
        case 2: col.arr_offsets = reinterpret_cast<const uint64_t*>(p); col.arr_offsets_bytes = n; break;
        default: col.nulls = p; col.nulls_bytes = n; break;
      }
    }
    b->lz_frame_bytes = ftotal;
  }
  // ---- counters (decoder.Counter, decoder.go:60-80) ------------------------------------
  dfgpu_counters& cn = b->counters;
  memset(&cn, 0, sizeof cn);
  cn.raw_count = nf;
  cn.count = nrows;
  for (uint32_t f = 0; f < nf; f++)
    if (ctx->h_frame_u32[f] || ctx->h_frame_u32[ctx->h_frames_cap + f] != NO_BAD) cn.error_count++;
  cn.l7_http_count = d->stats.l7_counts[0];
  cn.l7_dns_count = d->stats.l7_counts[1];
  cn.l7_sql_count = d->stats.l7_counts[2];
  cn.l7_nosql_count = d->stats.l7_counts[3];
  cn.l7_rpc_count = d->stats.l7_counts[4];
  cn.l7_mq_count = d->stats.l7_counts[5];
  cn.rows[sch->table_id] = nrows;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) cn.device_us = ms * 1000.0;
  if (ctx->cur_msg == DFGPU_MSG_PROTOCOLLOG) {
    int rc = l7_fetch_tables(ctx, b);
    if (rc) {
      dfgpu_batch_release(b);
      return rc;
    }
  }
  ctx->ran = false;
  *out = b;
  return DFGPU_OK;
}

int dfgpu_submit(dfgpu_ctx* ctx, int msg_type, const uint8_t* frames, size_t nbytes, const dfgpu_frame* tbl,
                 int nframes, uint32_t* id_counter, uint64_t* trace_id_index, dfgpu_batch** out) {
  int rc = dfgpu_stage(ctx, msg_type, frames, nbytes, tbl, nframes);
  if (rc) return rc;
  rc = dfgpu_run(ctx, id_counter, trace_id_index);
  if (rc) return rc;
  return dfgpu_fetch(ctx, out);
}

int dfgpu_batch_columns(const dfgpu_batch* b, int table, dfgpu_col* cols, int* ncols, uint64_t* nrows) {
  if (!b) return DFGPU_EINVAL;
  for (auto& t : b->tables)
    if (t.id == table) {
      if (ncols) *ncols = (int)t.cols.size();
      if (nrows) *nrows = t.nrows;
      if (cols) memcpy(cols, t.cols.data(), t.cols.size() * sizeof(dfgpu_col));
      return DFGPU_OK;
    }
  return DFGPU_ENOENT;
}

int dfgpu_batch_counters(const dfgpu_batch* b, dfgpu_counters* out) {
  if (!b || !out) return DFGPU_EINVAL;
  *out = b->counters;
  return DFGPU_OK;
}

void dfgpu_batch_release(dfgpu_batch* b) {
  if (!b) return;
  host_release(b->ctx, b->host);
  delete b;
}

int dfgpu_flush(dfgpu_ctx* ctx) { return dfgpu_sync(ctx); }

// Measurement hook (bench.py roofline leg): CUDA events recorded on the ctx stream
// right around the pack kernels (string / array columns) of every dfgpu_run.
int dfgpu_profile_enable(dfgpu_ctx* ctx, int on) {
  if (!ctx) return DFGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (on && !ctx->prof_ev[0])
    for (int i = 0; i < 2 * dfgpu_ctx::PROF_SLOTS; i++) CU(cudaEventCreate(&ctx->prof_ev[i]));
  if (on && ctx->stage_ev.empty()) {
    ctx->stage_ev.resize((size_t)dfgpu_ctx::PROF_SLOTS * dfgpu_ctx::STAGE_MARKS);
    for (auto& e : ctx->stage_ev) CU(cudaEventCreate(&e));
  }
  ctx->prof_on = on != 0;
  ctx->prof_used = 0;
  ctx->stage_used = 0;
  return DFGPU_OK;
}
// Sums and resets the stage times of the profiled runs (milliseconds):
// framing, prepare, decode pass, scans, pack pass, SmartEncoding, flow_tag stage.
int dfgpu_profile_stages(dfgpu_ctx* ctx, double* ms6, uint64_t* runs) {
  if (!ctx || !ms6) return DFGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < 7; k++) ms6[k] = 0;
  for (int r = 0; r < ctx->stage_used; r++)
    for (int k = 0; k < 7; k++) {
      float t = 0;
      CU(cudaEventElapsedTime(&t, ctx->stage_ev[(size_t)r * dfgpu_ctx::STAGE_MARKS + k],
                              ctx->stage_ev[(size_t)r * dfgpu_ctx::STAGE_MARKS + k + 1]));
      ms6[k] += t;
    }
  if (runs) *runs = (uint64_t)ctx->stage_used;
  ctx->stage_used = 0;
  return DFGPU_OK;
}
// Sums and resets: emit-kernel milliseconds, launches, input bytes of those launches.
int dfgpu_profile_read(dfgpu_ctx* ctx, double* emit_ms, uint64_t* launches, uint64_t* in_bytes) {
  if (!ctx) return DFGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->stream));
  double ms = 0;
  uint64_t bytes = 0;
  for (int i = 0; i < ctx->prof_used; i++) {
    float t = 0;
    CU(cudaEventElapsedTime(&t, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
    ms += t;
    bytes += ctx->prof_bytes[i];
  }
  if (emit_ms) *emit_ms = ms;
  if (launches) *launches = (uint64_t)ctx->prof_used;
  if (in_bytes) *in_bytes = bytes;
  ctx->prof_used = 0;
  return DFGPU_OK;
}

int dfgpu_batch_tag_fingerprints(const dfgpu_batch* b, int table, const uint64_t** fps, uint64_t* n) {
  if (!b || !fps || !n) return DFGPU_EINVAL;
  for (auto& t : b->tag_fps)
    if (t.table == table) {
      *fps = t.fps;
      *n = t.n;
      return DFGPU_OK;
    }
  return DFGPU_ENOENT;
}

int dfgpu_tags_reset(dfgpu_ctx* ctx) {
  if (!ctx) return DFGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  return l7_tags_reset(ctx);
}

int dfgpu_tags_import(dfgpu_ctx* ctx, const uint64_t* fps, const uint32_t* times, uint64_t n) {
  if (!ctx || (n && (!fps || !times))) return DFGPU_EINVAL;
  CU(cudaSetDevice(ctx->device));
  return l7_tags_import(ctx, fps, times, n);
}

static int metrics_table_arg(dfgpu_ctx* ctx, int table) {
  if (table < DFGPU_TABLE_METRICS_BASE || table >= DFGPU_TABLE_METRICS_BASE + 9)
    return fail(ctx, DFGPU_EINVAL, "table %d is not a flow_metrics table", table);
  return 0;
}

int dfgpu_rollup_add(dfgpu_ctx* ctx, int table, uint32_t bucket_seconds) {
  if (!ctx || bucket_seconds == 0) return DFGPU_EINVAL;
  if (int rc = metrics_table_arg(ctx, table)) return rc;
  if (!ctx->metrics_resident) return fail(ctx, DFGPU_ENOENT, "no metrics batch resident on the device (call after dfgpu_submit / dfgpu_fetch)");
  CU(cudaSetDevice(ctx->device));
  return rollup_add_resident(ctx, table - DFGPU_TABLE_METRICS_BASE, bucket_seconds);
}

int dfgpu_rollup_add_rows(dfgpu_ctx* ctx, int table, uint32_t bucket_seconds, const dfgpu_rollup_rows* rows) {
  if (!ctx || !rows || bucket_seconds == 0 || (rows->nrows && (!rows->tag_cols || !rows->raw))) return DFGPU_EINVAL;
  if (int rc = metrics_table_arg(ctx, table)) return rc;
  CU(cudaSetDevice(ctx->device));
  return rollup_add_rows(ctx, table - DFGPU_TABLE_METRICS_BASE, bucket_seconds, rows);
}

int dfgpu_rollup_describe(dfgpu_ctx* ctx, int table, int* ntag, int* nfields, uint8_t* elem_sizes, uint8_t* kinds) {
  if (!ctx) return DFGPU_EINVAL;
  if (int rc = metrics_table_arg(ctx, table)) return rc;
  CU(cudaSetDevice(ctx->device));
  return rollup_describe(ctx, table - DFGPU_TABLE_METRICS_BASE, ntag, nfields, elem_sizes, kinds);
}

int dfgpu_rollup_flush(dfgpu_ctx* ctx, int table, dfgpu_batch** out) {
  if (!ctx || !out) return DFGPU_EINVAL;
  if (int rc = metrics_table_arg(ctx, table)) return rc;
  CU(cudaSetDevice(ctx->device));
  dfgpu_batch* b = new dfgpu_batch();
  b->ctx = ctx;
  b->msg_type = DFGPU_MSG_METRICS;
  memset(&b->counters, 0, sizeof b->counters);
  int rc = rollup_flush(ctx, table - DFGPU_TABLE_METRICS_BASE, b, host_acquire);
  if (rc) {
    dfgpu_batch_release(b);
    return rc;
  }
  *out = b;
  return DFGPU_OK;
}

// one-shot form: the rows of the last metrics submit, reduced and returned (= add + flush; groups an
// earlier dfgpu_rollup_add left in the table's accumulator are folded in and flushed with them)
int dfgpu_metrics_reduce(dfgpu_ctx* ctx, int table, uint32_t bucket_seconds, dfgpu_batch** out) {
  int rc = dfgpu_rollup_add(ctx, table, bucket_seconds);
  if (rc) return rc;
  return dfgpu_rollup_flush(ctx, table, out);
}

}  // extern "C"
