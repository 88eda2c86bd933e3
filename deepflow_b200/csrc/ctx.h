/*
 * ctx.h — the context object behind dfgpu_ctx* (one per decoder goroutine).
 */
#pragma once
#include <cstdarg>
#include <mutex>

#include "pipeline.cuh"

using namespace dfgpu;

struct GeoTables {
  uint32_t* d_start = nullptr;  // one allocation: start[n] | end[n] | code[n]
  uint32_t* d_end = nullptr;
  uint8_t* d_code = nullptr;
  uint32_t n = 0;
  std::vector<std::string> names;  // by canonical code
  ~GeoTables() {
    if (d_start) cudaFree(d_start);
  }
};

struct HostBuf {
  uint8_t* p = nullptr;
  size_t cap = 0;
};

struct dfgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // side stream of the SmartEncoding kernel: it depends on the decode kernel only, so it runs beside the scans and the
  // string kernels (kg_overlap) and joins before the flow_tag stage
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int kg_overlap = 0, kg_ctas = 0, str_ctas = 0;  // 0: defaults
  dfgpu_cfg cfg{};
  std::string err;
  std::mutex dict_mu;
  std::map<uint16_t, std::shared_ptr<DeviceDict>> dicts;
  // geo: an immutable snapshot, swapped under dict_mu (dfgpu_geo_load may run on another thread);
  // a batch keeps the snapshot it was staged with alive until its fetch is over
  std::shared_ptr<GeoTables> geo;
  std::shared_ptr<GeoTables> cur_geo;

  // device work buffers
  uint8_t* d_in = nullptr;
  size_t d_in_cap = 0;
  dfgpu_frame* d_frames = nullptr;
  uint32_t* d_frame_u32 = nullptr;  // nrec | base | chain_err | bad
  uint32_t frames_cap = 0;
  uint32_t* d_rec_scratch = nullptr;  // framing scratch, cur_nbytes/4 + 2 entries
  size_t rec_scratch_cap = 0;
  uint32_t* d_rec_u32 = nullptr;  // off | len | frame
  uint32_t rec_cap = 0;
  uint32_t* d_warp = nullptr;  // tot[MAX_SLOTS][nwarps] | base[MAX_SLOTS][nwarps]
  size_t warp_cap = 0;         // in warps
  DynState* d_dyn = nullptr;
  PersistState* d_persist = nullptr;
  uint8_t* d_arena = nullptr;
  size_t arena_cap = 0;
  // pinned host
  DynState* h_dyn = nullptr;
  PersistState* h_persist = nullptr;
  uint32_t* h_frame_u32 = nullptr;
  uint32_t h_frames_cap = 0;
  std::mutex pool_mu;  // batches may be released from a writer thread while the decoder thread fetches
  std::vector<HostBuf> host_pool;

  // current batch
  int cur_msg = 0;
  const uint8_t* cur_in = nullptr;
  const dfgpu_frame* cur_frames = nullptr;  // device
  size_t cur_nbytes = 0;
  uint32_t cur_nframes = 0;
  uint16_t cur_org = 0;
  uint32_t cur_team = 0;
  bool staged = false, ran = false;
  BatchParams bp{};
  std::shared_ptr<DeviceDict> cur_dict;
  uint32_t* id_counter_out = nullptr;
  uint64_t* trace_idx_out = nullptr;
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // optional per-kernel timing (dfgpu_profile_enable)
  static constexpr int PROF_SLOTS = 256;
  bool prof_on = false;
  int prof_used = 0;
  cudaEvent_t prof_ev[2 * PROF_SLOTS] = {nullptr};
  uint64_t prof_bytes[PROF_SLOTS] = {0};
  // stage marks of every dfgpu_run while profiling: start, framing, prepare, decode, scans, pack, SmartEncoding, flow_tag
  static constexpr int STAGE_MARKS = 8;
  std::vector<cudaEvent_t> stage_ev;  // [PROF_SLOTS][STAGE_MARKS], created on enable
  int stage_used = 0;
  bool stage_open = false;
  void mark(int k) {
    if (prof_on && stage_open && !stage_ev.empty()) cudaEventRecord(stage_ev[(size_t)stage_used * STAGE_MARKS + k], stream);
  }

  // flow tag state (l7)
  void* ft = nullptr;
  // L4 per-record hand-over buffers (l4.cu)
  void* l4 = nullptr;
  // compressed-frame stage (lz4.cu)
  void* lz = nullptr;
  // flow_metrics state (metrics.cu)
  void* mt = nullptr;
  // flow_metrics rollup accumulators (rollup.cu)
  void* ru = nullptr;
  bool metrics_resident = false;  // the arena still holds the tables of the last metrics batch
};

inline int fail(dfgpu_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}
#define CU(call)                                                                       \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess)                                                             \
      return fail(ctx, DFGPU_ENODEV, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

