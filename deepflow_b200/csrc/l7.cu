/*
 * l7.cu — the L7 flow-log kernels: AppProtoLogsData record -> one l7_flow_log row,
 * plus flow_tag discovery.  One thread per record (warp = 32 consecutive records).
 * Replaces, fused:
 *   decode          gogo Unmarshal via SimpleDecoder.ReadPB (simple_codec.go:229-248)
 *   row build       ProtoLogToL7FlowLog  flow_log/log_data/l7_flow_log.go:638-644,
 *                   L7FlowLog.Fill :312-342, L7Base.Fill :514-597, fillL7FlowLog :363-452,
 *                   fillTraceIds :344-360, fillExceptionDesc :463-488 (exception_desc.go),
 *                   L7Protocol.String datatype/flow.go:623-803,
 *                   ParseTraceIdIndex :296-310 -> DJBHash libs/utils/hash.go:37-43
 *   SmartEncoding   KnowledgeGraph.FillL7 :599-613 -> kg.cuh ; QueryProcessInfo :589-596
 *   column pack     L7FlowLog.AppendToColumnBlock l7_column_block.go:167-208,376-422
 *   flow_tag        GenerateNewFlowTags :648-745 with lru.Cache AddOrGet/Add semantics
 *                   (libs/lru/lru.go:44-76), FlowTag.AppendToColumnBlock
 *                   flow_tag_column_block.go:79-101, AppServiceTagWriter.Write
 *                   app_service_tag_writer.go:92-113
 *   counters        updateCounter decoder.go:498-525
 */
#include "l7_dev.h"
#include "l7_tables_host.h"

namespace dfgpu {

struct L7Host {  // host mirror kept in ctx->ft
  L7Dev dev{};
  L7Dev* d_dev = nullptr;
  uint8_t* d_blob = nullptr;
  uint32_t* d_off = nullptr;
  TagSet set;
  uint8_t* d_import = nullptr;
  size_t import_cap = 0;
  uint8_t* d_ft_arena = nullptr;
  size_t ft_arena_cap = 0;
  // per-record hand-over buffers of the decode kernel (one allocation, carved for rec_cap records)
  uint8_t* d_rec_buf = nullptr;
  size_t rec_cap = 0;
  uint64_t* d_raw_scratch = nullptr;
  unsigned grid_decode = 0, grid_strings = 0, grid_kg = 0;
  uint32_t epoch = 0;
};

namespace {

}  // namespace

void l7_streams(BatchParams& bp) {
  bp.nslots = L7S_COUNT;
  bp.nstreams = L7V_COUNT;
  bp.n_lc = 1;
  for (int v = 0; v < 22; v++) bp.streams[v] = StreamSpec{-1, (int8_t)(L7S_STR0 + v), 1, 0};
  bp.streams[L7V_ATTR_NAMES] = StreamSpec{L7S_AN_ELEMS, L7S_AN_BYTES, 1, 1};
  bp.streams[L7V_ATTR_VALUES] = StreamSpec{L7S_AV_ELEMS, L7S_AV_BYTES, 1, 1};
  bp.streams[L7V_METRICS_NAMES] = StreamSpec{L7S_MN_ELEMS, L7S_MN_BYTES, 1, 1};
  bp.streams[L7V_METRICS_VALUES] = StreamSpec{L7S_MV_ELEMS, -1, 8, 0};
  // events: always empty; native encoding = one zero length byte per row
  bp.streams[L7V_EVENTS] = StreamSpec{-1, -1, (uint8_t)(bp.native_str ? 1 : 0), 0};
}

int l7_lc_space(int) { return LC7_TAPSIDE; }
const std::vector<std::string>& l7_lc_names(dfgpu_ctx*, int) {
  static const std::vector<std::string> tap(kTapSideNames, kTapSideNames + 18);
  return tap;
}

// =====================================================================================
// flow_tag discovery (columnar, runs on the freshly packed l7 columns)
// =====================================================================================
namespace {

// slots of the flow_tag stage scans
enum { FT_FV_ROWS = 0, FT_FV_NAME_B = 1, FT_FV_VAL_B = 2, FT_F_ROWS = 3, FT_F_NAME_B = 4, FT_AS_ROWS = 5, FT_AS_B = 6, FT_SLOTS = 7 };
// ft_off[] entries
enum {
  FO_FV_TIME = 0, FO_FV_VPC, FO_FV_PODNS, FO_FV_FTYPE, FO_FV_FVTYPE, FO_FV_TEAM, FO_FV_TABLE, FO_FV_NAME_POS,
  FO_FV_VAL_POS, FO_FV_COUNT, FO_FV_NAME_BUF, FO_FV_VAL_BUF,
  FO_F_TIME, FO_F_VPC, FO_F_PODNS, FO_F_FTYPE, FO_F_FVTYPE, FO_F_TEAM, FO_F_TABLE, FO_F_NAME_POS, FO_F_NAME_BUF,
  FO_AS_TIME, FO_AS_TABLE, FO_AS_SVC_POS, FO_AS_INST_POS, FO_AS_TEAM, FO_AS_BUF,
  FO_FV_FP, FO_F_FP, FO_AS_FP,  // 128-bit key fingerprint of every emitted row (multi-GPU tag merge)
  FO_COUNT
};

// find-or-insert in this context's tag set (tagset.cuh)
__device__ __forceinline__ FtEntry* ft_upsert(const L7Dev* d, U128 fp) { return ts_upsert(d->ft, d->ft_mask, fp, d->fts); }
__device__ __forceinline__ FtEntry* ft_find(const L7Dev* d, U128 fp) { return ts_find(d->ft, d->ft_mask, fp); }

// lru semantics (l7_flow_log.go:693-701): emit when the key is new, or when
// old + CacheFlushTimeout < time (uint32 arithmetic)
__device__ __forceinline__ bool ft_candidate(const FtEntry* e, uint32_t now, uint32_t timeout) {
  if (e->state == 0) return true;
  return (uint32_t)(e->time + timeout) < now;
}

enum { KIND_FIELD_VALUE = 1, KIND_FIELD = 2, KIND_APP_SERVICE = 3 };

struct RowView {  // one packed l7 row, as the flow_tag stage sees it
  uint32_t time;  // uint32(EndTime_us / 1e6)
  uint32_t time_as;  // L7Base.Time (app service tag)
  int32_t vpc[2];
  uint16_t pod_ns[2];
  uint16_t team;
  uint32_t n_sides;
  uint64_t an_lo, an_hi, av_lo, av_hi, mn_lo, mn_hi;  // element ranges
  const dfgpu_pos *an_pos, *av_pos, *mn_pos;
  const uint8_t *an_buf, *av_buf, *mn_buf;
  SRef app_service, endpoint;
};

__device__ __forceinline__ void load_row(const BatchParams& p, uint32_t row, RowView& v) {
  const DynState* dyn = p.dyn;
  auto fix = [&](int col) { return p.arena + dyn->fix_off[col]; };
  auto aux = [&](int col) { return p.arena + dyn->aux_off[col]; };
  int64_t end_us = reinterpret_cast<const int64_t*>(fix(L7C_END_TIME))[row];
  v.time = (uint32_t)(end_us / 1000000);
  v.time_as = reinterpret_cast<const uint32_t*>(fix(L7C_TIME))[row];
  v.vpc[0] = reinterpret_cast<const int32_t*>(fix(L7C_L3_EPC_ID_0))[row];
  v.vpc[1] = reinterpret_cast<const int32_t*>(fix(L7C_L3_EPC_ID_1))[row];
  v.pod_ns[0] = reinterpret_cast<const uint16_t*>(fix(L7C_POD_NS_ID_0))[row];
  v.pod_ns[1] = reinterpret_cast<const uint16_t*>(fix(L7C_POD_NS_ID_1))[row];
  v.team = reinterpret_cast<const uint16_t*>(fix(L7C_TEAM_ID))[row];
  v.n_sides = (v.vpc[0] == v.vpc[1] && v.pod_ns[0] == v.pod_ns[1]) ? 1 : 2;
  const uint64_t* ao = reinterpret_cast<const uint64_t*>(aux(L7C_ATTRIBUTE_NAMES));
  const uint64_t* vo = reinterpret_cast<const uint64_t*>(aux(L7C_ATTRIBUTE_VALUES));
  const uint64_t* mo = reinterpret_cast<const uint64_t*>(aux(L7C_METRICS_NAMES));
  v.an_lo = row ? ao[row - 1] : 0; v.an_hi = ao[row];
  v.av_lo = row ? vo[row - 1] : 0; v.av_hi = vo[row];
  v.mn_lo = row ? mo[row - 1] : 0; v.mn_hi = mo[row];
  v.an_pos = reinterpret_cast<const dfgpu_pos*>(p.arena + dyn->var_pos_off[L7V_ATTR_NAMES]);
  v.av_pos = reinterpret_cast<const dfgpu_pos*>(p.arena + dyn->var_pos_off[L7V_ATTR_VALUES]);
  v.mn_pos = reinterpret_cast<const dfgpu_pos*>(p.arena + dyn->var_pos_off[L7V_METRICS_NAMES]);
  v.an_buf = p.arena + dyn->var_buf_off[L7V_ATTR_NAMES];
  v.av_buf = p.arena + dyn->var_buf_off[L7V_ATTR_VALUES];
  v.mn_buf = p.arena + dyn->var_buf_off[L7V_METRICS_NAMES];
  // the Positions of the per-row strings: column aux region, or (native strings) the device-only region
  const uint8_t* sp_base = p.native_str ? p.arena + dyn->var_pos_off[L7V_APP_SERVICE] : aux(L7C_APP_SERVICE);
  const uint8_t* ep_base = p.native_str ? p.arena + dyn->var_pos_off[L7V_ENDPOINT] : aux(L7C_ENDPOINT);
  dfgpu_pos sp = reinterpret_cast<const dfgpu_pos*>(sp_base)[row];
  dfgpu_pos ep = reinterpret_cast<const dfgpu_pos*>(ep_base)[row];
  v.app_service = SRef{p.arena + dyn->var_buf_off[L7V_APP_SERVICE] + sp.start, (uint32_t)(sp.end - sp.start)};
  v.endpoint = SRef{p.arena + dyn->var_buf_off[L7V_ENDPOINT] + ep.start, (uint32_t)(ep.end - ep.start)};
}

// i-th (name, value) of attributeNames/attributeValues = h.Attribute* ++ the three
// extra fields (:660-663); n_names/n_values are the lengths BEFORE the append.
__device__ __forceinline__ SRef attr_name_at(const L7Dev* d, const RowView& v, uint32_t i) {
  uint32_t n = (uint32_t)(v.an_hi - v.an_lo);
  if (i < n) {
    dfgpu_pos q = v.an_pos[v.an_lo + i];
    return SRef{v.an_buf + q.start, (uint32_t)(q.end - q.start)};
  }
  return cs_ref(d, CS_FT_APP_SERVICE + (i - n));
}
__device__ __forceinline__ SRef attr_value_at(const RowView& v, uint32_t i) {
  uint32_t n = (uint32_t)(v.av_hi - v.av_lo);
  if (i < n) {
    dfgpu_pos q = v.av_pos[v.av_lo + i];
    return SRef{v.av_buf + q.start, (uint32_t)(q.end - q.start)};
  }
  uint32_t e = i - n;
  if (e == 0) return v.app_service;
  if (e == 1) return v.endpoint;
  return SRef{nullptr, 0};  // app_instance: always "" on this path
}

__device__ __forceinline__ U128 key_fp(const L7Dev* d, uint32_t kind, StrHash name, StrHash value, int32_t vpc,
                                       uint16_t pod_ns, uint8_t field_type, uint16_t team) {
  const uint64_t x = (uint64_t)(uint32_t)vpc << 32 | (uint64_t)pod_ns << 16 | (uint64_t)field_type << 8 | kind;
  const uint64_t y = (uint64_t)d->org_id << 16 | team;
  uint64_t lo = name.a ^ mix64(value.a + x * 0x9E3779B97F4A7C15ull) ^ (y * 0xD6E8FEB86659FD93ull);
  uint64_t hi = name.b + mix64(value.b ^ (x * 0xC2B2AE3D27D4EB4Full)) + (y * 0xA0761D6478BD642Full);
  return fp_finish(lo, hi);
}

__device__ __forceinline__ unsigned long long pos_key(const L7Dev* d, uint32_t row, uint32_t side, uint32_t metric,
                                                      uint32_t i) {
  unsigned long long pos = (unsigned long long)(row & 0x3ffffffu) << 22 | (unsigned long long)side << 21 |
                           (unsigned long long)metric << 20 | (i & 0xfffffu);
  return (unsigned long long)(d->epoch & 0xffffu) << 48 | (0xffffffffffffull - pos);
}

// The three phases share one walk over a row's keys.
//   PHASE 0: value / app-service keys: upsert + candidate atomicMin
//   PHASE 1: decide value winners (-> counts), set their time; field + metric keys upsert + atomicMin
//   PHASE 2: decide field winners (-> counts), set their time
//   PHASE 3: write rows
#ifndef FT_P0_MIN_BLOCKS
#define FT_P0_MIN_BLOCKS 4
#endif
template <int PHASE>
__global__ void __launch_bounds__(256, PHASE == 0 ? FT_P0_MIN_BLOCKS : 1) k_l7_flow_tag(const __grid_constant__ BatchParams p) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t warp = row >> 5, lane = threadIdx.x & 31;
  DynState* dyn = p.dyn;
  const L7Dev* d = reinterpret_cast<const L7Dev*>(p.ext);
  if ((dyn->stats.need_repair && !p.repair) || dyn->overflow) return;
  if (PHASE == 3 && dyn->stats.ft_overflow) return;
  // steady state: phase 0 flagged no row of the batch — the count / write walks have nothing to do (their per-warp
  // counts are then never read: k_l7_ft_layout takes the totals as zero)
  if (PHASE != 0 && !dyn->stats.ft_any) return;
  const uint32_t nrows = dyn->stats.nrows;
  const uint32_t nwarps = dyn->nwarps;
  const size_t wstride = p.warp_cap;
  // PHASE 0 notes per row whether ANY of its keys is new/expired; in steady state (a
  // saturated tag set) almost no row is, and phases 1-3 skip the row without hashing
  bool active = row < nrows;
  if (PHASE != 0 && active) active = d->row_flag[row] != 0;
  uint32_t c_fv = 0, b_fvn = 0, b_fvv = 0, c_f = 0, b_fn = 0, c_as = 0, b_as = 0;

  // PHASE 3 output cursors
  uint64_t o_fv = 0, o_fvn = 0, o_fvv = 0, o_f = 0, o_fn = 0, o_as = 0, o_asb = 0;
  if (PHASE == 3 && warp < nwarps) {
    // exclusive offsets = warp base + intra-warp exclusive scan of this thread's counts,
    // recomputed below after the counting walk; see the two-walk structure.
  }

  RowView v;
  const uint32_t timeout = d->ft_timeout;
  if (PHASE == 0) {
    // The steady-state walk (every batch, every row): each string is hashed ONCE and its key probed for
    // both sides; the first slot of both probes is loaded before either is consumed, and a key that is
    // present and fresh costs nothing beyond that one 32-byte sector.  The order of the probes is free in
    // this phase (insertion + atomicMax commute); phases 1-3 below keep the reference's side-major order.
    if (row < nrows) {
      load_row(p, row, v);
      const uint32_t ns = v.n_sides;
      const int32_t vpc0 = v.vpc[0], vpc1 = v.vpc[1];
      const uint16_t pns0 = v.pod_ns[0], pns1 = v.pod_ns[1];
      bool any = false;
      auto probe2 = [&](uint32_t kind, StrHash hn, StrHash hv, uint8_t ftype, uint32_t metric, uint32_t i) {
        const U128 f0 = key_fp(d, kind, hn, hv, vpc0, pns0, ftype, v.team);
        const U128 f1 = key_fp(d, kind, hn, hv, vpc1, pns1, ftype, v.team);
        TsPeek k0, k1;
        ts_peek(d->ft, d->ft_mask, f0, k0);
        if (ns > 1) ts_peek(d->ft, d->ft_mask, f1, k1);
        // a key the saturated set cannot hold (upsert == nullptr) is emitted by every row that carries
        // it: duplicates are idempotent in ClickHouse (flow_tag.go:153-160), dropping the batch is not
        if (!ts_peek_fresh(k0, f0, v.time, timeout)) {
          FtEntry* e = ft_upsert(d, f0);
          if (!e) any = true;
          else if (ft_candidate(e, v.time, timeout)) {
            atomicMax(&e->first_pos, pos_key(d, row, 0, metric, i));
            any = true;
          }
        }
        if (ns > 1 && !ts_peek_fresh(k1, f1, v.time, timeout)) {
          FtEntry* e = ft_upsert(d, f1);
          if (!e) any = true;
          else if (ft_candidate(e, v.time, timeout)) {
            atomicMax(&e->first_pos, pos_key(d, row, 1, metric, i));
            any = true;
          }
        }
      };
      const uint32_t n_names = (uint32_t)(v.an_hi - v.an_lo), n_values = (uint32_t)(v.av_hi - v.av_lo);
      const uint32_t tot_names = n_names + 3, tot_values = n_values + 3;
      const uint32_t min_len = tot_names > tot_values ? tot_values : tot_names;  // :666-674
      for (uint32_t i = 0; i < min_len; i++) {
        const SRef name = attr_name_at(d, v, i), value = attr_value_at(v, i);
        probe2(KIND_FIELD_VALUE, str_hash(name.p, name.n), str_hash(value.p, value.n), 0, 0, i);
      }
      const uint32_t n_metrics = (uint32_t)(v.mn_hi - v.mn_lo);  // :726-742
      for (uint32_t i = 0; i < n_metrics; i++) {
        const dfgpu_pos q = v.mn_pos[v.mn_lo + i];
        probe2(KIND_FIELD, str_hash(v.mn_buf + q.start, (uint32_t)(q.end - q.start)), StrHash{0, 0}, 1, 1, i);
      }
      if (v.app_service.n) {  // decoder.go:464-472
        const U128 fpa =
            key_fp(d, KIND_APP_SERVICE, str_hash(v.app_service.p, v.app_service.n), StrHash{0, 0}, 0, 0, 0, v.team);
        TsPeek k;
        ts_peek(d->ft, d->ft_mask, fpa, k);
        if (!ts_peek_fresh(k, fpa, v.time_as, timeout)) {
          FtEntry* e = ft_upsert(d, fpa);
          if (!e) any = true;
          else if (ft_candidate(e, v.time_as, timeout)) {
            atomicMax(&e->first_pos, pos_key(d, row, 0, 0, 0));
            any = true;
          }
        }
      }
      d->row_flag[row] = any ? 1 : 0;
      if (any) dyn->stats.ft_any = 1;
    }
    {  // newest lru time of the batch: the clock the rebuild measures expiry against
      uint32_t tmax = row < nrows ? (v.time > v.time_as ? v.time : v.time_as) : 0u;
      tmax = __reduce_max_sync(0xffffffffu, tmax);
      if (lane == 0 && tmax) atomicMax(&d->fts->now, tmax);
    }
    return;
  }
  constexpr uint32_t FT_HASH_CACHE = 8;
  StrHash hcache[2 * FT_HASH_CACHE];
  uint8_t* fa = d->ft_arena;
  const uint64_t* fo = dyn->stats.ft_off;

  // A PHASE-3 thread walks twice: first to count (for its offsets), then to write.
  for (int walk = 0; walk < (PHASE == 3 ? 2 : 1); walk++) {
    const bool writing = PHASE == 3 && walk == 1;
    if (PHASE == 3 && walk == 1) {
      uint32_t e_fv = warp_excl_scan(c_fv), e_fvn = warp_excl_scan(b_fvn), e_fvv = warp_excl_scan(b_fvv);
      uint32_t e_f = warp_excl_scan(c_f), e_fn = warp_excl_scan(b_fn), e_as = warp_excl_scan(c_as),
               e_asb = warp_excl_scan(b_as);
      const uint32_t wb = warp < nwarps ? warp : 0;
      o_fv = (uint64_t)p.warp_base[(size_t)FT_FV_ROWS * wstride + wb] + e_fv;
      o_fvn = (uint64_t)p.warp_base[(size_t)FT_FV_NAME_B * wstride + wb] + e_fvn;
      o_fvv = (uint64_t)p.warp_base[(size_t)FT_FV_VAL_B * wstride + wb] + e_fvv;
      o_f = (uint64_t)p.warp_base[(size_t)FT_F_ROWS * wstride + wb] + e_f;
      o_fn = (uint64_t)p.warp_base[(size_t)FT_F_NAME_B * wstride + wb] + e_fn;
      o_as = (uint64_t)p.warp_base[(size_t)FT_AS_ROWS * wstride + wb] + e_as;
      o_asb = (uint64_t)p.warp_base[(size_t)FT_AS_B * wstride + wb] + e_asb;
    }
    if (!active) continue;
    if (walk == 0) load_row(p, row, v);
    const uint32_t n_names = (uint32_t)(v.an_hi - v.an_lo), n_values = (uint32_t)(v.av_hi - v.av_lo);
    const uint32_t tot_names = n_names + 3, tot_values = n_values + 3;
    const uint32_t min_len = tot_names > tot_values ? tot_values : tot_names;  // :666-674
    for (uint32_t side = 0; side < v.n_sides; side++) {
      for (uint32_t i = 0; i < min_len; i++) {
        SRef name = attr_name_at(d, v, i), value = attr_value_at(v, i);
        // string hashes are side independent: computed at side 0, reused at side 1
        StrHash hn, hv;
        if (side == 0 || i >= FT_HASH_CACHE) {
          hn = str_hash(name.p, name.n);
          hv = str_hash(value.p, value.n);
          if (i < FT_HASH_CACHE) {
            hcache[2 * i] = hn;
            hcache[2 * i + 1] = hv;
          }
        } else {
          hn = hcache[2 * i];
          hv = hcache[2 * i + 1];
        }
        U128 fpv = key_fp(d, KIND_FIELD_VALUE, hn, hv, v.vpc[side], v.pod_ns[side], 0, v.team);
        unsigned long long pk = pos_key(d, row, side, 0, i);
        FtEntry* e = ft_find(d, fpv);
        if (e && e->first_pos != pk) continue;  // not the emitting occurrence (no entry: saturated set, emit)
        if (PHASE == 1 && !e) atomicAdd(&d->fts->unrecorded, 1u);
        if (PHASE == 1 && e) {
          e->time = v.time;   // lru Add(key, time)
          e->state = 1;
        }
        if (PHASE == 1 || (PHASE == 3 && walk == 0)) {
          c_fv++;
          b_fvn += name.n;
          b_fvv += value.n;
        }
        if (writing) {
          reinterpret_cast<uint32_t*>(fa + fo[FO_FV_TIME])[o_fv] = v.time;
          reinterpret_cast<int32_t*>(fa + fo[FO_FV_VPC])[o_fv] = v.vpc[side];
          reinterpret_cast<uint16_t*>(fa + fo[FO_FV_PODNS])[o_fv] = v.pod_ns[side];
          (fa + fo[FO_FV_FTYPE])[o_fv] = 0;   // "tag"
          (fa + fo[FO_FV_FVTYPE])[o_fv] = 0;  // "string" (:82-88: FieldType == tag)
          reinterpret_cast<uint16_t*>(fa + fo[FO_FV_TEAM])[o_fv] = v.team;
          (fa + fo[FO_FV_TABLE])[o_fv] = 0;
          reinterpret_cast<dfgpu_pos*>(fa + fo[FO_FV_NAME_POS])[o_fv] = dfgpu_pos{(int64_t)o_fvn, (int64_t)(o_fvn + name.n)};
          reinterpret_cast<dfgpu_pos*>(fa + fo[FO_FV_VAL_POS])[o_fv] = dfgpu_pos{(int64_t)o_fvv, (int64_t)(o_fvv + value.n)};
          reinterpret_cast<uint64_t*>(fa + fo[FO_FV_COUNT])[o_fv] = 1;
          reinterpret_cast<U128*>(fa + fo[FO_FV_FP])[o_fv] = fpv;
          for (uint32_t c = 0; c < name.n; c++) (fa + fo[FO_FV_NAME_BUF])[o_fvn + c] = name.p[c];
          for (uint32_t c = 0; c < value.n; c++) (fa + fo[FO_FV_VAL_BUF])[o_fvv + c] = value.p[c];
          o_fv++;
          o_fvn += name.n;
          o_fvv += value.n;
        }
        // the tag-only key is only consulted when the value row was emitted (:703-722)
        if (i >= n_names) continue;
        U128 fpf = key_fp(d, KIND_FIELD, hn, StrHash{0, 0}, v.vpc[side], v.pod_ns[side], 0, v.team);
        if (PHASE == 1) {
          FtEntry* ef = ft_upsert(d, fpf);
          if (ef && ft_candidate(ef, v.time, timeout)) atomicMax(&ef->first_pos, pk);
          continue;
        }
        FtEntry* ef = ft_find(d, fpf);
        if (ef && ef->first_pos != pk) continue;
        if (PHASE == 2 && !ef) atomicAdd(&d->fts->unrecorded, 1u);
        if (PHASE == 2 && ef) {
          ef->time = v.time;
          ef->state = 1;
        }
        if (PHASE == 2 || (PHASE == 3 && walk == 0)) {
          c_f++;
          b_fn += name.n;
        }
        if (writing) {
          reinterpret_cast<uint32_t*>(fa + fo[FO_F_TIME])[o_f] = v.time;
          reinterpret_cast<int32_t*>(fa + fo[FO_F_VPC])[o_f] = v.vpc[side];
          reinterpret_cast<uint16_t*>(fa + fo[FO_F_PODNS])[o_f] = v.pod_ns[side];
          (fa + fo[FO_F_FTYPE])[o_f] = 0;
          (fa + fo[FO_F_FVTYPE])[o_f] = 0;
          reinterpret_cast<uint16_t*>(fa + fo[FO_F_TEAM])[o_f] = v.team;
          (fa + fo[FO_F_TABLE])[o_f] = 0;
          reinterpret_cast<U128*>(fa + fo[FO_F_FP])[o_f] = fpf;
          reinterpret_cast<dfgpu_pos*>(fa + fo[FO_F_NAME_POS])[o_f] = dfgpu_pos{(int64_t)o_fn, (int64_t)(o_fn + name.n)};
          for (uint32_t c = 0; c < name.n; c++) (fa + fo[FO_F_NAME_BUF])[o_fn + c] = name.p[c];
          o_f++;
          o_fn += name.n;
        }
      }
      // metrics names: FieldType = metrics, value "" (:726-742)
      const uint32_t n_metrics = (uint32_t)(v.mn_hi - v.mn_lo);
      for (uint32_t i = 0; i < n_metrics; i++) {
        dfgpu_pos q = v.mn_pos[v.mn_lo + i];
        SRef name{v.mn_buf + q.start, (uint32_t)(q.end - q.start)};
        U128 fpf = key_fp(d, KIND_FIELD, str_hash(name.p, name.n), StrHash{0, 0}, v.vpc[side], v.pod_ns[side], 1, v.team);
        unsigned long long pk = pos_key(d, row, side, 1, i);
        if (PHASE == 1) continue;
        FtEntry* ef = ft_find(d, fpf);
        if (ef && ef->first_pos != pk) continue;
        if (PHASE == 2 && !ef) atomicAdd(&d->fts->unrecorded, 1u);
        if (PHASE == 2 && ef) {
          ef->time = v.time;
          ef->state = 1;
        }
        if (PHASE == 2 || (PHASE == 3 && walk == 0)) {
          c_f++;
          b_fn += name.n;
        }
        if (writing) {
          reinterpret_cast<uint32_t*>(fa + fo[FO_F_TIME])[o_f] = v.time;
          reinterpret_cast<int32_t*>(fa + fo[FO_F_VPC])[o_f] = v.vpc[side];
          reinterpret_cast<uint16_t*>(fa + fo[FO_F_PODNS])[o_f] = v.pod_ns[side];
          (fa + fo[FO_F_FTYPE])[o_f] = 1;   // "metrics"
          (fa + fo[FO_F_FVTYPE])[o_f] = 1;  // "float": empty value and FieldType != tag
          reinterpret_cast<uint16_t*>(fa + fo[FO_F_TEAM])[o_f] = v.team;
          (fa + fo[FO_F_TABLE])[o_f] = 0;
          reinterpret_cast<U128*>(fa + fo[FO_F_FP])[o_f] = fpf;
          reinterpret_cast<dfgpu_pos*>(fa + fo[FO_F_NAME_POS])[o_f] = dfgpu_pos{(int64_t)o_fn, (int64_t)(o_fn + name.n)};
          for (uint32_t c = 0; c < name.n; c++) (fa + fo[FO_F_NAME_BUF])[o_fn + c] = name.p[c];
          o_f++;
          o_fn += name.n;
        }
      }
    }
    // app_service tag (decoder.go:464-472): AppInstance is always "" here
    if (v.app_service.n) {
      U128 fpa = key_fp(d, KIND_APP_SERVICE, str_hash(v.app_service.p, v.app_service.n), StrHash{0, 0}, 0, 0, 0, v.team);
      unsigned long long pk = pos_key(d, row, 0, 0, 0);
      {
        FtEntry* e = ft_find(d, fpa);
        if (!e || e->first_pos == pk) {
          if (PHASE == 1 && !e) atomicAdd(&d->fts->unrecorded, 1u);
          if (PHASE == 1 && e) {
            e->time = v.time_as;
            e->state = 1;
          }
          if (PHASE == 1 || (PHASE == 3 && walk == 0)) {
            c_as++;
            b_as += v.app_service.n;
          }
          if (writing) {
            reinterpret_cast<uint32_t*>(fa + fo[FO_AS_TIME])[o_as] = v.time_as;
            (fa + fo[FO_AS_TABLE])[o_as] = 0;
            reinterpret_cast<U128*>(fa + fo[FO_AS_FP])[o_as] = fpa;
            reinterpret_cast<dfgpu_pos*>(fa + fo[FO_AS_SVC_POS])[o_as] =
                dfgpu_pos{(int64_t)o_asb, (int64_t)(o_asb + v.app_service.n)};
            reinterpret_cast<dfgpu_pos*>(fa + fo[FO_AS_INST_POS])[o_as] = dfgpu_pos{0, 0};
            reinterpret_cast<uint16_t*>(fa + fo[FO_AS_TEAM])[o_as] = v.team;
            for (uint32_t c = 0; c < v.app_service.n; c++) (fa + fo[FO_AS_BUF])[o_asb + c] = v.app_service.p[c];
            o_as++;
            o_asb += v.app_service.n;
          }
        }
      }
    }
  }
  if (PHASE == 1 || PHASE == 2) {
    // per-warp totals of the rows/bytes decided in this phase
    uint32_t vals[FT_SLOTS] = {c_fv, b_fvn, b_fvv, c_f, b_fn, c_as, b_as};
    const int lo = PHASE == 1 ? 0 : FT_F_ROWS, hi = PHASE == 1 ? FT_F_ROWS : FT_AS_ROWS;
#pragma unroll
    for (int s = 0; s < FT_SLOTS; s++) {
      bool mine = (s >= lo && s < hi) || (PHASE == 1 && s >= FT_AS_ROWS);
      if (!mine) continue;
      uint32_t t = warp_sum(vals[s]);
      if (lane == 0 && warp < nwarps) p.warp_tot[(size_t)s * wstride + warp] = t;
    }
  }
}

__global__ void k_l7_ft_clear_pos(FtEntry* t, uint32_t mask) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= mask; i += gridDim.x * blockDim.x) t[i].first_pos = 0;
}

// the tag set's bookkeeping, for the host (fetch reads DynState)
__global__ void k_l7_ft_stats(const BatchParams p) {
  const L7Dev* d = reinterpret_cast<const L7Dev*>(p.ext);
  p.dyn->stats.ft_count = d->fts->count;
  p.dyn->stats.ft_unrecorded = d->fts->unrecorded;
  p.dyn->stats.ft_evicted = d->fts->evicted;
}

// placement of the three flow_tag tables inside the ft arena
__global__ void k_l7_ft_layout(const BatchParams p) {
  DynState* dyn = p.dyn;
  const L7Dev* d = reinterpret_cast<const L7Dev*>(p.ext);
  if ((dyn->stats.need_repair && !p.repair) || dyn->overflow) return;
  uint64_t* tot = dyn->stats.ft_tot;
  if (!dyn->stats.ft_any)  // the counting walks did not run: what the scan summed is stale
    for (int i = 0; i < 8; i++) tot[i] = 0;
  uint64_t* fo = dyn->stats.ft_off;
  uint64_t cur = 0;
  auto take = [&](int id, uint64_t bytes) {
    fo[id] = cur;
    cur += (bytes + 255) & ~255ull;
  };
  uint64_t nfv = tot[FT_FV_ROWS], nf = tot[FT_F_ROWS], nas = tot[FT_AS_ROWS];
  take(FO_FV_TIME, nfv * 4); take(FO_FV_VPC, nfv * 4); take(FO_FV_PODNS, nfv * 2); take(FO_FV_FTYPE, nfv);
  take(FO_FV_FVTYPE, nfv); take(FO_FV_TEAM, nfv * 2); take(FO_FV_TABLE, nfv); take(FO_FV_NAME_POS, nfv * 16);
  take(FO_FV_VAL_POS, nfv * 16); take(FO_FV_COUNT, nfv * 8); take(FO_FV_NAME_BUF, tot[FT_FV_NAME_B]);
  take(FO_FV_VAL_BUF, tot[FT_FV_VAL_B]);
  take(FO_F_TIME, nf * 4); take(FO_F_VPC, nf * 4); take(FO_F_PODNS, nf * 2); take(FO_F_FTYPE, nf);
  take(FO_F_FVTYPE, nf); take(FO_F_TEAM, nf * 2); take(FO_F_TABLE, nf); take(FO_F_NAME_POS, nf * 16);
  take(FO_F_NAME_BUF, tot[FT_F_NAME_B]);
  take(FO_AS_TIME, nas * 4); take(FO_AS_TABLE, nas); take(FO_AS_SVC_POS, nas * 16); take(FO_AS_INST_POS, nas * 16);
  take(FO_AS_TEAM, nas * 2); take(FO_AS_BUF, tot[FT_AS_B]);
  take(FO_FV_FP, nfv * 16); take(FO_F_FP, nf * 16); take(FO_AS_FP, nas * 16);
  dyn->stats.ft_end = cur;
  dyn->stats.ft_overflow = cur > d->ft_arena_bytes ? 1 : 0;
}

}  // namespace

static int l7_ensure_state(dfgpu_ctx* ctx) {
  if (ctx->ft) return 0;
  L7Host* h = new L7Host();
  // ---- constant strings + lookup tables (l7_tables_host.h) ----
  std::vector<std::string> cs;
  L7Dev& d = h->dev;
  memset(&d, 0, sizeof d);
  l7_build_tables(cs, d.tb);
  {
    std::vector<uint32_t> t = l7_tok_table();
    memcpy(d.tok_tbl, t.data(), sizeof d.tok_tbl);
  }
  std::vector<uint8_t> blob;
  std::vector<uint32_t> off;
  l7_layout_strings(cs, blob, off);
  CU(cudaMalloc(&h->d_blob, blob.size() + 16));
  CU(cudaMalloc(&h->d_off, off.size() * 4));
  CU(cudaMemcpy(h->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(h->d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
  d.tb.cs_blob = h->d_blob;
  d.tb.cs_off = h->d_off;
  // ---- device tag set: the three LRUs of the reference (field_value, field, app_service) in one table,
  // 2x their combined capacity; rebuilt (expiry + oldest-first eviction) when more than half full ----
  if (tagset_create(h->set, (uint64_t)ctx->cfg.flow_tag_cache_max_size * 3 * 2, ctx->cfg.flow_tag_cache_flush_timeout))
    return fail(ctx, DFGPU_ENOMEM, "flow_tag set allocation failed");
  d.ft_timeout = ctx->cfg.flow_tag_cache_flush_timeout;
  h->ft_arena_cap = 16u << 20;
  CU(cudaMalloc(&h->d_ft_arena, h->ft_arena_cap));
  d.ft_arena = h->d_ft_arena;
  d.ft_arena_bytes = h->ft_arena_cap;
  CU(cudaMalloc(&h->d_dev, sizeof(L7Dev)));
  // persistent grids: one CTA (or a few) per SM, sized once
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  h->grid_decode = (unsigned)sms * 5;
  h->grid_strings = (unsigned)sms * 6;
  h->grid_kg = (unsigned)sms * 4;
  CU(cudaMalloc(&h->d_raw_scratch, (size_t)h->grid_decode * l7_decode_warps_per_cta() * R7_COUNT * 32 * sizeof(uint64_t)));
  d.raw_scratch = h->d_raw_scratch;
  ctx->ft = h;
  return 0;
}

// per-record buffers, carved from one allocation (first L7 batch of the context, or after the record
// index grew)
static int l7_ensure_rec_buffers(dfgpu_ctx* ctx, L7Host* h, size_t cap) {
  if (cap <= h->rec_cap) return 0;
  if (h->d_rec_buf) CU(cudaFree(h->d_rec_buf));
  h->d_rec_buf = nullptr;
  const size_t per_rec = (size_t)L7REF_COUNT * 8 + (size_t)L7E_COUNT * 8 + 8 + (size_t)L7ARR_COUNT * 4 + 4 + (size_t)L7KG_WORDS * 4 + 4 + 1;
  CU(cudaMalloc(&h->d_rec_buf, per_rec * cap + 4096));
  uint8_t* q = h->d_rec_buf;
  L7Dev& d = h->dev;
  d.refs = reinterpret_cast<uint64_t*>(q); q += (size_t)L7REF_COUNT * 8 * cap;
  d.elems = reinterpret_cast<uint64_t*>(q); q += (size_t)L7E_COUNT * 8 * cap;
  d.thash = reinterpret_cast<uint64_t*>(q); q += 8 * cap;
  d.arr = reinterpret_cast<uint32_t*>(q); q += (size_t)L7ARR_COUNT * 4 * cap;
  d.meta = reinterpret_cast<uint32_t*>(q); q += 4 * cap;
  d.kgin = reinterpret_cast<uint32_t*>(q); q += (size_t)L7KG_WORDS * 4 * cap;
  d.rec_row = reinterpret_cast<uint32_t*>(q); q += 4 * cap;
  d.row_flag = q;
  d.rec_cap = (uint32_t)cap;
  h->rec_cap = cap;
  return 0;
}

// called by dfgpu_run before the passes: (re)publish the device-side L7 state
int l7_prepare(dfgpu_ctx* ctx, cudaStream_t s) {
  int rc = l7_ensure_state(ctx);
  if (rc) return rc;
  L7Host* h = reinterpret_cast<L7Host*>(ctx->ft);
  rc = l7_ensure_rec_buffers(ctx, h, ctx->bp.rec_cap);
  if (rc) return rc;
  h->epoch++;
  if ((h->epoch & 0xffff) == 0) {
    // the 16-bit epoch inside first_pos wrapped: forget every stale candidate
    k_l7_ft_clear_pos<<<1024, 256, 0, s>>>(h->set.cur, h->set.mask);
    h->epoch = 1;
  }
  tagset_maybe_rebuild(h->set, s, &ctx->launches);  // between batches, asynchronous
  h->dev.ft = h->set.cur;
  h->dev.ft_mask = h->set.mask;
  h->dev.fts = h->set.d_state;
  h->dev.epoch = h->epoch;
  h->dev.org_id = ctx->cur_org;
  h->dev.ft_arena = h->d_ft_arena;
  h->dev.ft_arena_bytes = h->ft_arena_cap;
  CU(cudaMemcpyAsync(h->d_dev, &h->dev, sizeof(L7Dev), cudaMemcpyHostToDevice, s));
  ctx->bp.ext = h->d_dev;
  return 0;
}

void l7_grids(dfgpu_ctx* ctx, unsigned* decode, unsigned* strings, unsigned* kg) {
  L7Host* h = reinterpret_cast<L7Host*>(ctx->ft);
  *decode = h->grid_decode;
  *strings = h->grid_strings;
  *kg = h->grid_kg;
}

static void launch_ft_writer(dfgpu_ctx* ctx, cudaStream_t s) {
  BatchParams& bp = ctx->bp;
  k_l7_flow_tag<3><<<rec_grid(bp), 256, 0, s>>>(bp);
  ctx->launches++;
}

int l7_flow_tags(dfgpu_ctx* ctx, cudaStream_t s) {
  BatchParams& bp = ctx->bp;
  {
    dim3 grid(rec_grid(bp)), block(256);
    k_l7_flow_tag<0><<<grid, block, 0, s>>>(bp);
    k_l7_flow_tag<1><<<grid, block, 0, s>>>(bp);
    k_l7_flow_tag<2><<<grid, block, 0, s>>>(bp);
    launch_slot_scan_to(bp, FT_SLOTS, ctx->d_dyn->stats.ft_tot, s);
    ctx->launches += 4;
  }
  k_l7_ft_layout<<<1, 1, 0, s>>>(bp);
  ctx->launches++;
  launch_ft_writer(ctx, s);
  k_l7_ft_stats<<<1, 1, 0, s>>>(bp);
  launch_commit(bp, 1, s);
  ctx->launches += 2;
  CU(cudaGetLastError());
  return 0;
}

static void ft_lc_const(dfgpu_batch* b, dfgpu_col& col, const std::vector<std::string>& dict) {
  b->owned.emplace_back();
  b->owned_pos.emplace_back();
  auto& buf = b->owned.back();
  auto& pos = b->owned_pos.back();
  for (auto& s : dict) {
    pos.push_back(dfgpu_pos{(int64_t)buf.size(), (int64_t)(buf.size() + s.size())});
    buf.insert(buf.end(), s.begin(), s.end());
  }
  col.dict_data = buf.data();
  col.dict_pos = pos.data();
  col.dict_n = pos.size();
}

// LowCardinality dictionaries of the flow_tag tables hold at most two entries per
// column; keys are written as raw codes and remapped here (rows are few).
static void ft_lc_coded(dfgpu_batch* b, dfgpu_col& col, uint8_t* keys, uint64_t n, const char* name0,
                        const char* name1) {
  int first[2] = {-1, -1};
  for (uint64_t i = 0; i < n && (first[0] < 0 || first[1] < 0); i++)
    if (first[keys[i]] < 0) first[keys[i]] = (int)i;
  std::vector<std::string> dict;
  uint8_t remap[2] = {0, 0};
  if (first[0] >= 0 && (first[1] < 0 || first[0] < first[1])) {
    dict.push_back(name0);
    if (first[1] >= 0) { dict.push_back(name1); remap[1] = 1; }
  } else if (first[1] >= 0) {
    dict.push_back(name1);
    remap[1] = 0;
    if (first[0] >= 0) { dict.push_back(name0); remap[0] = 1; }
  }
  if (remap[0] != 0 || remap[1] != 1)
    for (uint64_t i = 0; i < n; i++) keys[i] = remap[keys[i]];
  ft_lc_const(b, col, dict);
}

int l7_fetch_tables(dfgpu_ctx* ctx, dfgpu_batch* b) {
  L7Host* h = reinterpret_cast<L7Host*>(ctx->ft);
  cudaStream_t s = ctx->stream;
  DynState* d = ctx->h_dyn;
  h->set.count_seen = d->stats.ft_count;
  b->counters.flow_tag_unrecorded = d->stats.ft_unrecorded;
  b->counters.flow_tag_evicted = d->stats.ft_evicted;
  if (d->stats.ft_overflow) {
    // grow the ft arena and re-run only the writer: the emit decisions are already
    // recorded in the tag set and do not change
    size_t want = (size_t)d->stats.ft_end + (1u << 20);
    CU(cudaFree(h->d_ft_arena));
    h->d_ft_arena = nullptr;
    CU(cudaMalloc(&h->d_ft_arena, want));
    h->ft_arena_cap = want;
    h->dev.ft_arena = h->d_ft_arena;
    h->dev.ft_arena_bytes = want;
    CU(cudaMemcpyAsync(h->d_dev, &h->dev, sizeof(L7Dev), cudaMemcpyHostToDevice, s));
    uint32_t zero = 0;
    CU(cudaMemcpyAsync(&ctx->d_dyn->stats.ft_overflow, &zero, 4, cudaMemcpyHostToDevice, s));
    launch_ft_writer(ctx, s);
    CU(cudaStreamSynchronize(s));
  }
  uint64_t bytes = d->stats.ft_end;
  b->owned.emplace_back(bytes + 16);
  uint8_t* hb = b->owned.back().data();
  if (bytes) {
    CU(cudaMemcpyAsync(hb, h->d_ft_arena, bytes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  const uint64_t* fo = d->stats.ft_off;
  uint64_t nfv = d->stats.ft_tot[FT_FV_ROWS], nf = d->stats.ft_tot[FT_F_ROWS], nas = d->stats.ft_tot[FT_AS_ROWS];
  auto fixed = [&](dfgpu_batch::Table& t, const char* name, int ch, int es, uint64_t off, uint64_t n) {
    dfgpu_col c;
    memset(&c, 0, sizeof c);
    c.name = name;
    c.ch_type = ch;
    c.elem_size = es;
    c.nrows = n;
    c.data = hb + off;
    c.data_bytes = n * es;
    t.cols.push_back(c);
    return &t.cols.back();
  };
  auto strcol = [&](dfgpu_batch::Table& t, const char* name, int ch, uint64_t pos_off, uint64_t buf_off,
                    uint64_t buf_bytes, uint64_t n) {
    dfgpu_col c;
    memset(&c, 0, sizeof c);
    c.name = name;
    c.ch_type = ch;
    c.nrows = n;
    c.pos = (const dfgpu_pos*)(hb + pos_off);
    c.data = hb + buf_off;
    c.data_bytes = buf_bytes;
    t.cols.push_back(c);
  };
  for (int which = 0; which < 2; which++) {  // 0: custom_field, 1: custom_field_value
    b->tables.emplace_back();
    dfgpu_batch::Table& t = b->tables.back();
    t.cols.reserve(12);
    bool val = which == 1;
    uint64_t n = val ? nfv : nf;
    t.id = val ? DFGPU_TABLE_L7_CUSTOM_FIELD_VALUE : DFGPU_TABLE_L7_CUSTOM_FIELD;
    t.nrows = n;
    fixed(t, "time", DFGPU_CH_DATETIME, 4, fo[val ? FO_FV_TIME : FO_F_TIME], n);
    ft_lc_const(b, *fixed(t, "table", DFGPU_CH_LC_STR, 1, fo[val ? FO_FV_TABLE : FO_F_TABLE], n),
                n ? std::vector<std::string>{"l7_flow_log"} : std::vector<std::string>{});
    fixed(t, "vpc_id", DFGPU_CH_INT32, 4, fo[val ? FO_FV_VPC : FO_F_VPC], n);
    fixed(t, "pod_ns_id", DFGPU_CH_UINT16, 2, fo[val ? FO_FV_PODNS : FO_F_PODNS], n);
    ft_lc_coded(b, *fixed(t, "field_type", DFGPU_CH_LC_STR, 1, fo[val ? FO_FV_FTYPE : FO_F_FTYPE], n),
                hb + fo[val ? FO_FV_FTYPE : FO_F_FTYPE], n, "tag", "metrics");
    strcol(t, "field_name", DFGPU_CH_LC_STR_ROWS, fo[val ? FO_FV_NAME_POS : FO_F_NAME_POS],
           fo[val ? FO_FV_NAME_BUF : FO_F_NAME_BUF], d->stats.ft_tot[val ? FT_FV_NAME_B : FT_F_NAME_B], n);
    ft_lc_coded(b, *fixed(t, "field_value_type", DFGPU_CH_LC_STR, 1, fo[val ? FO_FV_FVTYPE : FO_F_FVTYPE], n),
                hb + fo[val ? FO_FV_FVTYPE : FO_F_FVTYPE], n, "string", "float");
    fixed(t, "team_id", DFGPU_CH_UINT16, 2, fo[val ? FO_FV_TEAM : FO_F_TEAM], n);
    if (val) {
      strcol(t, "field_value", DFGPU_CH_STR, fo[FO_FV_VAL_POS], fo[FO_FV_VAL_BUF], d->stats.ft_tot[FT_FV_VAL_B], n);
      fixed(t, "count", DFGPU_CH_UINT64, 8, fo[FO_FV_COUNT], n);
    }
  }
  {
    b->tables.emplace_back();
    dfgpu_batch::Table& t = b->tables.back();
    t.cols.reserve(8);
    t.id = DFGPU_TABLE_L7_APP_SERVICE;
    t.nrows = nas;
    fixed(t, "time", DFGPU_CH_DATETIME, 4, fo[FO_AS_TIME], nas);
    ft_lc_const(b, *fixed(t, "table", DFGPU_CH_LC_STR, 1, fo[FO_AS_TABLE], nas),
                nas ? std::vector<std::string>{"l7_flow_log"} : std::vector<std::string>{});
    strcol(t, "app_service", DFGPU_CH_LC_STR_ROWS, fo[FO_AS_SVC_POS], fo[FO_AS_BUF], d->stats.ft_tot[FT_AS_B], nas);
    strcol(t, "app_instance", DFGPU_CH_LC_STR_ROWS, fo[FO_AS_INST_POS], fo[FO_AS_BUF], 0, nas);
    fixed(t, "team_id", DFGPU_CH_UINT16, 2, fo[FO_AS_TEAM], nas);
  }
  b->tag_fps.push_back({DFGPU_TABLE_L7_CUSTOM_FIELD, (const uint64_t*)(hb + fo[FO_F_FP]), nf});
  b->tag_fps.push_back({DFGPU_TABLE_L7_CUSTOM_FIELD_VALUE, (const uint64_t*)(hb + fo[FO_FV_FP]), nfv});
  b->tag_fps.push_back({DFGPU_TABLE_L7_APP_SERVICE, (const uint64_t*)(hb + fo[FO_AS_FP]), nas});
  b->counters.new_field_count = (int64_t)nf;
  b->counters.new_field_value_count = (int64_t)nfv;
  b->counters.rows[DFGPU_TABLE_L7_CUSTOM_FIELD] = (int64_t)nf;
  b->counters.rows[DFGPU_TABLE_L7_CUSTOM_FIELD_VALUE] = (int64_t)nfv;
  b->counters.rows[DFGPU_TABLE_L7_APP_SERVICE] = (int64_t)nas;
  return 0;
}

// ---- multi-GPU tag merge: keys another rank emitted become "seen" here too -----------------------
__global__ void k_l7_ft_import(FtEntry* tbl, uint32_t mask, const U128* fps, const uint32_t* times, uint32_t n,
                               TagSetState* st) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  FtEntry* e = ts_upsert(tbl, mask, fps[i], st);
  if (!e) return;  // saturated: this rank may emit the key once more (idempotent)
  // the same key may arrive from several ranks: keep the newest emission time
  atomicMax(&e->time, times[i]);
  e->state = 1;
}

int l7_tags_import(dfgpu_ctx* ctx, const uint64_t* fps, const uint32_t* times, uint64_t n) {
  if (!n) return 0;
  int rc = l7_ensure_state(ctx);
  if (rc) return rc;
  L7Host* h = reinterpret_cast<L7Host*>(ctx->ft);
  cudaStream_t s = ctx->stream;
  // the staging buffer of the merge: grown only when a step brings more keys than any step before
  if (n > h->import_cap) {
    if (h->d_import) CU(cudaFree(h->d_import));
    h->d_import = nullptr;
    const size_t cap = (size_t)n + n / 2 + 4096;
    CU(cudaMalloc(&h->d_import, cap * 20 + 16));
    h->import_cap = cap;
  }
  U128* d_fps = reinterpret_cast<U128*>(h->d_import);
  uint32_t* d_times = reinterpret_cast<uint32_t*>(d_fps + h->import_cap);
  // host or device pointers (unified addressing): a merge that all-gathers over NCCL hands device buffers in
  CU(cudaMemcpyAsync(d_fps, fps, n * 16, cudaMemcpyDefault, s));
  CU(cudaMemcpyAsync(d_times, times, n * 4, cudaMemcpyDefault, s));
  k_l7_ft_import<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->set.cur, h->set.mask, d_fps, d_times, (uint32_t)n, h->set.d_state);
  ctx->launches++;
  CU(cudaStreamSynchronize(s));  // the caller's arrays may go away
  h->set.count_seen += (uint32_t)n;  // upper bound until the next fetch
  return 0;
}

// forget every flow_tag key (what restarting the decoder does to its lru caches)
int l7_tags_reset(dfgpu_ctx* ctx) {
  int rc = l7_ensure_state(ctx);
  if (rc) return rc;
  L7Host* h = reinterpret_cast<L7Host*>(ctx->ft);
  CU(cudaMemsetAsync(h->set.cur, 0, (size_t)h->set.slots() * sizeof(FtEntry), ctx->stream));
  CU(cudaMemsetAsync(h->set.d_state, 0, sizeof(TagSetState), ctx->stream));
  h->set.count_seen = 0;
  return 0;
}

}  // namespace dfgpu

extern "C" void dfgpu_l7_free_state(dfgpu_ctx* ctx) {
  using namespace dfgpu;
  if (!ctx || !ctx->ft) return;
  L7Host* h = reinterpret_cast<L7Host*>(ctx->ft);
  cudaFree(h->d_blob);
  cudaFree(h->d_off);
  tagset_destroy(h->set);
  cudaFree(h->d_ft_arena);
  cudaFree(h->d_rec_buf);
  cudaFree(h->d_raw_scratch);
  cudaFree(h->d_import);
  cudaFree(h->d_dev);
  delete h;
  ctx->ft = nullptr;
}
