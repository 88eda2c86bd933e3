/*
 * l4.cu — the L4 flow-log kernels: TaggedFlow record -> one l4_flow_log row.
 *
 * Same shape as the L7 path (l7_pass.cu): one thread per record, persistent warps striding over the batch, the
 * record decoded ONCE by the table-driven tokenizer (pb_tok.cuh + l4_tok.h) into shared-memory words, the row
 * built by every lane of the warp together (l4_row.cuh) with every column leaving as one coalesced, full-sector
 * store.  (Round 1 decoded twice with one switch per message into a 480-byte per-thread struct: a 1 KB local
 * frame + 1.6 KB of spills per thread, 23.8 ms per 625 k records.)  Fuses what the reference does in three
 * goroutine hops:
 *   decode            gogo Unmarshal via SimpleDecoder.ReadPB (simple_codec.go:229-248)
 *   row build         TaggedFlowToL4FlowLog  flow_log/log_data/l4_flow_log.go:955-971 (l4_row.cuh)
 *   geo               geo.QueryProvince flow_log/geo/geo.go:29 (k_l4_geo)
 *   SmartEncoding     KnowledgeGraph.FillL4 :730-743 -> kg_kernel.cuh (k_l4_kg)
 *   column pack       L4FlowLog.AppendToColumnBlock l4_column_block.go:848-859
 */
#include "ctx.h"
#include "kg_kernel.cuh"
#include "l4_decode.cuh"
#include "l4_row.cuh"

namespace dfgpu {

static_assert(L4C_REGION_ID_1 == L4C_REGION_ID_0 + 1 && L4C_AZ_ID_0 == L4C_REGION_ID_0 + 2 && L4C_HOST_ID_0 == L4C_REGION_ID_0 + 4 &&
                  L4C_L3_DEVICE_TYPE_0 == L4C_REGION_ID_0 + 6 && L4C_L3_DEVICE_ID_0 == L4C_REGION_ID_0 + 8 &&
                  L4C_POD_NODE_ID_0 == L4C_REGION_ID_0 + 10 && L4C_POD_NS_ID_0 == L4C_REGION_ID_0 + 12 &&
                  L4C_POD_GROUP_ID_0 == L4C_REGION_ID_0 + 14 && L4C_POD_ID_0 == L4C_REGION_ID_0 + 16 &&
                  L4C_POD_CLUSTER_ID_0 == L4C_REGION_ID_0 + 18 && L4C_L3_EPC_ID_0 == L4C_REGION_ID_0 + 20 &&
                  L4C_EPC_ID_0 == L4C_REGION_ID_0 + 22 && L4C_SUBNET_ID_0 == L4C_REGION_ID_0 + 24 &&
                  L4C_SERVICE_ID_0 == L4C_REGION_ID_0 + 26 && L4C_AUTO_INSTANCE_ID_0 == L4C_REGION_ID_0 + 28 &&
                  L4C_AUTO_INSTANCE_TYPE_0 == L4C_REGION_ID_0 + 29 && L4C_AUTO_SERVICE_ID_0 == L4C_REGION_ID_0 + 30 &&
                  L4C_AUTO_SERVICE_TYPE_0 == L4C_REGION_ID_0 + 31 && L4C_AUTO_INSTANCE_ID_1 == L4C_REGION_ID_0 + 32 &&
                  L4C_TAG_SOURCE_0 == L4C_REGION_ID_0 + 36 && L4C_TEAM_ID == L4C_REGION_ID_0 + 38,
              "KG_COLUMNS order (kg_kernel.cuh)");

// per-record hand-over of the decode kernel (stride rec_cap)
enum : int { L4ARR_ACL_N = 0, L4ARR_AGG_N, L4ARR_AGG_CHARS, L4ARR_COUNT };
constexpr uint32_t L4META_OK = 1u;  // bits 8..15 tap-side code, 16..23 / 24..31 province codes (k_l4_geo)

struct L4Dev {
  uint32_t tok_tbl[M4_COUNT * L4_MAXF];
  uint64_t* elems;   // [rec_cap][L4E_COUNT] captured acl_gids / aggregated_flow_ids values
  uint32_t* kgin;    // [KG_WORDS][rec_cap]
  uint64_t* ip6;     // [2][rec_cap] offset | length << 32
  uint64_t* dom;     // [rec_cap] request_domain reference
  uint32_t* arr;     // [L4ARR_COUNT][rec_cap]
  uint32_t* meta;    // [rec_cap]
  uint32_t* rec_row; // [rec_cap]
  uint32_t rec_cap;
};

struct L4Host {
  L4Dev dev{};
  L4Dev* d_dev = nullptr;
  uint8_t* d_rec_buf = nullptr;
  size_t rec_cap = 0;
  unsigned grid_decode = 0, grid_pack = 0, grid_kg = 0;
};

namespace {

constexpr int D_THREADS = 64;  // 131 words per thread: 33.5 KB per CTA, 6 CTAs per SM

template <class T>
__device__ __forceinline__ void st(const BatchParams& p, int col, uint32_t row, T v) {
  reinterpret_cast<T*>(p.arena + p.dyn->fix_off[col])[row] = v;
}

struct EmitSink {
  uint16_t* acl_out;  // next acl_gids slot
  uint8_t* agg_out;   // next aggregated_flow_ids character
  uint32_t n_agg;
  __device__ __forceinline__ void acl(uint32_t v) { *acl_out++ = (uint16_t)v; }
  __device__ void agg(uint64_t v) {  // strconv.FormatUint(n, 10) joined by ','
    if (n_agg++) *agg_out++ = ',';
    char tmp[20];
    int n = 0;
    do {
      tmp[n++] = (char)('0' + v % 10);
      v /= 10;
    } while (v);
    while (n) *agg_out++ = (uint8_t)tmp[--n];
  }
};


struct SlotSink4 {
  uint32_t* w;                // s_words + tid, stride D_THREADS
  unsigned long long* elems;  // this record's block of captured values
  __device__ __forceinline__ void u32(uint32_t s, uint32_t v) { w[s * D_THREADS] = v; }
  __device__ __forceinline__ void u64(uint32_t s, uint32_t lo, uint32_t hi) {
    w[s * D_THREADS] = lo;
    w[(s + 1) * D_THREADS] = hi;
  }
  __device__ __forceinline__ void add(uint32_t s, uint32_t v) { w[s * D_THREADS] += v; }
  __device__ __forceinline__ uint32_t word(uint32_t s) const { return w[s * D_THREADS]; }
  __device__ __forceinline__ void ref(uint32_t, uint32_t, uint32_t) {}
  __device__ __forceinline__ void elem(uint32_t, uint32_t, uint32_t, uint32_t) {}
  __device__ __forceinline__ void velem(uint32_t base, uint32_t i, uint32_t lo, uint32_t hi) {
    if (i < L4E_CAP) elems[base + i] = (unsigned long long)hi << 32 | lo;
  }
};
struct SlotView4 {
  const uint32_t* wp;
  __device__ __forceinline__ uint32_t w(int s) const { return wp[s * D_THREADS]; }
  __device__ __forceinline__ uint64_t w64(int s) const { return (uint64_t)w(s + 1) << 32 | w(s); }
};
struct RowEmit4 {
  const BatchParams& p;
  const L4Dev* dv;
  const uint64_t* s_fix;
  bool ok, in_range;
  uint32_t rec, row;
  template <class T>
  __device__ __forceinline__ void fix(int col, T v) {
    if (ok) reinterpret_cast<T*>(p.arena + s_fix[col])[row] = v;
  }
  __device__ __forceinline__ void ip6(int col, uint32_t off, uint32_t len) {
    // ckdb.AppendIPv6 (ckdb.go:46-56): nil -> zeros, short -> zero padded
    uint64_t lo = 0, hi = 0;
    const uint32_t n = len > 16 ? 16 : len;
    for (uint32_t i = 0; i < n; i++) {
      const uint64_t v = (uint64_t)__ldg(p.in + off + i) << ((i & 7) * 8);
      if (i < 8) lo |= v; else hi |= v;
    }
    if (ok) reinterpret_cast<ulonglong2*>(p.arena + s_fix[col])[row] = make_ulonglong2(lo, hi);
  }
  __device__ __forceinline__ void kg(uint32_t vtap, uint32_t tap_side, uint32_t proto, uint32_t port_dst, bool v6, bool vip0,
                                     bool vip1, uint32_t ip0, uint32_t ip1, uint32_t epc0, uint32_t epc1, uint64_t mac0,
                                     uint64_t mac1, uint32_t gpid0, uint32_t gpid1, uint32_t o6s, uint32_t l6s, uint32_t o6d,
                                     uint32_t l6d) {
    if (!in_range) return;
    kg_write_inputs(dv->kgin, dv->rec_cap, rec, vtap, tap_side, proto, port_dst, v6, vip0, vip1, ip0, ip1, epc0, epc1, mac0,
                    mac1, gpid0, gpid1, 0u, 0u, 0u, 0u);
    dv->ip6[rec] = ok ? ((unsigned long long)l6s << 32 | o6s) : 0ull;
    dv->ip6[(size_t)dv->rec_cap + rec] = ok ? ((unsigned long long)l6d << 32 | o6d) : 0ull;
  }
};

}  // namespace

__global__ void __launch_bounds__(D_THREADS, 6) k_l4_decode(const __grid_constant__ BatchParams p) {
  __shared__ uint32_t s_tbl[M4_COUNT * L4_MAXF];
  __shared__ uint32_t s_words[W4_COUNT * D_THREADS];
  __shared__ uint64_t s_fix[L4C_COUNT];
  DynState* dyn = p.dyn;
  const L4Dev* dv = reinterpret_cast<const L4Dev*>(p.ext);
  for (uint32_t i = threadIdx.x; i < M4_COUNT * L4_MAXF; i += D_THREADS) s_tbl[i] = dv->tok_tbl[i];
  for (uint32_t i = threadIdx.x; i < L4C_COUNT; i += D_THREADS) s_fix[i] = dyn->fix_off[i];
  __syncthreads();
  if (dyn->overflow) return;
  const uint32_t nrec = batch_nrec(p), nwarps = dyn->nwarps;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * (D_THREADS / 32) + (threadIdx.x >> 5), total_warps = gridDim.x * (D_THREADS / 32);
  uint32_t* my_words = s_words + threadIdx.x;
  TokWin win;
  win.init(p.in, p.in_bytes);
  for (uint32_t wtile = gwarp; wtile < nwarps; wtile += total_warps) {
    const uint32_t rec = wtile * 32 + lane;
    const bool in_range = rec < nrec;
#pragma unroll
    for (int i = 0; i < W4_COUNT; i++) my_words[i * D_THREADS] = 0;
    SlotSink4 sink{my_words, reinterpret_cast<unsigned long long*>(dv->elems) + (size_t)(in_range ? rec : 0) * L4E_COUNT};
    bool ok = false;
    uint32_t frame = 0;
    if (in_range) {
      frame = p.rec_frame[rec];
      if (!(p.repair && rec >= p.frame_bad[frame])) {
        ok = tk_decode<L4_MAXF>(win, p.rec_off[rec], p.rec_len[rec], s_tbl, sink, 0u);
        if (!ok) {
          if (!p.repair) {
            atomicMin(&p.frame_bad[frame], rec);
            dyn->stats.need_repair = 1;
          }
          for (int i = 0; i < W4_COUNT; i++) my_words[i * D_THREADS] = 0;
        }
      }
    }
    // TaggedFlow.IsValid() (datatype/pb/flow_log.go:91-101) cannot fail: ResetAll keeps every sub-message
    // allocated.  Rows keep the input order: row = record index - records dropped in earlier frames.
    const uint32_t row = ok ? rec - p.frame_drop[frame] : 0xffffffffu;
    const SlotView4 sl{my_words};
    RowEmit4 em{p, dv, s_fix, ok, in_range, rec, row};
    const L4RowInfo ri = l4_build_row(sl, em, p.dict.analyzer_id, p.persist->id_counter + row + 1);
    if (ok) reinterpret_cast<uint16_t*>(p.arena + s_fix[L4C_TEAM_ID])[row] = (uint16_t)p.frames[frame].team_id;
#define WTOT(s, v)                                                        \
  {                                                                       \
    const uint32_t t_ = __reduce_add_sync(0xffffffffu, ok ? (v) : 0u);    \
    if (lane == 0) p.warp_tot[(size_t)(s) * p.warp_cap + wtile] = t_;     \
  }
    WTOT(L4S_ROWS, 1u)
    WTOT(L4S_AGG_BYTES, ri.agg_chars)
    WTOT(L4S_ACL_ELEMS, ri.n_acl)
    WTOT(L4S_DOM_BYTES, ri.dom_len)
#undef WTOT
    if (in_range) {
      const size_t st = dv->rec_cap;
      dv->dom[rec] = ok ? ((unsigned long long)ri.dom_len << 32 | ri.dom_off) : 0ull;
      dv->arr[L4ARR_ACL_N * st + rec] = ok ? ri.n_acl : 0u;
      dv->arr[L4ARR_AGG_N * st + rec] = ok ? ri.n_agg : 0u;
      dv->arr[L4ARR_AGG_CHARS * st + rec] = ok ? ri.agg_chars : 0u;
      dv->meta[rec] = (ok ? L4META_OK : 0u) | (ok ? ri.tside : 0u) << 8;
      dv->rec_row[rec] = row;
    }
    lc_note(dyn->stats.lc_first[LC_TAPSIDE], ok ? ri.tside : 0u, rec, ok);
  }
}

// province of both endpoints (Internet.Fill :531-534 always takes the v4 fields): a binary search over the
// sorted geo ranges per address, in a kernel of its own so that its dependent loads run at full occupancy
__global__ void __launch_bounds__(256) k_l4_geo(const __grid_constant__ BatchParams p) {
  DynState* dyn = p.dyn;
  const L4Dev* dv = reinterpret_cast<const L4Dev*>(p.ext);
  if (dyn->overflow) return;
  const uint32_t nrec = batch_nrec(p);
  const size_t cap = dv->rec_cap;
  for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < nrec; base += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t rec = (uint32_t)(base + threadIdx.x);
    const bool in_range = rec < nrec;
    const uint32_t meta = in_range ? dv->meta[rec] : 0u;
    const bool ok = (meta & L4META_OK) != 0;
    uint32_t prov0 = 0, prov1 = 0;
    if (ok) {
      prov0 = geo_code(p.geo, dv->kgin[KG_IP_SRC * cap + rec]);
      prov1 = geo_code(p.geo, dv->kgin[KG_IP_DST * cap + rec]);
      dv->meta[rec] = meta | prov0 << 16 | prov1 << 24;
    }
    lc_note(dyn->stats.lc_first[LC_PROVINCE_0], prov0, rec, ok);
    lc_note(dyn->stats.lc_first[LC_PROVINCE_1], prov1, rec, ok);
  }
}

// variable-length columns + the LowCardinality keys, after the scans: work item = (warp of 32 records, task)
//   0 request_domain (ColStr)   1 aggregated_flow_ids (ColStr: decimal join)   2 acl_gids (ColArr[uint16])
//   3 province_0 / province_1 / observation_point keys (ranks of first appearance, k_finish)
__global__ void __launch_bounds__(256) k_l4_pack(const __grid_constant__ BatchParams p) {
  DynState* dyn = p.dyn;
  const L4Dev* dv = reinterpret_cast<const L4Dev*>(p.ext);
  if ((dyn->stats.need_repair && !p.repair) || dyn->overflow) return;
  const uint32_t nrec = batch_nrec(p), nwarps = dyn->nwarps;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), total_warps = gridDim.x * (blockDim.x >> 5);
  const size_t cap = dv->rec_cap;
  const uint64_t nitems = (uint64_t)nwarps * 4;
  for (uint64_t item = gwarp; item < nitems; item += total_warps) {
    __syncwarp();
    const uint32_t wtile = (uint32_t)(item >> 2), task = (uint32_t)(item & 3);
    const uint32_t rec = wtile * 32 + lane;
    const bool in_range = rec < nrec;
    const uint32_t row = in_range ? dv->rec_row[rec] : 0xffffffffu;
    const bool ok = row != 0xffffffffu;
#define WBASE(s, e) ((uint64_t)p.warp_base[(size_t)(s) * p.warp_cap + wtile] + (e))
    if (task == 0) {
      const unsigned long long ref = ok ? dv->dom[rec] : 0ull;
      const uint32_t len = (uint32_t)(ref >> 32);
      const uint64_t off = WBASE(L4S_DOM_BYTES, warp_excl_scan(len));
      if (ok) {
        reinterpret_cast<dfgpu_pos*>(p.arena + dyn->aux_off[L4C_REQUEST_DOMAIN])[row] = dfgpu_pos{(int64_t)off, (int64_t)(off + len)};
        uint8_t* dst = p.arena + dyn->var_buf_off[2] + off;
        const uint8_t* src = p.in + (uint32_t)ref;
        for (uint32_t i = 0; i < len; i++) dst[i] = __ldg(src + i);
      }
    } else if (task == 1 || task == 2) {
      const uint32_t n_acl = ok ? dv->arr[L4ARR_ACL_N * cap + rec] : 0u, n_agg = ok ? dv->arr[L4ARR_AGG_N * cap + rec] : 0u;
      const uint32_t chars = ok ? dv->arr[L4ARR_AGG_CHARS * cap + rec] : 0u;
      const uint64_t agg_off = WBASE(L4S_AGG_BYTES, warp_excl_scan(task == 1 ? chars : 0u));
      const uint64_t acl_off = WBASE(L4S_ACL_ELEMS, warp_excl_scan(task == 2 ? n_acl : 0u));
      if (ok) {
        const unsigned long long* E = reinterpret_cast<const unsigned long long*>(dv->elems) + (size_t)rec * L4E_COUNT;
        EmitSink sink;
        sink.acl_out = reinterpret_cast<uint16_t*>(p.arena + dyn->var_buf_off[1]) + acl_off;
        sink.agg_out = p.arena + dyn->var_buf_off[0] + agg_off;
        sink.n_agg = 0;
        if (task == 1) {
          reinterpret_cast<dfgpu_pos*>(p.arena + dyn->aux_off[L4C_AGGREGATED_FLOW_IDS])[row] =
              dfgpu_pos{(int64_t)agg_off, (int64_t)(agg_off + chars)};
          if (n_agg && n_agg <= L4E_CAP) {
            for (uint32_t i = 0; i < n_agg; i++) sink.agg(E[L4E_AGG + i]);
          } else if (n_agg) {  // more values than a block holds: walk the record again, keep only this column
            EmitSink walk = sink;
            walk.acl_out = nullptr;
            L4Rec scratch;
            memset(&scratch, 0, sizeof scratch);
            dec_tagged_flow(p.in + p.rec_off[rec], p.rec_len[rec], scratch, walk);
          }
        } else {
          reinterpret_cast<uint64_t*>(p.arena + dyn->aux_off[L4C_ACL_GIDS])[row] = acl_off + n_acl;
          if (n_acl && n_acl <= L4E_CAP) {
            for (uint32_t i = 0; i < n_acl; i++) sink.acl((uint32_t)E[L4E_ACL + i]);
          } else if (n_acl) {
            EmitSink walk = sink;
            walk.agg_out = nullptr;
            L4Rec scratch;
            memset(&scratch, 0, sizeof scratch);
            dec_tagged_flow(p.in + p.rec_off[rec], p.rec_len[rec], scratch, walk);
          }
        }
      }
    } else if (ok) {
      const uint32_t meta = dv->meta[rec];
      (p.arena + dyn->fix_off[L4C_PROVINCE_0])[row] = dyn->lc_rank[LC_PROVINCE_0][(meta >> 16) & 0xffu];
      (p.arena + dyn->fix_off[L4C_PROVINCE_1])[row] = dyn->lc_rank[LC_PROVINCE_1][(meta >> 24) & 0xffu];
      (p.arena + dyn->fix_off[L4C_OBSERVATION_POINT])[row] = dyn->lc_rank[LC_TAPSIDE][(meta >> 8) & 0xffu];
    }
#undef WBASE
  }
}

__global__ void __launch_bounds__(256, 4) k_l4_kg(const __grid_constant__ BatchParams p) {
  __shared__ uint64_t s_fix[L4C_COUNT];
  for (uint32_t i = threadIdx.x; i < L4C_COUNT; i += blockDim.x) s_fix[i] = p.dyn->fix_off[i];
  __syncthreads();
  const L4Dev* dv = reinterpret_cast<const L4Dev*>(p.ext);
  KgSrc q;
  q.kgin = dv->kgin;
  q.ip6_src = dv->ip6;
  q.ip6_dst = dv->ip6 + dv->rec_cap;
  q.rec_row = dv->rec_row;
  q.rec_cap = dv->rec_cap;
  kg_rows<L4C_REGION_ID_0, -1>(p, q, s_fix);
}

// ---- host side ---------------------------------------------------------------------------------------------
int l4_prepare(dfgpu_ctx* ctx, cudaStream_t s) {
  if (!ctx->l4) {
    L4Host* h = new L4Host();
    std::vector<uint32_t> t = l4_tok_table();
    memcpy(h->dev.tok_tbl, t.data(), sizeof h->dev.tok_tbl);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    h->grid_decode = (unsigned)sms * 6;
    h->grid_pack = (unsigned)sms * 6;
    h->grid_kg = (unsigned)sms * 4;
    CU(cudaMalloc(&h->d_dev, sizeof(L4Dev)));
    cudaFuncSetAttribute(k_l4_decode, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    ctx->l4 = h;
  }
  L4Host* h = reinterpret_cast<L4Host*>(ctx->l4);
  const size_t cap = ctx->bp.rec_cap;
  if (cap > h->rec_cap) {  // first L4 batch of the context, or the record index grew
    if (h->d_rec_buf) CU(cudaFree(h->d_rec_buf));
    h->d_rec_buf = nullptr;
    const size_t per_rec = (size_t)L4E_COUNT * 8 + (size_t)KG_WORDS * 4 + 16 + 8 + (size_t)L4ARR_COUNT * 4 + 4 + 4;
    CU(cudaMalloc(&h->d_rec_buf, per_rec * cap + 4096));
    uint8_t* q = h->d_rec_buf;
    L4Dev& d = h->dev;
    d.elems = reinterpret_cast<uint64_t*>(q); q += (size_t)L4E_COUNT * 8 * cap;
    d.ip6 = reinterpret_cast<uint64_t*>(q); q += 16 * cap;
    d.dom = reinterpret_cast<uint64_t*>(q); q += 8 * cap;
    d.kgin = reinterpret_cast<uint32_t*>(q); q += (size_t)KG_WORDS * 4 * cap;
    d.arr = reinterpret_cast<uint32_t*>(q); q += (size_t)L4ARR_COUNT * 4 * cap;
    d.meta = reinterpret_cast<uint32_t*>(q); q += 4 * cap;
    d.rec_row = reinterpret_cast<uint32_t*>(q);
    d.rec_cap = (uint32_t)cap;
    h->rec_cap = cap;
  }
  CU(cudaMemcpyAsync(h->d_dev, &h->dev, sizeof(L4Dev), cudaMemcpyHostToDevice, s));
  ctx->bp.ext = h->d_dev;
  return 0;
}

void l4_free_state(dfgpu_ctx* ctx) {
  L4Host* h = reinterpret_cast<L4Host*>(ctx->l4);
  if (!h) return;
  cudaFree(h->d_rec_buf);
  cudaFree(h->d_dev);
  delete h;
  ctx->l4 = nullptr;
}

// decode (+ province look-ups): before the scans
void launch_l4_decode(dfgpu_ctx* ctx, cudaStream_t s) {
  L4Host* h = reinterpret_cast<L4Host*>(ctx->l4);
  k_l4_decode<<<h->grid_decode, D_THREADS, 0, s>>>(ctx->bp);
  k_l4_geo<<<h->grid_pack, 256, 0, s>>>(ctx->bp);
  ctx->launches += 2;
}
// variable-length columns, LowCardinality keys, SmartEncoding: after the scans
void launch_l4_pack(dfgpu_ctx* ctx, cudaStream_t s) {
  L4Host* h = reinterpret_cast<L4Host*>(ctx->l4);
  k_l4_pack<<<h->grid_pack, 256, 0, s>>>(ctx->bp);
  ctx->launches++;
}
void launch_l4_kg(dfgpu_ctx* ctx, cudaStream_t s) {
  L4Host* h = reinterpret_cast<L4Host*>(ctx->l4);
  k_l4_kg<<<h->grid_kg, 256, 0, s>>>(ctx->bp);
  ctx->launches++;
}

}  // namespace dfgpu
