/*
 * ORACLE — test infrastructure only.
 * PINNING (what ties this restatement to the reference; full list: DESIGN.md §4).  Pinned by the reference's own
 * tests / sources: the wire schema (field lists machine-extracted from message/*.proto), every output column list
 * (machine-extracted from the ToInput functions), tag Code / table codes, trident enums, L7Protocol.String, the
 * TaggedFlow / MiniTag literals, meter merge KATs, TAP-side strings, DJBHash, lru order, netmask tree, queue
 * semantics.  PARITY UNPINNED by the reference (it has no tests for them): KnowledgeGraph.fill, DocumentExpand,
 * GenerateNewFlowTags, every AppendToColumnBlock, genID, getStatus — for these a second, independent Python
 * restatement (tests/kg_second_opinion.py, tests/test_*_crosscheck_cpu.py) is compared with this oracle instead.
 * oracle_api.cc: C entry points (ctypes) that drive the CPU restatement the way
 *   Decoder.Run / handleTaggedFlow / handleProtoLog   decoder.go:151-234
 *   SimpleDecoder.ReadPB / IsEnd / Failed              libs/codec/simple_codec.go:229-248,400-406
 * drive the reference: one frame (= RecvBuffer) after another, `{u32 LE len, pb}`
 * records inside a frame, a malformed record abandons the rest of its frame.
 * The entry points mirror include/dfgpu.h one-to-one so the parity tests can run
 * the same call sequence against both libraries.
 */
#include <algorithm>
#include <unordered_map>
#include <map>
#include <memory>

#include "l7.h"
#include "metrics.h"

using namespace oracle;

struct oracle_ctx {
  dfgpu_cfg cfg{};
  std::map<uint16_t, std::unique_ptr<PlatformInfoTable>> tables;
  GeoTree geo;
  std::unique_ptr<L7State> l7state;
  std::unique_ptr<L7State> metrics_state;  // the unmarshaller's own AppServiceTagWriter cache
  std::unique_ptr<L4Block> l4;
  std::unique_ptr<L7Blocks> l7;
  std::unique_ptr<MetricsBlocks> metrics;
  std::unique_ptr<Msg> pbTaggedFlow, pbAppLog, pbDoc;
  dfgpu_counters counters{};
  std::vector<std::vector<dfgpu_col>> views;
};

extern "C" {

oracle_ctx* oracle_new(const dfgpu_cfg* cfg) {
  auto* c = new oracle_ctx();
  if (cfg) c->cfg = *cfg;
  L7Config lc;
  lc.trace_id_index_disabled = c->cfg.trace_id_index_disabled != 0;
  if (c->cfg.flow_tag_cache_flush_timeout) lc.flow_tag_cache_flush_timeout = c->cfg.flow_tag_cache_flush_timeout;
  if (c->cfg.flow_tag_cache_max_size) lc.flow_tag_cache_max_size = c->cfg.flow_tag_cache_max_size;
  c->l7state.reset(new L7State(lc));
  c->metrics_state.reset(new L7State(lc));
  c->pbTaggedFlow.reset(new Msg(&schema().tagged_flow));
  c->pbAppLog.reset(new Msg(&schema().app_log));
  c->pbDoc.reset(new Msg(&metric_schema().document));
  return c;
}
void oracle_free(oracle_ctx* c) { delete c; }

int oracle_dict_load(oracle_ctx* c, uint16_t org, const dfgpu_dict_snapshot* s) {
  auto t = std::make_unique<PlatformInfoTable>();
  t->load(*s);
  c->tables[org] = std::move(t);
  return 0;
}
int oracle_geo_load(oracle_ctx* c, const dfgpu_geo_entry* e, uint64_t n, const char* const* names,
                    uint32_t n_names) {
  c->geo.build(e, n, names, n_names);
  return 0;
}

static bool read_len(const uint8_t* buf, size_t len, size_t& off, const uint8_t*& rec, size_t& n) {
  // ReadPB simple_codec.go:229-240
  off += 4;
  if (off > len) return false;
  uint32_t l = (uint32_t)buf[off - 4] | (uint32_t)buf[off - 3] << 8 | (uint32_t)buf[off - 2] << 16 |
               (uint32_t)buf[off - 1] << 24;
  off += l;
  if (off > len) return false;
  rec = buf + off - l;
  n = l;
  return true;
}

int oracle_submit(oracle_ctx* c, int msg_type, const uint8_t* frames, size_t nbytes,
                  const dfgpu_frame* tbl, int nframes, uint32_t* id_counter, uint64_t* trace_idx) {
  (void)nbytes;
  memset(&c->counters, 0, sizeof c->counters);
  uint64_t dummy_trace = 0;
  if (!trace_idx) trace_idx = &dummy_trace;
  L7Counters l7c;
  if (msg_type == DFGPU_MSG_TAGGEDFLOW) {
    if (!c->l4) {
      c->l4.reset(new L4Block());
      c->l4->init();
    } else {
      c->l4->b.clear();
    }
  } else if (msg_type == DFGPU_MSG_PROTOCOLLOG) {
    if (!c->l7) {
      c->l7.reset(new L7Blocks());
      c->l7->init();
    } else {
      c->l7->l7.clear();
      c->l7->field.clear();
      c->l7->field_value.clear();
      c->l7->app_service.clear();
    }
  } else if (msg_type == DFGPU_MSG_METRICS) {
    if (!c->metrics) {
      c->metrics.reset(new MetricsBlocks());
      c->metrics->init();
    } else {
      c->metrics->clear();
    }
  } else {
    return DFGPU_EINVAL;
  }
  for (int fi = 0; fi < nframes; fi++) {
    const dfgpu_frame& fr = tbl[fi];
    c->counters.raw_count++;
    auto it = c->tables.find(fr.org_id);
    if (it == c->tables.end()) return DFGPU_ENOENT;
    const PlatformInfoTable* pd = it->second.get();
    const uint8_t* buf = frames + fr.offset;
    size_t len = fr.len, off = 0;
    while (off < len) {  // !decoder.IsEnd()
      const uint8_t* rec;
      size_t n;
      if (msg_type == DFGPU_MSG_TAGGEDFLOW) {
        Msg& m = *c->pbTaggedFlow;
        m.reset();  // ResetAll, datatype/pb/flow_log.go:122-165
        if (!read_len(buf, len, off, rec, n) || !decode_msg(m, rec, n)) {
          c->counters.error_count++;  // decoder.go:206-209
          break;
        }
        // IsValid (:91-101) holds: every sub-message is pre-allocated
        c->counters.count++;  // sendFlow :412
        c->l4->append_record(m, fr.org_id, (uint16_t)fr.team_id, pd, &c->geo, id_counter);
      } else if (msg_type == DFGPU_MSG_PROTOCOLLOG) {
        Msg& m = *c->pbAppLog;
        m.reset();
        if (!read_len(buf, len, off, rec, n) || !decode_msg(m, rec, n)) {
          c->counters.error_count++;  // decoder.go:225-229
          break;
        }
        c->l7->append_record(m, fr.org_id, (uint16_t)fr.team_id, pd, *c->l7state, id_counter, trace_idx, l7c);
        c->counters.count++;  // updateCounter :499
      } else {
        Msg& m = *c->pbDoc;
        m.reset();
        if (!read_len(buf, len, off, rec, n) || !decode_msg(m, rec, n)) {
          c->counters.error_count++;  // unmarshaller.go:236-240 (ErrDocCount)
          break;
        }
        // an unknown meter id is a decode error too: it abandons the frame (codec.go:60-63)
        uint8_t meter = (uint8_t)m.sub(3).u32(1);
        if (meter != MT_METER_FLOW && meter != MT_METER_USAGE && meter != MT_METER_APP) {
          c->counters.error_count++;
          break;
        }
        c->counters.count++;  // DocCount, unmarshaller.go:124
        c->metrics->append_record(m, fr.org_id, (uint16_t)fr.team_id, pd, c->counters, *c->metrics_state);
      }
    }
  }
  c->counters.l7_http_count = l7c.http;
  c->counters.l7_dns_count = l7c.dns;
  c->counters.l7_sql_count = l7c.sql;
  c->counters.l7_nosql_count = l7c.nosql;
  c->counters.l7_rpc_count = l7c.rpc;
  c->counters.l7_mq_count = l7c.mq;
  c->counters.new_field_count = l7c.new_field;
  c->counters.new_field_value_count = l7c.new_field_value;
  if (c->l4 && msg_type == DFGPU_MSG_TAGGEDFLOW) c->counters.rows[DFGPU_TABLE_L4_FLOW_LOG] = c->l4->b.nrows;
  if (c->l7 && msg_type == DFGPU_MSG_PROTOCOLLOG) {
    c->counters.rows[DFGPU_TABLE_L7_FLOW_LOG] = c->l7->l7.nrows;
    c->counters.rows[DFGPU_TABLE_L7_CUSTOM_FIELD] = c->l7->field.nrows;
    c->counters.rows[DFGPU_TABLE_L7_CUSTOM_FIELD_VALUE] = c->l7->field_value.nrows;
    c->counters.rows[DFGPU_TABLE_L7_APP_SERVICE] = c->l7->app_service.nrows;
  }
  if (c->metrics && msg_type == DFGPU_MSG_METRICS)
  {
    for (int t = 0; t < 9; t++) c->counters.rows[DFGPU_TABLE_METRICS_BASE + t] = c->metrics->tables[t].nrows;
    c->counters.rows[DFGPU_TABLE_METRICS_APP_SERVICE] = c->metrics->app_service.nrows;
  }
  return 0;
}

static Block* find_block(oracle_ctx* c, int table) {
  switch (table) {
    case DFGPU_TABLE_L4_FLOW_LOG: return c->l4 ? &c->l4->b : nullptr;
    case DFGPU_TABLE_L7_FLOW_LOG: return c->l7 ? &c->l7->l7 : nullptr;
    case DFGPU_TABLE_L7_CUSTOM_FIELD: return c->l7 ? &c->l7->field : nullptr;
    case DFGPU_TABLE_L7_CUSTOM_FIELD_VALUE: return c->l7 ? &c->l7->field_value : nullptr;
    case DFGPU_TABLE_L7_APP_SERVICE: return c->l7 ? &c->l7->app_service : nullptr;
  }
  if (table >= DFGPU_TABLE_METRICS_BASE && table < DFGPU_TABLE_METRICS_BASE + 9 && c->metrics)
    return &c->metrics->tables[table - DFGPU_TABLE_METRICS_BASE];
  if (table == DFGPU_TABLE_METRICS_APP_SERVICE && c->metrics) return &c->metrics->app_service;
  if (table == DFGPU_TABLE_METRICS_REDUCED && c->metrics) return &c->metrics->reduced;
  return nullptr;
}

int oracle_columns(oracle_ctx* c, int table, dfgpu_col* cols, int* ncols, uint64_t* nrows) {
  Block* b = find_block(c, table);
  if (!b) return DFGPU_ENOENT;
  if (ncols) *ncols = (int)b->cols.size();
  if (nrows) *nrows = b->nrows;
  if (cols)
    for (size_t i = 0; i < b->cols.size(); i++) cols[i] = b->cols[i].view();
  return 0;
}

int oracle_counters(oracle_ctx* c, dfgpu_counters* out) {
  *out = c->counters;
  return 0;
}

// reduce the rows of the last metrics submit (table `table`) by
// (time bucket, tag columns) with ConcurrentMerge semantics.
int oracle_rollup_add(oracle_ctx* c, int table, uint32_t bucket_seconds) {
  if (!c || !c->metrics) return DFGPU_ENOENT;
  return c->metrics->rollup_add(table - DFGPU_TABLE_METRICS_BASE, bucket_seconds);
}
int oracle_rollup_flush(oracle_ctx* c, int table) {
  if (!c || !c->metrics) return DFGPU_ENOENT;
  return c->metrics->rollup_flush(table - DFGPU_TABLE_METRICS_BASE);
}
// CPU baseline of BASELINE config 3 on DECODED rows (what dfgpu_rollup_add_rows consumes): packed tag
// bytes per row (key_stride, time already the bucket start) + row-major raw meters; ops = the
// ConcurrentMerge rule per field (0 sum u64, 1 sum u32, 2/3 max, 4 first, 5 last).  One thread, one hash
// map — the single-goroutine shape of the reference's merge.  Returns the number of groups; out_sums[nf]
// = per-field fold over all groups (a checksum the caller compares with the GPU's).
uint64_t oracle_rollup_rows(const uint8_t* keys, uint32_t key_stride, const uint64_t* raw, uint32_t nf,
                            const uint8_t* ops, uint64_t n, uint64_t* out_sums) {
  std::unordered_map<std::string, size_t> idx;
  std::vector<uint64_t> acc;
  idx.reserve((size_t)std::min<uint64_t>(n, 1u << 22));
  for (uint64_t r = 0; r < n; r++) {
    std::string k((const char*)keys + r * key_stride, key_stride);
    const uint64_t* v = raw + r * nf;
    auto it = idx.find(k);
    if (it == idx.end()) {
      idx.emplace(std::move(k), acc.size() / nf);
      acc.insert(acc.end(), v, v + nf);
      for (uint32_t f = 0; f < nf; f++)
        if (ops[f] == 1) acc[acc.size() - nf + f] = (uint32_t)v[f];
    } else {
      uint64_t* a = acc.data() + it->second * nf;
      for (uint32_t f = 0; f < nf; f++) {
        switch (ops[f]) {
          case 0: a[f] += v[f]; break;
          case 1: a[f] = (uint32_t)((uint32_t)a[f] + (uint32_t)v[f]); break;
          case 2:
          case 3: if (a[f] < v[f]) a[f] = v[f]; break;
          case 5: a[f] = v[f]; break;
          default: break;  // first: kept
        }
      }
    }
  }
  if (out_sums)
    for (uint32_t f = 0; f < nf; f++) {
      uint64_t t = 0;
      for (size_t g = 0; g < acc.size() / nf; g++) t += acc[g * nf + f];
      out_sums[f] = t;
    }
  return acc.size() / nf;
}

int oracle_metrics_reduce(oracle_ctx* c, int table, uint32_t bucket_seconds) {
  if (!c->metrics) return DFGPU_ENOENT;
  return c->metrics->reduce(table - DFGPU_TABLE_METRICS_BASE, bucket_seconds);
}

// ---- known-answer hooks for the reference's own unit tests (tests/test_reference_kats.py)
// libs/geo/netmask_tree_test.go:24-52 TestGetFirstMask
void oracle_kat_first_mask_sequence(uint32_t start, uint32_t end, uint32_t* prefixes, uint8_t* masklens, int n) {
  for (int k = 0; k < n; k++) {
    int maxLen = getFirstMask(start, end);
    prefixes[k] = start;
    masklens[k] = (uint8_t)maxLen;
    start += 1u << (uint32_t)(32 - maxLen);
  }
}
const char* oracle_kat_province(oracle_ctx* c, uint32_t ip) {
  static thread_local std::string s;
  s = std::string(c->geo.province(ip));
  return s.c_str();
}
int32_t oracle_kat_marshal_int32(int32_t v) { return MarshalInt32WithSpecialID(v); }
const char* oracle_kat_tapside(uint32_t v) {
  bool ok;
  return TAPSideString(v, ok);
}
int oracle_kat_is_local_ip(int v6, uint32_t ip4, const uint8_t* ip6, uint32_t len) { return IsLocalIP(v6, ip4, ip6, len); }
int oracle_kat_is_loopback(int v6, uint32_t ip4, const uint8_t* ip6, uint32_t len) { return IsLoopback(v6, ip4, ip6, len); }
// common.GetAutoInstance / GetAutoService / IsPodServiceIP (ingester/common/common.go:160-198): out = {id, type}
void oracle_kat_auto_instance(uint32_t pod, uint32_t gpid, uint32_t pod_node, uint32_t l3dev, uint32_t subnet,
                              uint32_t l3type, int32_t l3epc, uint32_t* out) {
  uint32_t id = 0;
  uint8_t type = 0;
  GetAutoInstance(pod, gpid, pod_node, l3dev, subnet, (uint8_t)l3type, l3epc, id, type);
  out[0] = id;
  out[1] = type;
}
void oracle_kat_auto_service(uint32_t custom, uint32_t pod_service, uint32_t pod_group, uint32_t gpid,
                             uint32_t pod_cluster, uint32_t l3dev, uint32_t subnet, uint32_t l3type,
                             uint32_t pod_group_type, int32_t l3epc, uint32_t* out) {
  uint32_t id = 0;
  uint8_t type = 0;
  GetAutoService(custom, pod_service, pod_group, gpid, pod_cluster, l3dev, subnet, (uint8_t)l3type,
                 (uint8_t)pod_group_type, l3epc, id, type);
  out[0] = id;
  out[1] = type;
}
int oracle_kat_is_pod_service_ip(uint32_t device_type, uint32_t pod, uint32_t pod_node) {
  return IsPodServiceIP((uint8_t)device_type, pod, pod_node) ? 1 : 0;
}
// L7Protocol.String(isTLS) (datatype/flow.go:623-803), NUL terminated copy
const char* oracle_kat_l7_protocol_string(uint32_t p, int tls) {
  static thread_local std::string s;
  s = std::string(L7ProtocolString((uint8_t)p, tls != 0));
  return s.c_str();
}
uint64_t oracle_kat_djb(uint64_t base, const char* s, uint32_t n) { return DJBHash(base, std::string_view(s, n)); }
// libs/lru/lru_test.go TestKeys + eviction order: add keys[0..n) to a cache of `cap`,
// touch `touch` (AddOrGet) when >= 0, then list the keys oldest -> newest.
struct IntHash {
  size_t operator()(int v) const { return (size_t)v * 0x9E3779B97F4A7C15ull; }
};
int oracle_kat_lru(int cap, int n, int touch, int* out_keys) {
  LruCache<int, IntHash> c(cap);
  for (int i = 0; i < n; i++) c.Add(i, (uint32_t)i);
  if (touch >= 0) c.AddOrGet(touch, 0);
  return c.Keys(out_keys);
}

// flow-metrics/basic_meter_test.go:112-135 TestMerge, usage_meter_test.go:57-79
// TestVTAPMeterMerge: a = every uint64/uint32 field set to `na` (uint8 fields stay 0, like
// the test's reflect-based initMeter), b likewise with `nb`; a.ConcurrentMerge(b); the
// fields come back in struct order.  kind: 1 flow, 4 usage, 5 app.
extern "C++" {
template <class T, size_t N>
static int dump_fields(T (&p)[N], uint64_t* out) {
  for (size_t i = 0; i < N; i++) out[i] = p[i];
  return (int)N;
}
}
int oracle_kat_meter_merge(int kind, uint64_t na, uint64_t nb, uint64_t* out) {
  int n = 0;
  if (kind == MT_METER_FLOW) {
    MFlowMeter a, b;
    MFlowMeter* ms[2] = {&a, &b};
    uint64_t vs[2] = {na, nb};
    for (int k = 0; k < 2; k++) {
      MFlowMeter& m = *ms[k];
      uint64_t v = vs[k];
#define SETALL(M)                                                                                                   \
  M(packet_tx) M(packet_rx) M(byte_tx) M(byte_rx) M(l3_byte_tx) M(l3_byte_rx) M(l4_byte_tx) M(l4_byte_rx)            \
  M(new_flow) M(closed_flow) M(l7_request) M(l7_response) M(syn_count) M(synack_count) M(rtt_max) M(rtt_client_max)  \
  M(rtt_server_max) M(srt_max) M(art_max) M(rrt_max) M(cit_max) M(rtt_sum) M(rtt_client_sum) M(rtt_server_sum)       \
  M(srt_sum) M(art_sum) M(rrt_sum) M(cit_sum) M(rtt_count) M(rtt_client_count) M(rtt_server_count) M(srt_count)      \
  M(art_count) M(rrt_count) M(cit_count) M(retrans_tx) M(retrans_rx) M(zero_win_tx) M(zero_win_rx) M(retrans_syn)    \
  M(retrans_synack) M(client_rst_flow) M(server_rst_flow) M(server_syn_miss) M(client_ack_miss)                      \
  M(client_half_close_flow) M(server_half_close_flow) M(client_source_port_reuse) M(client_establish_reset)          \
  M(server_reset) M(server_queue_lack) M(server_establish_reset) M(tcp_timeout) M(client_establish_fail)             \
  M(server_establish_fail) M(tcp_establish_fail) M(tcp_transfer_fail) M(tcp_rst_fail) M(l7_client_error)             \
  M(l7_server_error) M(l7_timeout) M(ooo_tx) M(ooo_rx) M(flow_load)
#define M(f) m.f = (decltype(m.f))v;
      SETALL(M)
#undef M
    }
    a.ConcurrentMerge(b);
#define M(f) out[n++] = a.f;
    SETALL(M)
#undef M
    out[n++] = a.direction_score;
  } else if (kind == MT_METER_USAGE) {
    MUsageMeter a, b;
    // fillMetrics(hint) usage_meter_test.go:25-34
    auto fill = [](MUsageMeter& m, uint64_t h) {
      m.packet_tx = h * 3; m.packet_rx = h * 2; m.byte_tx = h * 97; m.byte_rx = h * 89;
      m.l3_byte_tx = h * 96; m.l3_byte_rx = h * 88; m.l4_byte_tx = h * 95; m.l4_byte_rx = h * 87;
    };
    fill(a, na);
    fill(b, nb);
    a.ConcurrentMerge(b);
    uint64_t f[] = {a.packet_tx, a.packet_rx, a.byte_tx, a.byte_rx, a.l3_byte_tx, a.l3_byte_rx, a.l4_byte_tx, a.l4_byte_rx};
    n = dump_fields(f, out);
  } else if (kind == MT_METER_APP) {
    MAppMeter a, b;
    MAppMeter* ms[2] = {&a, &b};
    uint64_t vs[2] = {na, nb};
    for (int k = 0; k < 2; k++) {
      MAppMeter& m = *ms[k];
      m.request = m.response = m.rrt_max = m.rrt_count = m.client_error = m.server_error = m.timeout = (uint32_t)vs[k];
      m.rrt_sum = vs[k];
    }
    a.ConcurrentMerge(b);
    uint64_t f[] = {a.request, a.response, a.direction_score, a.rrt_max, a.rrt_sum, a.rrt_count, a.client_error, a.server_error, a.timeout};
    n = dump_fields(f, out);
  }
  return n;
}
// flow-metrics/tag_test.go:259-352 TestEncodeMiniTag: decode a pb.MiniTag and run
// Tag.ReadFromPB; returns {code, ip4_0, ip4_1, l3_epc_0, l3_epc_1, vtap, protocol, server_port,
// role, tap_type, acl_gid, is_ipv4, ip6_0[15], ip6_1[15], pod_id_0, pod_id_1, tunnel_ip_id}
int oracle_kat_tag_read(const uint8_t* minitag, uint32_t n, int64_t* out) {
  Msg m(&metric_schema().mini_tag);
  m.reset();
  if (!decode_msg(m, minitag, n)) return -1;
  MTag t;
  TagReadFromPB(t, m);
  int64_t v[] = {(int64_t)t.code, t.ip4[0], t.ip4[1], t.l3_epc[0], t.l3_epc[1], t.vtap_id, t.protocol, t.server_port,
                 t.role, t.tap_type, t.acl_gid, t.is_ipv4, t.ip6[0][15], t.ip6[1][15], t.pod_id[0], t.pod_id[1],
                 t.tunnel_ip_id};
  for (size_t i = 0; i < sizeof v / sizeof v[0]; i++) out[i] = v[i];
  return (int)(sizeof v / sizeof v[0]);
}
// app.DecodePB (codec.go:28-65) of ONE Document, before DocumentExpand: hdr = {timestamp,
// flags, uint8(meter_id)}; tag as oracle_kat_tag_read plus {gpid_0, gpid_1, signal_source,
// tap_side, l7_protocol, biz_type, tap_port, tap_port_type, nat_source, tunnel_type, mac_0, mac_1, gtid};
// meter fields in struct order (flow: SETALL order + direction_score).  Returns the meter field count, -1 on a decode error.
int oracle_kat_doc_read(const uint8_t* doc, uint32_t n, uint32_t* hdr, int64_t* tag, uint64_t* meter) {
  Msg m(&metric_schema().document);
  m.reset();
  if (!decode_msg(m, doc, n)) return -1;
  hdr[0] = m.u32(1);
  hdr[1] = m.u32(4);
  hdr[2] = (uint8_t)m.sub(3).u32(1);
  MTag t;
  TagReadFromPB(t, m.sub(2));
  int64_t v[] = {(int64_t)t.code, t.ip4[0], t.ip4[1], t.l3_epc[0], t.l3_epc[1], t.vtap_id, t.protocol, t.server_port,
                 t.role, t.tap_type, t.acl_gid, t.is_ipv4, t.ip6[0][15], t.ip6[1][15], t.pod_id[0], t.pod_id[1],
                 t.tunnel_ip_id, t.gpid[0], t.gpid[1], t.signal_source, t.tap_side, t.l7_protocol, t.biz_type,
                 t.tap_port, t.tap_port_type, t.nat_source, t.tunnel_type, (int64_t)t.mac[0], (int64_t)t.mac[1], t.gtid};
  for (size_t i = 0; i < sizeof v / sizeof v[0]; i++) tag[i] = v[i];
  int k = 0;
  if (hdr[2] == MT_METER_FLOW) {
    MFlowMeter a;
    a.ReadFromPB(m.sub(3).sub(2));
#define M(f) meter[k++] = a.f;
    SETALL(M)
#undef M
    meter[k++] = a.direction_score;
  } else if (hdr[2] == MT_METER_USAGE) {
    MUsageMeter a;
    a.ReadFromPB(m.sub(3).sub(3));
    uint64_t f[] = {a.packet_tx, a.packet_rx, a.byte_tx, a.byte_rx, a.l3_byte_tx, a.l3_byte_rx, a.l4_byte_tx, a.l4_byte_rx};
    k = dump_fields(f, meter);
  } else if (hdr[2] == MT_METER_APP) {
    MAppMeter a;
    a.ReadFromPB(m.sub(3).sub(4));
    uint64_t f[] = {a.request, a.response, a.direction_score, a.rrt_max, a.rrt_sum, a.rrt_count, a.client_error, a.server_error, a.timeout};
    k = dump_fields(f, meter);
  }
  return k;
}
#undef SETALL
// Tag.TableID over a raw code (tag.go:920-931)
int oracle_kat_table_id(uint64_t code, int is_second) { return TagTableID(code, is_second != 0); }
// column names / container kinds of one metrics table, '\n'-joined "name:ch_type"
const char* oracle_kat_metrics_columns(int table) {
  static thread_local std::string s;
  Block b;
  MetricsBlocks::add_columns(b, table);
  s.clear();
  for (auto& c : b.cols) s += c.name + ":" + std::to_string(c.ch_type) + ":" + std::to_string(c.elem_size) + "\n";
  return s.c_str();
}

}  // extern "C"
