/*
 * ORACLE — test infrastructure only.
 * oracle_api.cc: C entry points (ctypes) that drive the CPU restatement the way
 *   Decoder.Run / handleTaggedFlow / handleProtoLog   decoder.go:151-234
 *   SimpleDecoder.ReadPB / IsEnd / Failed              libs/codec/simple_codec.go:229-248,400-406
 * drive the reference: one frame (= RecvBuffer) after another, `{u32 LE len, pb}`
 * records inside a frame, a malformed record abandons the rest of its frame.
 * The entry points mirror include/dfgpu.h one-to-one so the parity tests can run
 * the same call sequence against both libraries.
 */
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
    c->metrics.reset(new MetricsBlocks());
    c->metrics->init();
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
        c->counters.count++;
        c->metrics->append_record(m, fr.org_id, (uint16_t)fr.team_id, pd, c->counters);
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
    for (int t = 0; t < 9; t++) c->counters.rows[DFGPU_TABLE_METRICS_BASE + t] = c->metrics->tables[t].nrows;
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
  if (table == DFGPU_TABLE_METRICS_BASE + 15 && c->metrics) return &c->metrics->reduced;
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

}  // extern "C"
