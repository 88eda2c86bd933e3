/*
 * gpu_decoder.cc — see gpu_decoder.h.  Also exports a small C surface (dfhost_*) so the
 * Python tests can drive the queue and the decoder the way the reference's own tests do
 * (server/libs/queue/queue_test.go:54-176).
 */
#include "gpu_decoder.h"

#include "../../include/dfgpu_wire.h"

#include <cstdlib>
#include <cstring>

namespace dfhost {

GpuDecoder::GpuDecoder(int index, int msg_type, int device, const dfgpu_cfg* cfg, OverwriteQueue* in_queue,
                       std::function<void(RecvBuffer*)> release_recv_buffer, BlockSink* sink)
    : index_(index), msg_type_(msg_type), in_(in_queue), release_(std::move(release_recv_buffer)), sink_(sink) {
  if (dfgpu_init(device, cfg, &ctx_) != DFGPU_OK) ctx_ = nullptr;  // no CPU fallback: the caller sees ok() == false
  buffer_.resize(BUFFER_SIZE);
}

GpuDecoder::~GpuDecoder() {
  if (ctx_) dfgpu_close(ctx_);
}

void GpuDecoder::submit_org(uint16_t org, const std::vector<RecvBuffer*>& bufs) {
  size_t total = 0;
  for (auto* rb : bufs) total += (size_t)(rb->end - rb->begin);
  if (staging_.size() < total + 64) staging_.resize(total + total / 4 + 4096);
  frames_.clear();
  size_t used = 0;
  for (auto* rb : bufs) {
    const size_t n = (size_t)(rb->end - rb->begin);
    memcpy(staging_.data() + used, rb->buffer + rb->begin, n);
    dfgpu_frame fr;
    memset(&fr, 0, sizeof fr);
    fr.offset = used;
    fr.len = (uint32_t)n;
    fr.vtap_id = rb->vtap_id;
    fr.org_id = rb->org_id;
    fr.team_id = rb->team_id;
    frames_.push_back(fr);
    used += n;
  }
  dfgpu_batch* batch = nullptr;
  int rc = dfgpu_submit(ctx_, msg_type_, staging_.data(), used, frames_.data(), (int)frames_.size(), &id_counter_,
                        msg_type_ == DFGPU_MSG_PROTOCOLLOG ? &trace_id_index_ : nullptr, &batch);
  counter_.submits++;
  if (rc != DFGPU_OK) {  // never returned upward (decoder.go:206-209): the frames count as errors
    counter_.error_count += (int64_t)bufs.size();
    return;
  }
  dfgpu_counters c;
  dfgpu_batch_counters(batch, &c);
  counter_.count += c.count;
  counter_.error_count += c.error_count;
  counter_.drop_count += c.drop_count;
  counter_.l7_http += c.l7_http_count;
  counter_.l7_dns += c.l7_dns_count;
  counter_.l7_sql += c.l7_sql_count;
  counter_.l7_nosql += c.l7_nosql_count;
  counter_.l7_rpc += c.l7_rpc_count;
  counter_.l7_mq += c.l7_mq_count;
  if (sink_) sink_->PutBlocks(msg_type_, org, batch);
  else dfgpu_batch_release(batch);
}

bool GpuDecoder::RunOnce() {
  if (closed_) return false;
  const size_t n = in_->Gets(buffer_.data(), buffer_.size());  // decoder.go:161
  // one submit per org, frames in arrival order (the reference handles one RecvBuffer at a time;
  // rows of different orgs go to different databases anyway, ckwriter.go:221-227 OrgID())
  std::map<uint16_t, std::vector<RecvBuffer*>> by_org;
  bool flush = false;
  for (size_t i = 0; i < n; i++) {
    if (buffer_[i] == nullptr) {  // flush tick (decoder.go:164-168)
      flush = true;
      continue;
    }
    counter_.raw_count++;
    RecvBuffer* rb = reinterpret_cast<RecvBuffer*>(buffer_[i]);
    by_org[rb->org_id].push_back(rb);
  }
  for (auto& kv : by_org) {
    submit_org(kv.first, kv.second);
    for (auto* rb : kv.second) release_(rb);  // receiver.ReleaseRecvBuffer, exactly once (decoder.go:196)
  }
  if (flush) {
    counter_.flushes++;
    dfgpu_flush(ctx_);
    if (sink_) sink_->Flush();
  }
  return true;
}

void GpuDecoder::Run() {
  while (RunOnce()) {
  }
}

}  // namespace dfhost

// ---- C surface for the tests -----------------------------------------------------------------------
using namespace dfhost;

namespace {
struct TestQueue {
  long long released_product = 1;
  long long released_count = 0;
  OverwriteQueue* q = nullptr;
};
struct CollectSink : BlockSink {
  std::vector<dfgpu_batch*> batches;
  int64_t rows[DFGPU_TABLE_MAX] = {0};
  int flushes = 0;
  void PutBlocks(int, uint16_t, dfgpu_batch* b) override {
    dfgpu_counters c;
    dfgpu_batch_counters(b, &c);
    for (int t = 0; t < DFGPU_TABLE_MAX; t++) rows[t] += c.rows[t];
    batches.push_back(b);
  }
  void Flush() override { flushes++; }
  ~CollectSink() override {
    for (auto* b : batches) dfgpu_batch_release(b);
  }
};
struct TestDecoder {
  OverwriteQueue* q = nullptr;
  CollectSink sink;
  GpuDecoder* dec = nullptr;
  int64_t acquired = 0, released = 0;
};
}  // namespace

extern "C" {

void* dfhost_queue_new(int size, int with_release, int flush_ms) {
  auto* t = new TestQueue();
  OverwriteQueue::Release rel = nullptr;
  if (with_release)
    rel = [t](OverwriteQueue::Item x) {
      t->released_product *= (long long)(intptr_t)x;
      t->released_count++;
    };
  t->q = new OverwriteQueue((size_t)size, rel, flush_ms);
  return t;
}
void dfhost_queue_free(void* h) {
  auto* t = (TestQueue*)h;
  delete t->q;
  delete t;
}
int dfhost_queue_put(void* h, const long long* vals, int n) {
  auto* t = (TestQueue*)h;
  std::vector<OverwriteQueue::Item> items((size_t)n);
  for (int i = 0; i < n; i++) items[(size_t)i] = (OverwriteQueue::Item)(intptr_t)vals[i];
  return t->q->Put(items.data(), (size_t)n) ? 0 : -1;
}
long long dfhost_queue_get(void* h) { return (long long)(intptr_t)((TestQueue*)h)->q->Get(); }
int dfhost_queue_gets(void* h, long long* out, int cap) {
  auto* t = (TestQueue*)h;
  std::vector<OverwriteQueue::Item> items((size_t)cap + 1);
  size_t n = t->q->Gets(items.data(), (size_t)cap);
  for (size_t i = 0; i < n; i++) out[i] = (long long)(intptr_t)items[i];
  return (int)n;
}
int dfhost_queue_len(void* h) { return (int)((TestQueue*)h)->q->Len(); }
int dfhost_queue_size(void* h) { return (int)((TestQueue*)h)->q->Size(); }
long long dfhost_queue_released(void* h, int reset) {
  auto* t = (TestQueue*)h;
  long long v = t->released_product;
  if (reset) t->released_product = 1;
  return v;
}

// decoder harness: queue of `qsize` RecvBuffers feeding one GpuDecoder
void* dfhost_decoder_new(int index, int msg_type, int device, const dfgpu_cfg* cfg, int qsize) {
  auto* t = new TestDecoder();
  auto release = [t](RecvBuffer* rb) {
    free(rb->buffer);
    delete rb;
    t->released++;
  };
  t->q = new OverwriteQueue((size_t)qsize, [release](OverwriteQueue::Item x) { release((RecvBuffer*)x); }, 0);
  t->dec = new GpuDecoder(index, msg_type, device, cfg, t->q, release, &t->sink);
  if (!t->dec->ok()) {
    delete t->dec;
    delete t->q;
    delete t;
    return nullptr;
  }
  return t;
}
void dfhost_decoder_free(void* h) {
  auto* t = (TestDecoder*)h;
  delete t->dec;
  // whatever is still queued is released by draining
  while (t->q->Len()) {
    OverwriteQueue::Item it = t->q->Get();
    if (it) {
      auto* rb = (RecvBuffer*)it;
      free(rb->buffer);
      delete rb;
      t->released++;
    }
  }
  delete t->q;
  delete t;
}
int dfhost_decoder_dict_load(void* h, uint16_t org, const dfgpu_dict_snapshot* s) { return ((TestDecoder*)h)->dec->LoadDict(org, s); }
int dfhost_decoder_geo_load(void* h, const dfgpu_geo_entry* e, uint64_t n, const char* const* names, uint32_t nn) {
  return ((TestDecoder*)h)->dec->LoadGeo(e, n, names, nn);
}
// enqueue frames [first, first+count) of a workload as RecvBuffers (payload copied at a non-zero
// Begin offset, like the receiver's pooled buffers); flush != 0 appends a nil flush tick
int dfhost_decoder_feed(void* h, const uint8_t* data, const dfgpu_frame* tbl, int first, int count, int flush) {
  auto* t = (TestDecoder*)h;
  for (int i = first; i < first + count; i++) {
    auto* rb = new RecvBuffer();
    const int pad = 19;  // the message headers precede the payload in the real buffer (droplet-message.go:150-219)
    rb->buffer = (uint8_t*)malloc((size_t)tbl[i].len + pad + 8);
    memcpy(rb->buffer + pad, data + tbl[i].offset, tbl[i].len);
    rb->begin = pad;
    rb->end = pad + (int)tbl[i].len;
    rb->vtap_id = tbl[i].vtap_id;
    rb->org_id = tbl[i].org_id;
    rb->team_id = tbl[i].team_id;
    t->acquired++;
    t->q->Put(rb);
  }
  if (flush) t->q->Put(nullptr);
  return 0;
}
// run Gets() batches until the queue is empty; returns the number of batches
int dfhost_decoder_drain(void* h) {
  auto* t = (TestDecoder*)h;
  int n = 0;
  while (t->q->Len()) {
    t->dec->RunOnce();
    n++;
  }
  return n;
}
void dfhost_decoder_stats(void* h, int64_t* out /*[16 + DFGPU_TABLE_MAX]*/) {
  auto* t = (TestDecoder*)h;
  const Counter& c = t->dec->counter();
  int64_t v[16] = {c.raw_count, c.count, c.error_count, c.drop_count, c.l7_http, c.l7_dns, c.l7_sql, c.l7_nosql,
                   c.l7_rpc, c.l7_mq, c.submits, c.flushes, t->acquired, t->released, (int64_t)t->sink.batches.size(),
                   (int64_t)t->sink.flushes};
  memcpy(out, v, sizeof v);
  memcpy(out + 16, t->sink.rows, sizeof t->sink.rows);
}
// A1 (include/dfgpu_wire.h): one received message -> frame fields; out = {rc, msg_type, payload_off, len, vtap,
// org, team, encoder, invalid}
void dfhost_parse_message(const uint8_t* msg, uint64_t len, int64_t* out) {
  dfgpu_frame fr{};
  int mt = -1;
  size_t off = 0;
  uint8_t enc = 0;
  uint64_t invalid = 0;
  const int rc = dfgpu_frame_from_message(msg, (size_t)len, &fr, &mt, &off, &enc, &invalid);
  const int64_t v[9] = {rc, mt, (int64_t)off, fr.len, fr.vtap_id, fr.org_id, fr.team_id, enc, (int64_t)invalid};
  memcpy(out, v, sizeof v);
}
// i-th block handed to the sink (owned by the harness)
dfgpu_batch* dfhost_decoder_batch(void* h, int i) {
  auto* t = (TestDecoder*)h;
  return i >= 0 && (size_t)i < t->sink.batches.size() ? t->sink.batches[(size_t)i] : nullptr;
}

}  // extern "C"
