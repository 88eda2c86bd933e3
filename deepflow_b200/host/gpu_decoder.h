/*
 * gpu_decoder.h — host side of the drop-in: the decoder goroutine of the ingester, with its
 * body replaced by one dfgpu_submit per queue batch.  C++ mirror (no Go toolchain here; the
 * cgo version of the same loop is in INTEGRATION.md) of
 *   decoder.NewDecoder / (*Decoder).Run   server/ingester/flow_log/decoder/decoder.go:106-199
 *   Unmarshaller.QueueProcess             server/ingester/flow_metrics/unmarshaller/unmarshaller.go:215-275
 *   receiver.RecvBuffer                   server/libs/receiver/receiver.go:69-78
 * Contract kept: items come from a QueueReader in batches of up to BUFFER_SIZE = 1024
 * (decoder.go:56,155-161); nil = flush tick; every RecvBuffer is released exactly once
 * (decoder.go:196); errors are counted, never returned upward (decoder.go:206-209).
 */
#pragma once
#include <functional>
#include <map>
#include <vector>

#include "../../include/dfgpu.h"
#include "overwrite_queue.h"

namespace dfhost {

constexpr size_t BUFFER_SIZE = 1024;  // decoder.go:56

struct RecvBuffer {  // receiver.go:69-78 (the fields the decoders read)
  int begin = 0, end = 0;
  uint8_t* buffer = nullptr;
  uint16_t vtap_id = 0;
  uint32_t team_id = 0;
  uint16_t org_id = 0;
};

struct Counter {  // decoder.go:60-80
  int64_t raw_count = 0, count = 0, error_count = 0, drop_count = 0;
  int64_t l7_http = 0, l7_dns = 0, l7_sql = 0, l7_nosql = 0, l7_rpc = 0, l7_mq = 0;
  int64_t submits = 0, flushes = 0;
};

// The ckwriter hand-off: receives finished column blocks (one dfgpu_batch per submit) and owns
// them from then on (it calls dfgpu_batch_release when the INSERT is done).
struct BlockSink {
  virtual ~BlockSink() = default;
  virtual void PutBlocks(int msg_type, uint16_t org_id, dfgpu_batch* batch) = 0;
  virtual void Flush() {}
};

class GpuDecoder {
 public:
  GpuDecoder(int index, int msg_type, int device, const dfgpu_cfg* cfg, OverwriteQueue* in_queue,
             std::function<void(RecvBuffer*)> release_recv_buffer, BlockSink* sink);
  ~GpuDecoder();
  bool ok() const { return ctx_ != nullptr; }
  dfgpu_ctx* ctx() { return ctx_; }
  int LoadDict(uint16_t org, const dfgpu_dict_snapshot* snap) { return dfgpu_dict_load(ctx_, org, snap); }
  int LoadGeo(const dfgpu_geo_entry* e, uint64_t n, const char* const* names, uint32_t n_names) {
    return dfgpu_geo_load(ctx_, e, n, names, n_names);
  }
  void Run();      // until Close()
  bool RunOnce();  // one Gets() batch; false when closed
  void Close() { closed_ = true; }
  const Counter& counter() const { return counter_; }

 private:
  void submit_org(uint16_t org, const std::vector<RecvBuffer*>& bufs);
  int index_, msg_type_;
  dfgpu_ctx* ctx_ = nullptr;
  OverwriteQueue* in_;
  std::function<void(RecvBuffer*)> release_;
  BlockSink* sink_;
  std::vector<uint8_t> staging_;
  std::vector<dfgpu_frame> frames_;
  std::vector<OverwriteQueue::Item> buffer_;
  uint32_t id_counter_ = 0;     // L4FlowCounter / L7FlowLogCounter (l4_flow_log.go:947)
  uint64_t trace_id_index_ = 0; // lastTraceIdIndex (l7_flow_log.go:294)
  Counter counter_;
  bool closed_ = false;
};

}  // namespace dfhost
