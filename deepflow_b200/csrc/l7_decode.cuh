/*
 * l7_decode.cuh — device decoder for one flow_log.AppProtoLogsData record
 * (message/flow_log.proto:179-310).  Replaces the gogo-generated
 * (*AppProtoLogsData).Unmarshal on a fresh object whose Base and Base.Head are
 * pre-allocated (server/libs/datatype/pb/flow_log.go:23-30): req / resp /
 * trace_info / ext_info are "non-nil" exactly when they appear on the wire.
 * Strings are kept as (pointer,length) into the input frame; repeated fields are
 * counted here and streamed out by a second walk (walk_l7_repeated).
 */
#pragma once
#include "pb.cuh"

namespace dfgpu {

struct SRef {
  const uint8_t* p;
  uint32_t n;
};

struct L7Rec {
  // AppProtoLogsBaseInfo
  uint64_t start_time, end_time, flow_id, tap_port, mac_src, mac_dst;
  uint32_t vtap_id, tap_type, is_ipv6, tap_side;
  uint32_t ip_src, ip_dst;
  SRef ip6_src, ip6_dst;
  int32_t l3_epc_id_src, l3_epc_id_dst;
  uint32_t port_src, port_dst, protocol, is_vip_interface_src, is_vip_interface_dst;
  uint32_t req_tcp_seq, resp_tcp_seq, process_id_0, process_id_1;
  SRef process_kname_0, process_kname_1;
  uint64_t syscall_trace_id_request, syscall_trace_id_response, syscall_coroutine_0, syscall_coroutine_1;
  uint32_t syscall_trace_id_thread_0, syscall_trace_id_thread_1, syscall_cap_seq_0, syscall_cap_seq_1;
  uint32_t gpid_0, gpid_1, pod_id_0, pod_id_1, biz_type;
  // AppProtoHead
  uint32_t head_proto, head_msg_type;
  uint64_t head_rrt;
  // AppProtoLogsData
  int32_t req_len, resp_len;
  SRef version;
  uint32_t row_effect, direction_score, flags, captured_request_byte, captured_response_byte;
  SRef biz_code, biz_scenario, biz_response_code;
  uint32_t has_req, has_resp, has_trace, has_ext;
  // L7Request
  SRef req_type, domain, resource, endpoint;
  // L7Response
  uint32_t resp_status;
  int32_t resp_code;
  SRef exception, result;
  // TraceInfo
  SRef trace_id, span_id, parent_span_id, trace_ids0, trace_ids1;
  uint32_t n_trace_ids;
  // ExtendedInfo
  SRef service_name, client_ip, x_request_id_0, http_user_agent, http_referer, rpc_service, protocol_str,
      x_request_id_1;
  uint32_t request_id;
  uint32_t n_attr_names, b_attr_names, n_attr_values, b_attr_values, n_metrics_names, b_metrics_names,
      n_metrics_values;
};

#define PB_SREF(dst)                              \
  {                                               \
    uint32_t _n;                                  \
    if (wt != 2 || !rd_len(r, _n)) return false;  \
    (dst).p = r.p + r.i;                          \
    (dst).n = _n;                                 \
    r.i += _n;                                    \
    break;                                        \
  }

__device__ inline bool dec_l7_head(Rd r, L7Rec& o) {
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_U32(o.head_proto)
      case 2: PB_U32(o.head_msg_type)
      case 5: PB_U64(o.head_rrt)
      default: PB_SKIP()
    }
  }
  return true;
}

__device__ inline bool dec_l7_base(Rd r, L7Rec& o) {
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_U64(o.start_time)
      case 2: PB_U64(o.end_time)
      case 3: PB_U64(o.flow_id)
      case 4: PB_U64(o.tap_port)
      case 5: PB_U32(o.vtap_id)
      case 6: PB_U32(o.tap_type)
      case 7: PB_U32(o.is_ipv6)
      case 8: PB_U32(o.tap_side)
      case 9: {
        uint32_t n;
        if (wt != 2 || !rd_len(r, n)) return false;
        Rd s{r.p + r.i, 0, n};
        if (!dec_l7_head(s, o)) return false;
        r.i += n;
        break;
      }
      case 10: PB_U64(o.mac_src)
      case 11: PB_U64(o.mac_dst)
      case 12: PB_U32(o.ip_src)
      case 13: PB_U32(o.ip_dst)
      case 14: PB_SREF(o.ip6_src)
      case 15: PB_SREF(o.ip6_dst)
      case 16: PB_I32(o.l3_epc_id_src)
      case 17: PB_I32(o.l3_epc_id_dst)
      case 18: PB_U32(o.port_src)
      case 19: PB_U32(o.port_dst)
      case 20: PB_U32(o.protocol)
      case 21: PB_U32(o.is_vip_interface_src)
      case 22: PB_U32(o.is_vip_interface_dst)
      case 23: PB_U32(o.req_tcp_seq)
      case 24: PB_U32(o.resp_tcp_seq)
      case 25: PB_U32(o.process_id_0)
      case 26: PB_U32(o.process_id_1)
      case 27: PB_SREF(o.process_kname_0)
      case 28: PB_SREF(o.process_kname_1)
      case 29: PB_U64(o.syscall_trace_id_request)
      case 30: PB_U64(o.syscall_trace_id_response)
      case 31: PB_U32(o.syscall_trace_id_thread_0)
      case 32: PB_U32(o.syscall_trace_id_thread_1)
      case 33: PB_U32(o.syscall_cap_seq_0)
      case 34: PB_U32(o.syscall_cap_seq_1)
      case 35: PB_U32(o.gpid_0)
      case 36: PB_U32(o.gpid_1)
      case 39: PB_U64(o.syscall_coroutine_0)
      case 40: PB_U64(o.syscall_coroutine_1)
      case 41: PB_U32(o.pod_id_0)
      case 42: PB_U32(o.pod_id_1)
      case 43: PB_U32(o.biz_type)
      default: PB_SKIP()
    }
  }
  return true;
}

__device__ inline bool dec_l7_req(Rd r, L7Rec& o) {
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_SREF(o.req_type)
      case 2: PB_SREF(o.domain)
      case 3: PB_SREF(o.resource)
      case 4: PB_SREF(o.endpoint)
      default: PB_SKIP()
    }
  }
  return true;
}

__device__ inline bool dec_l7_resp(Rd r, L7Rec& o) {
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_U32(o.resp_status)
      case 2: PB_I32(o.resp_code)
      case 3: PB_SREF(o.exception)
      case 4: PB_SREF(o.result)
      default: PB_SKIP()
    }
  }
  return true;
}

__device__ inline bool dec_l7_trace(Rd r, L7Rec& o) {
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_SREF(o.trace_id)
      case 2: PB_SREF(o.span_id)
      case 3: PB_SREF(o.parent_span_id)
      case 4: {  // repeated string trace_ids: only [0] and [1] are read (l7_flow_log.go:349-356)
        uint32_t n;
        if (wt != 2 || !rd_len(r, n)) return false;
        if (o.n_trace_ids == 0) { o.trace_ids0.p = r.p + r.i; o.trace_ids0.n = n; }
        else if (o.n_trace_ids == 1) { o.trace_ids1.p = r.p + r.i; o.trace_ids1.n = n; }
        o.n_trace_ids++;
        r.i += n;
        break;
      }
      default: PB_SKIP()
    }
  }
  return true;
}

// Sink of the repeated ExtendedInfo fields (16..19).
struct L7NoSink {
  __device__ __forceinline__ void attr_name(const uint8_t*, uint32_t) {}
  __device__ __forceinline__ void attr_value(const uint8_t*, uint32_t) {}
  __device__ __forceinline__ void metric_name(const uint8_t*, uint32_t) {}
  __device__ __forceinline__ void metric_value(uint64_t) {}
};

template <class Sink>
__device__ inline bool dec_l7_ext(Rd r, L7Rec& o, Sink& sink) {
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_SREF(o.service_name)
      case 2: PB_SREF(o.client_ip)
      case 3: PB_U32(o.request_id)
      case 4: PB_SREF(o.x_request_id_0)
      case 6: PB_SREF(o.http_user_agent)
      case 7: PB_SREF(o.http_referer)
      case 8: PB_SREF(o.rpc_service)
      case 9: PB_SREF(o.protocol_str)
      case 10: PB_SREF(o.x_request_id_1)
      case 16:
      case 17:
      case 18: {
        uint32_t n;
        if (wt != 2 || !rd_len(r, n)) return false;
        const uint8_t* p = r.p + r.i;
        if (f == 16) { o.n_attr_names++; o.b_attr_names += n; sink.attr_name(p, n); }
        else if (f == 17) { o.n_attr_values++; o.b_attr_values += n; sink.attr_value(p, n); }
        else { o.n_metrics_names++; o.b_metrics_names += n; sink.metric_name(p, n); }
        r.i += n;
        break;
      }
      case 19: {  // repeated double
        if (wt == 1) {
          if (r.l - r.i < 8) return false;
          uint64_t v = 0;
          for (int k = 0; k < 8; k++) v |= (uint64_t)ldb(r.p + r.i + k) << (8 * k);
          r.i += 8;
          o.n_metrics_values++;
          sink.metric_value(v);
        } else if (wt == 2) {
          uint32_t n;
          if (!rd_len(r, n)) return false;
          uint32_t end = r.i + n;
          while (r.i < end) {
            if (r.l - r.i < 8) return false;
            uint64_t v = 0;
            for (int k = 0; k < 8; k++) v |= (uint64_t)ldb(r.p + r.i + k) << (8 * k);
            r.i += 8;
            o.n_metrics_values++;
            sink.metric_value(v);
          }
        } else {
          return false;
        }
        break;
      }
      default: PB_SKIP()
    }
  }
  return true;
}

template <class Sink>
__device__ inline bool dec_app_proto_log(const uint8_t* rec, uint32_t len, L7Rec& o, Sink& sink) {
  Rd r{rec, 0, len};
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1:
      case 11:
      case 12:
      case 14:
      case 15: {
        uint32_t n;
        if (wt != 2 || !rd_len(r, n)) return false;
        Rd s{r.p + r.i, 0, n};
        bool ok;
        if (f == 1) ok = dec_l7_base(s, o);
        else if (f == 11) { o.has_req = 1; ok = dec_l7_req(s, o); }
        else if (f == 12) { o.has_resp = 1; ok = dec_l7_resp(s, o); }
        else if (f == 14) { o.has_trace = 1; ok = dec_l7_trace(s, o); }
        else { o.has_ext = 1; ok = dec_l7_ext(s, o, sink); }
        if (!ok) return false;
        r.i += n;
        break;
      }
      case 9: PB_I32(o.req_len)
      case 10: PB_I32(o.resp_len)
      case 13: PB_SREF(o.version)
      case 16: PB_U32(o.row_effect)
      case 17: PB_U32(o.direction_score)
      case 18: PB_U32(o.flags)
      case 19: PB_U32(o.captured_request_byte)
      case 20: PB_U32(o.captured_response_byte)
      case 21: PB_SREF(o.biz_code)
      case 22: PB_SREF(o.biz_scenario)
      case 23: PB_SREF(o.biz_response_code)
      default: PB_SKIP()
    }
  }
  return true;
}

}  // namespace dfgpu
