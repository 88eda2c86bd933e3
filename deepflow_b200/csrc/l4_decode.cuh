/*
 * l4_decode.cuh — device decoder for one flow_log.TaggedFlow record
 * (message/flow_log.proto:14-177).  Replaces the gogo-generated
 * (*TaggedFlow).Unmarshal + pb.TaggedFlow.ResetAll
 * (server/libs/datatype/pb/flow_log.go:122-165): every sub-message is treated as
 * pre-allocated (never nil), repeated occurrences merge, last scalar wins.
 */
#pragma once
#include "pb.cuh"

namespace dfgpu {

struct L4Peer {  // FlowMetricsPeer
  uint64_t byte_count, l3_byte_count, l4_byte_count, packet_count, total_byte_count, total_packet_count;
  uint32_t tcp_flags;
  int32_t l3_epc_id;
  uint32_t is_l2_end, is_l3_end, is_vip_interface, real_ip, real_port, gpid;
};

struct L4Rec {
  // FlowKey
  uint32_t vtap_id, tap_type;
  uint64_t tap_port, mac_src, mac_dst;
  uint32_t ip_src, ip_dst;
  uint32_t ip6_src_off, ip6_src_len, ip6_dst_off, ip6_dst_len;
  uint32_t port_src, port_dst, proto;
  L4Peer peer[2];
  // TunnelField
  uint32_t tx_ip0, tx_ip1, rx_ip0, rx_ip1, tx_mac0, tx_mac1, rx_mac0, rx_mac1, tx_id, rx_id, tunnel_type, tier;
  // Flow
  uint64_t flow_id, start_time, end_time, duration;
  uint32_t vlan, eth_type, has_perf_stats, close_type, signal_source, is_new_flow, tap_side;
  uint32_t syn_seq, synack_seq, last_keepalive_seq, last_keepalive_ack;
  uint32_t direction_score, init_ipid;
  uint32_t domain_off, domain_len;
  uint32_t n_acl_gids;       // element count over all occurrences
  uint32_t agg_ids_chars;    // len(Uint64SliceToString(aggregated_flow_ids))
  uint32_t n_agg_ids;
  // FlowPerfStats
  uint32_t l7_protocol, l7_failed_count;
  // TCPPerfStats
  uint32_t rtt_client_max, rtt_server_max, srt_max, art_max, rtt, srt_sum, art_sum, srt_count, art_count;
  uint32_t syn_count, synack_count, cit_max, cit_sum, cit_count, fin_count;
  uint32_t tx_retrans, tx_zero_win, tx_ooo, rx_retrans, rx_zero_win, rx_ooo;
  // L7PerfStats
  uint32_t request_count, response_count, err_client_count, err_server_count, err_timeout, rrt_count,
      rrt_max, tls_rtt;
  uint64_t rrt_sum;
};

__device__ __forceinline__ uint32_t dec_digits(uint64_t v) {
  uint32_t n = 1;
  while (v >= 10) {
    v /= 10;
    n++;
  }
  return n;
}

__device__ inline bool dec_flow_key(Rd r, const uint8_t* base, L4Rec& o) {
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_U32(o.vtap_id)
      case 2: PB_U32(o.tap_type)
      case 3: PB_U64(o.tap_port)
      case 4: PB_U64(o.mac_src)
      case 5: PB_U64(o.mac_dst)
      case 6: PB_U32(o.ip_src)
      case 7: PB_U32(o.ip_dst)
      case 8: PB_STR(o.ip6_src_off, o.ip6_src_len)
      case 9: PB_STR(o.ip6_dst_off, o.ip6_dst_len)
      case 10: PB_U32(o.port_src)
      case 11: PB_U32(o.port_dst)
      case 12: PB_U32(o.proto)
      default: PB_SKIP()
    }
  }
  return true;
}

__device__ inline bool dec_peer(Rd r, L4Peer& o) {
  uint32_t f, wt, dummy32;
  uint64_t dummy64;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_U64(o.byte_count)
      case 2: PB_U64(o.l3_byte_count)
      case 3: PB_U64(o.l4_byte_count)
      case 4: PB_U64(o.packet_count)
      case 5: PB_U64(o.total_byte_count)
      case 6: PB_U64(o.total_packet_count)
      case 7: PB_U64(dummy64)  // first
      case 8: PB_U64(dummy64)  // last
      case 9: PB_U32(o.tcp_flags)
      case 10: PB_I32(o.l3_epc_id)
      case 11: PB_U32(o.is_l2_end)
      case 12: PB_U32(o.is_l3_end)
      case 13: PB_U32(dummy32)  // is_active_host
      case 14: PB_U32(dummy32)  // is_device
      case 15: PB_U32(o.is_vip_interface)
      case 16: PB_U32(dummy32)  // is_vip
      case 20: PB_U32(o.real_ip)
      case 21: PB_U32(o.real_port)
      case 22: PB_U32(o.gpid)
      default: PB_SKIP()
    }
  }
  (void)dummy32;
  (void)dummy64;
  return true;
}

__device__ inline bool dec_tunnel(Rd r, L4Rec& o) {
  uint32_t f, wt, dummy32;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_U32(o.tx_ip0)
      case 2: PB_U32(o.tx_ip1)
      case 3: PB_U32(o.rx_ip0)
      case 4: PB_U32(o.rx_ip1)
      case 5: PB_U32(o.tx_mac0)
      case 6: PB_U32(o.tx_mac1)
      case 7: PB_U32(o.rx_mac0)
      case 8: PB_U32(o.rx_mac1)
      case 9: PB_U32(o.tx_id)
      case 10: PB_U32(o.rx_id)
      case 11: PB_U32(o.tunnel_type)
      case 12: PB_U32(o.tier)
      case 13: PB_U32(dummy32)  // is_ipv6
      default: PB_SKIP()
    }
  }
  (void)dummy32;
  return true;
}

__device__ inline bool dec_counts(Rd r, uint32_t& retrans, uint32_t& zero_win, uint32_t& ooo) {
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_U32(retrans)
      case 2: PB_U32(zero_win)
      case 3: PB_U32(ooo)
      default: PB_SKIP()
    }
  }
  return true;
}

__device__ inline bool dec_tcp(Rd r, L4Rec& o) {
  uint32_t f, wt, dummy32;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_U32(o.rtt_client_max)
      case 2: PB_U32(o.rtt_server_max)
      case 3: PB_U32(o.srt_max)
      case 4: PB_U32(o.art_max)
      case 5: PB_U32(o.rtt)
      case 8: PB_U32(o.srt_sum)
      case 9: PB_U32(o.art_sum)
      case 12: PB_U32(o.srt_count)
      case 13: PB_U32(o.art_count)
      case 14:
      case 15: {
        uint32_t n;
        if (wt != 2 || !rd_len(r, n)) return false;
        Rd s{r.p + r.i, 0, n};
        bool ok = f == 14 ? dec_counts(s, o.tx_retrans, o.tx_zero_win, o.tx_ooo)
                          : dec_counts(s, o.rx_retrans, o.rx_zero_win, o.rx_ooo);
        if (!ok) return false;
        r.i += n;
        break;
      }
      case 16: PB_U32(dummy32)  // total_retrans_count
      case 17: PB_U32(o.syn_count)
      case 18: PB_U32(o.synack_count)
      case 19: PB_U32(o.cit_max)
      case 20: PB_U32(o.cit_sum)
      case 21: PB_U32(o.cit_count)
      case 22: PB_U32(o.fin_count)
      default: PB_SKIP()
    }
  }
  (void)dummy32;
  return true;
}

__device__ inline bool dec_l7perf(Rd r, L4Rec& o) {
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: PB_U32(o.request_count)
      case 2: PB_U32(o.response_count)
      case 3: PB_U32(o.err_client_count)
      case 4: PB_U32(o.err_server_count)
      case 5: PB_U32(o.err_timeout)
      case 6: PB_U32(o.rrt_count)
      case 7: PB_U64(o.rrt_sum)
      case 8: PB_U32(o.rrt_max)
      case 9: PB_U32(o.tls_rtt)
      default: PB_SKIP()
    }
  }
  return true;
}

__device__ inline bool dec_perf(Rd r, L4Rec& o) {
  uint32_t f, wt, dummy32;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1:
      case 2: {
        uint32_t n;
        if (wt != 2 || !rd_len(r, n)) return false;
        Rd s{r.p + r.i, 0, n};
        if (!(f == 1 ? dec_tcp(s, o) : dec_l7perf(s, o))) return false;
        r.i += n;
        break;
      }
      case 3: PB_U32(dummy32)  // l4_protocol
      case 4: PB_U32(o.l7_protocol)
      case 5: PB_U32(o.l7_failed_count)
      default: PB_SKIP()
    }
  }
  (void)dummy32;
  return true;
}

// Repeated scalar sinks: the same walk counts (decode) or emits (pack pass).
struct NoSink {
  __device__ __forceinline__ void acl(uint32_t) {}
  __device__ __forceinline__ void agg(uint64_t) {}
};

template <class Sink>
__device__ inline bool dec_flow(Rd r, const uint8_t* base, L4Rec& o, Sink& sink) {
  uint32_t f, wt, dummy32;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1:
      case 2:
      case 3:
      case 4:
      case 13: {
        uint32_t n;
        if (wt != 2 || !rd_len(r, n)) return false;
        Rd s{r.p + r.i, 0, n};
        bool ok;
        if (f == 1) ok = dec_flow_key(s, base, o);
        else if (f == 2) ok = dec_peer(s, o.peer[0]);
        else if (f == 3) ok = dec_peer(s, o.peer[1]);
        else if (f == 4) ok = dec_tunnel(s, o);
        else ok = dec_perf(s, o);
        if (!ok) return false;
        r.i += n;
        break;
      }
      case 5: PB_U64(o.flow_id)
      case 6: PB_U64(o.start_time)
      case 7: PB_U64(o.end_time)
      case 8: PB_U64(o.duration)
      case 10: PB_U32(o.vlan)
      case 11: PB_U32(o.eth_type)
      case 12: PB_U32(o.has_perf_stats)
      case 14: PB_U32(o.close_type)
      case 15: PB_U32(o.signal_source)
      case 16: PB_U32(dummy32)  // is_active_service
      case 17: PB_U32(dummy32)  // queue_hash
      case 18: PB_U32(o.is_new_flow)
      case 19: PB_U32(o.tap_side)
      case 20: PB_U32(o.syn_seq)
      case 21: PB_U32(o.synack_seq)
      case 22: PB_U32(o.last_keepalive_seq)
      case 23: PB_U32(o.last_keepalive_ack)
      case 24:    // repeated uint32 acl_gids
      case 27: {  // repeated uint64 aggregated_flow_ids
        if (wt == 0) {
          uint64_t v;
          if (!rd_varint(r, v)) return false;
          if (f == 24) { o.n_acl_gids++; sink.acl((uint32_t)v); }
          else { o.agg_ids_chars += dec_digits(v) + (o.n_agg_ids ? 1 : 0); o.n_agg_ids++; sink.agg(v); }
        } else if (wt == 2) {
          uint32_t n;
          if (!rd_len(r, n)) return false;
          uint32_t end = r.i + n;
          while (r.i < end) {
            uint64_t v;
            if (!rd_varint(r, v)) return false;  // bounded by the message, like gogo
            if (f == 24) { o.n_acl_gids++; sink.acl((uint32_t)v); }
            else { o.agg_ids_chars += dec_digits(v) + (o.n_agg_ids ? 1 : 0); o.n_agg_ids++; sink.agg(v); }
          }
        } else {
          return false;
        }
        break;
      }
      case 25: PB_U32(o.direction_score)
      case 26: PB_STR(o.domain_off, o.domain_len)
      case 28: PB_U32(o.init_ipid)
      default: PB_SKIP()
    }
  }
  (void)dummy32;
  return true;
}

template <class Sink>
__device__ inline bool dec_tagged_flow(const uint8_t* rec, uint32_t len, L4Rec& o, Sink& sink) {
  Rd r{rec, 0, len};
  const uint8_t* base = rec;
  uint32_t f, wt;
  while (r.i < r.l) {
    if (!rd_tag(r, f, wt)) return false;
    switch (f) {
      case 1: {
        uint32_t n;
        if (wt != 2 || !rd_len(r, n)) return false;
        Rd s{r.p + r.i, 0, n};
        if (!dec_flow(s, base, o, sink)) return false;
        r.i += n;
        break;
      }
      default: PB_SKIP()
    }
  }
  return true;
}

}  // namespace dfgpu
