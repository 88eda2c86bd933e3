/*
 * ORACLE — test infrastructure only.
 * l4.cc: CPU restatement of the L4 flow-log row path
 *   TaggedFlowToL4FlowLog   server/ingester/flow_log/log_data/l4_flow_log.go:955-971
 *   *.Fill                  l4_flow_log.go:467-534, 730-743, 779-884
 *   getStatus               l4_flow_log.go:745-760 ; CloseType datatype/flow.go:33-84
 *   genID                   l4_flow_log.go:949-953
 *   geo.QueryProvince       flow_log/geo/geo.go:29 ; libs/geo/netmask_tree.go:24-104
 *   AppendToColumnBlock     l4_column_block.go:53-59,145-172,213-223,243-246,272-276,
 *                           410-451,561-593,748-796,848-859 ; ToInput order :822-833
 * Pinned: the TaggedFlow literal (tagged_flow_test.go:29-103), the column list (machine-extracted from ToInput), the
 * netmask tree.  Parity unpinned by the reference: the Fill functions, getStatus, genID, AppendToColumnBlock —
 * cross-checked by tests/test_l4_row_crosscheck_cpu.py.
 */
#include "l4.h"

#include <cstdio>

namespace oracle {

// ---- geo: netmask tree (libs/geo/netmask_tree.go) ------------------------------
static uint32_t maskLenToNetmask(int i) { return i == 0 ? 0 : 0xFFFFFFFFu << (32 - i); }

int getFirstMask(uint32_t start, uint32_t end) {  // netmask_tree.go:37-50
  const int MIN_MASKLEN = 16, MAX_MASKLEN = 32;
  int maxLen = MAX_MASKLEN;
  for (; maxLen > MIN_MASKLEN; maxLen--) {
    if (start & (1u << (uint32_t)(MAX_MASKLEN - maxLen))) break;
    if ((uint32_t)(start + ~maskLenToNetmask(maxLen)) >= end ||
        (uint32_t)(start + ~maskLenToNetmask(maxLen - 1)) > end)
      break;
  }
  return maxLen;
}

void GeoTree::build(const dfgpu_geo_entry* e, uint64_t n, const char* const* names, uint32_t n_names) {
  region_names.clear();
  for (uint32_t i = 0; i < n_names; i++) region_names.emplace_back(names[i]);
  maxMask.assign(1 << 16, 0);
  cache.clear();
  const int MIN_MASKLEN = 16, MAX_MASKLEN = 32;
  for (uint64_t i = 0; i < n; i++) {  // updateMask :52-62
    uint32_t start = e[i].ip_start, end = e[i].ip_end;
    while (end >= start) {
      int maskLen = getFirstMask(start, end);
      uint32_t key = start >> (MAX_MASKLEN - MIN_MASKLEN);
      if (maskLen > maxMask[key]) maxMask[key] = (uint8_t)maskLen;
      uint32_t step = 1u << (uint32_t)(MAX_MASKLEN - maskLen);
      uint32_t next = start + step;
      if (next < start) break;  // uint32 wrap: Go would loop forever at 255.255.255.255; guard
      start = next;
    }
  }
  for (uint64_t i = 0; i < n; i++) {  // :78-86
    for (uint32_t ip = e[i].ip_start; ip <= e[i].ip_end && ip > 0;) {
      int mask = maxMask[ip >> 16];
      uint32_t key = ip & maskLenToNetmask(mask);
      cache[key] = (uint16_t)((uint16_t)e[i].region << 8);
      uint32_t next = ip + (1u << (32 - (uint32_t)mask));
      if (next < ip) break;
      ip = next;
    }
  }
}

uint8_t GeoTree::query_region(uint32_t ip) const {  // Query :97-104
  if (maxMask.empty()) return 0;
  int mm = maxMask[ip >> 16];
  uint32_t key = ip & maskLenToNetmask(mm);
  auto it = cache.find(key);
  if (it == cache.end()) return 0;
  return (uint8_t)(it->second >> 8);
}
std::string_view GeoTree::province(uint32_t ip) const {  // DecodeRegion utils.go:27-40
  uint8_t r = query_region(ip);
  if (r >= region_names.size()) return unknown;
  return region_names[r];
}

// ---- CloseType (datatype/flow.go:33-84) -----------------------------------------
enum {
  CloseTypeFinish = 1, CloseTypeTCPServerRst = 2, CloseTypeTimeout = 3, CloseTypeForcedReport = 5,
  CloseTypeServerSynMiss = 7, CloseTypeServerHalfClose = 8, CloseTypeTCPClientRst = 9,
  CloseTypeClientAckMiss = 10, CloseTypeClientHalfClose = 11, CloseTypeClientSourcePortReuse = 13,
  CloseTypeServerReset = 15, CloseTypeServerQueueLack = 17, CloseTypeClientEstablishReset = 18,
  CloseTypeServerEstablishReset = 19, CloseTypeTCPFinClientRst = 20
};
static uint8_t getStatus(uint8_t t, uint8_t p) {  // l4_flow_log.go:745-760; CloseType is uint8
  bool tcp = p == 6;
  if (t == CloseTypeFinish || t == CloseTypeForcedReport || t == CloseTypeTCPFinClientRst ||
      (!tcp && t == CloseTypeTimeout) || t == CloseTypeClientHalfClose || t == CloseTypeServerHalfClose)
    return 0;  // STATUS_OK
  if (t == CloseTypeClientAckMiss || t == CloseTypeTCPClientRst || t == CloseTypeClientSourcePortReuse ||
      t == CloseTypeClientEstablishReset)
    return 4;  // STATUS_CLIENT_ERROR
  if (tcp && (t == CloseTypeTCPServerRst || t == CloseTypeTimeout || t == CloseTypeServerSynMiss ||
              t == CloseTypeServerReset || t == CloseTypeServerQueueLack ||
              t == CloseTypeServerEstablishReset))
    return 3;  // STATUS_SERVER_ERROR
  return 2;    // STATUS_TIMEOUT
}

uint64_t genID(uint32_t time, uint32_t* counter, uint32_t analyzerID) {  // :949-953
  uint32_t count = ++(*counter);
  return (uint64_t)time << 32 | (uint64_t)(analyzerID & 0x3ff) << 22 | ((uint64_t)count & 0x3fffff);
}

// ---- column block ------------------------------------------------------------------
void L4Block::init() {
  auto U8 = [&](const char* n) { b.add(n, DFGPU_CH_UINT8, 1); };
  auto U16 = [&](const char* n) { b.add(n, DFGPU_CH_UINT16, 2); };
  auto U32 = [&](const char* n) { b.add(n, DFGPU_CH_UINT32, 4); };
  auto U64 = [&](const char* n) { b.add(n, DFGPU_CH_UINT64, 8); };
  auto F64 = [&](const char* n) { b.add(n, DFGPU_CH_FLOAT64, 8); };
  auto IP4 = [&](const char* n) { b.add(n, DFGPU_CH_IPV4, 4); };
  auto IP6c = [&](const char* n) { b.add(n, DFGPU_CH_IPV6, 16); };
  U64("_id");
  // DataLinkLayerBlock :40
  U64("mac_0"); U64("mac_1"); U16("eth_type"); U16("vlan");
  // NetworkLayerBlock :113
  IP4("ip4_0"); IP4("ip4_1"); IP6c("ip6_0"); IP6c("ip6_1"); U8("is_ipv4"); U8("protocol");
  U8("tunnel_tier"); U16("tunnel_type"); U32("tunnel_tx_id"); U32("tunnel_rx_id");
  IP4("tunnel_tx_ip4_0"); IP4("tunnel_tx_ip4_1"); IP4("tunnel_rx_ip4_0"); IP4("tunnel_rx_ip4_1");
  IP6c("tunnel_tx_ip6_0"); IP6c("tunnel_tx_ip6_1"); IP6c("tunnel_rx_ip6_0"); IP6c("tunnel_rx_ip6_1");
  U8("tunnel_is_ipv4"); U32("tunnel_tx_mac_0"); U32("tunnel_tx_mac_1"); U32("tunnel_rx_mac_0");
  U32("tunnel_rx_mac_1");
  // TransportLayerBlock :196
  U16("client_port"); U16("server_port"); U16("tcp_flags_bit_0"); U16("tcp_flags_bit_1");
  U32("syn_seq"); U32("syn_ack_seq"); U32("last_keepalive_seq"); U32("last_keepalive_ack");
  // ApplicationLayerBlock :233
  U8("l7_protocol");
  // InternetBlock :258
  b.add("province_0", DFGPU_CH_LC_STR, 4); b.add("province_1", DFGPU_CH_LC_STR, 4);
  // KnowledgeGraphBlock :362
  kg_base = b.cols.size();
  KnowledgeGraph::add_columns(b);
  // FlowInfoBlock :519
  fi_base = b.cols.size();
  b.add("time", DFGPU_CH_DATETIME, 4); U16("close_type"); U16("signal_source"); U64("flow_id");
  b.add("aggregated_flow_ids", DFGPU_CH_STR); U32("init_ipid"); U8("capture_network_type_id");
  U8("nat_source"); U8("capture_nic_type"); U32("capture_nic");
  b.add("observation_point", DFGPU_CH_LC_STR, 4); U16("agent_id"); U8("l2_end_0"); U8("l2_end_1");
  U8("l3_end_0"); U8("l3_end_1"); b.add("start_time", DFGPU_CH_DATETIME64, 8);
  b.add("end_time", DFGPU_CH_DATETIME64, 8); U64("duration"); U8("is_new_flow"); U8("status");
  b.add("acl_gids", DFGPU_CH_ARR_UINT16, 2); U32("gprocess_id_0"); U32("gprocess_id_1");
  IP4("nat_real_ip4_0"); IP4("nat_real_ip4_1"); U16("nat_real_port_0"); U16("nat_real_port_1");
  U8("direction_score"); b.add("request_domain", DFGPU_CH_STR);
  // MetricsBlock :693
  m_base = b.cols.size();
  U64("packet_tx"); U64("packet_rx"); U64("byte_tx"); U64("byte_rx"); U64("l3_byte_tx");
  U64("l3_byte_rx"); U64("l4_byte_tx"); U64("l4_byte_rx"); U64("total_packet_tx");
  U64("total_packet_rx"); U64("total_byte_tx"); U64("total_byte_rx"); U32("l7_request");
  U32("l7_response"); U32("l7_parse_failed"); F64("rtt"); F64("rtt_client"); F64("rtt_server");
  F64("tls_rtt"); F64("srt_sum"); F64("art_sum"); F64("rrt_sum"); F64("cit_sum"); U64("srt_count");
  U64("art_count"); U64("rrt_count"); U64("cit_count"); U32("srt_max"); U32("art_max");
  U32("rrt_max"); U32("cit_max"); U32("retrans_tx"); U32("retrans_rx"); U32("zero_win_tx");
  U32("zero_win_rx"); U32("syn_count"); U32("synack_count"); U32("retrans_syn");
  U32("retrans_synack"); U32("l7_client_error"); U32("l7_server_error"); U32("l7_server_timeout");
  U32("l7_error"); U32("ooo_rx"); U32("ooo_tx"); U32("fin_count");
}

// One TaggedFlow record -> one appended row.  `tf` is the decoded pb (with the
// pre-allocated sub-messages of pb.NewTaggedFlow, datatype/pb/flow_log.go:103-119).
void L4Block::append_record(const Msg& tf, uint16_t orgId, uint16_t teamId,
                            const PlatformInfoTable* pd, const GeoTree* geo, uint32_t* id_counter) {
  using namespace fl;
  const Msg& f = tf.sub(TaggedFlow::flow);
  const Msg& k = f.sub(Flow::flow_key);
  const Msg& s = f.sub(Flow::metrics_peer_src);
  const Msg& d = f.sub(Flow::metrics_peer_dst);
  const Msg& t = f.sub(Flow::tunnel);
  const Msg& p = f.sub(Flow::perf_stats);
  const Msg& tcp = p.sub(PerfStats::tcp);
  const Msg& l7 = p.sub(PerfStats::l7);
  const Msg& ctx = tcp.sub(Tcp::counts_peer_tx);
  const Msg& crx = tcp.sub(Tcp::counts_peer_rx);

  bool isIPV6 = f.u32(Flow::eth_type) == 0x86DD;  // :956
  size_t c = 0;
  // _id :960
  b[c++].append<uint64_t>(genID((uint32_t)(f.u64(Flow::end_time) / 1000000000ull), id_counter, pd->QueryAnalyzerID()));
  // DataLinkLayer.Fill :467-472
  b[c++].append<uint64_t>(k.u64(FlowKey::mac_src));
  b[c++].append<uint64_t>(k.u64(FlowKey::mac_dst));
  b[c++].append<uint16_t>((uint16_t)f.u32(Flow::eth_type));
  b[c++].append<uint16_t>((uint16_t)f.u32(Flow::vlan));
  // NetworkLayer.Fill :484-512
  std::string_view ip60, ip61;
  uint32_t ip40 = 0, ip41 = 0;
  if (isIPV6) {
    ip60 = k.str(FlowKey::ip6_src);
    ip61 = k.str(FlowKey::ip6_dst);
  } else {
    ip40 = k.u32(FlowKey::ip_src);
    ip41 = k.u32(FlowKey::ip_dst);
  }
  b[c++].append<uint32_t>(ip40);
  b[c++].append<uint32_t>(ip41);
  b[c++].append_ip6((const uint8_t*)ip60.data(), ip60.size());
  b[c++].append_ip6((const uint8_t*)ip61.data(), ip61.size());
  b[c++].append<uint8_t>(isIPV6 ? 0 : 1);
  b[c++].append<uint8_t>((uint8_t)k.u32(FlowKey::proto));
  bool tun = t.u32(Tunnel::tunnel_type) != 0;
  b[c++].append<uint8_t>(tun ? (uint8_t)t.u32(Tunnel::tier) : 0);
  b[c++].append<uint16_t>(tun ? (uint16_t)t.u32(Tunnel::tunnel_type) : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::tx_id) : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::rx_id) : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::tx_ip0) : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::tx_ip1) : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::rx_ip0) : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::rx_ip1) : 0);
  for (int i = 0; i < 4; i++) b[c++].append_ip6(nullptr, 0);  // tunnel ip6 never filled
  b[c++].append<uint8_t>(tun ? 1 : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::tx_mac0) : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::tx_mac1) : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::rx_mac0) : 0);
  b[c++].append<uint32_t>(tun ? t.u32(Tunnel::rx_mac1) : 0);
  // TransportLayer.Fill :514-523
  b[c++].append<uint16_t>((uint16_t)k.u32(FlowKey::port_src));
  b[c++].append<uint16_t>((uint16_t)k.u32(FlowKey::port_dst));
  b[c++].append<uint16_t>((uint16_t)s.u32(Peer::tcp_flags));
  b[c++].append<uint16_t>((uint16_t)d.u32(Peer::tcp_flags));
  b[c++].append<uint32_t>(f.u32(Flow::syn_seq));
  b[c++].append<uint32_t>(f.u32(Flow::synack_seq));
  b[c++].append<uint32_t>(f.u32(Flow::last_keepalive_seq));
  b[c++].append<uint32_t>(f.u32(Flow::last_keepalive_ack));
  // ApplicationLayer.Fill :525-529
  bool perf = f.u32(Flow::has_perf_stats) == 1;
  b[c++].append<uint8_t>(perf ? (uint8_t)p.u32(PerfStats::l7_protocol) : 0);
  // Internet.Fill :531-534 (always on the v4 fields, even for v6 flows)
  b[c++].append_lc(geo->province(k.u32(FlowKey::ip_src)));
  b[c++].append_lc(geo->province(k.u32(FlowKey::ip_dst)));
  // KnowledgeGraph.FillL4 :730-743
  KnowledgeGraph kg;
  kg.OrgId = orgId;
  kg.TeamID = teamId;
  kg.fill(pd, isIPV6, s.u32(Peer::is_vip_interface) == 1, d.u32(Peer::is_vip_interface) == 1,
          MarshalInt32WithSpecialID(s.i32(Peer::l3_epc_id)), MarshalInt32WithSpecialID(d.i32(Peer::l3_epc_id)),
          k.u32(FlowKey::ip_src), k.u32(FlowKey::ip_dst), k.str(FlowKey::ip6_src), k.str(FlowKey::ip6_dst),
          k.u64(FlowKey::mac_src), k.u64(FlowKey::mac_dst), s.u32(Peer::gpid), d.u32(Peer::gpid),
          (uint16_t)k.u32(FlowKey::vtap_id), 0, 0, (uint16_t)k.u32(FlowKey::port_dst),
          f.u32(Flow::tap_side), (uint8_t)k.u32(FlowKey::proto));
  kg.append(b, c);
  c += 39;
  // FlowInfo.Fill :779-815
  uint64_t tapPort = k.u64(FlowKey::tap_port);
  uint16_t closeType = (uint16_t)f.u32(Flow::close_type);
  b[c++].append<uint32_t>((uint32_t)(f.u64(Flow::end_time) / 1000000000ull));  // time
  b[c++].append<uint16_t>(closeType);
  b[c++].append<uint16_t>((uint16_t)f.u32(Flow::signal_source));
  b[c++].append<uint64_t>(f.u64(Flow::flow_id));
  {  // Uint64SliceToString :762-777
    std::string sj;
    const auto& ids = f.ru[Flow::aggregated_flow_ids];
    char tmp[24];
    for (size_t i = 0; i < ids.size(); i++) {
      if (i) sj.push_back(',');
      snprintf(tmp, sizeof tmp, "%llu", (unsigned long long)ids[i]);
      sj += tmp;
    }
    b[c++].append_str(sj);
  }
  b[c++].append<uint32_t>(f.u32(Flow::init_ipid));
  b[c++].append<uint8_t>((uint8_t)k.u32(FlowKey::tap_type));
  b[c++].append<uint8_t>((uint8_t)((tapPort >> 36) & 0xf));  // nat_source
  b[c++].append<uint8_t>((uint8_t)(tapPort >> 60));          // capture_nic_type
  b[c++].append<uint32_t>((uint32_t)tapPort);                // capture_nic
  bool ok;
  b[c++].append_lc(TAPSideString(f.u32(Flow::tap_side) & 0xff, ok));  // TAPSideEnum is uint8
  b[c++].append<uint16_t>((uint16_t)k.u32(FlowKey::vtap_id));
  b[c++].append<uint8_t>(s.u32(Peer::is_l2_end) == 1);
  b[c++].append<uint8_t>(d.u32(Peer::is_l2_end) == 1);
  b[c++].append<uint8_t>(s.u32(Peer::is_l3_end) == 1);
  b[c++].append<uint8_t>(d.u32(Peer::is_l3_end) == 1);
  b[c++].append<int64_t>((int64_t)f.u64(Flow::start_time) / 1000);
  b[c++].append<int64_t>((int64_t)f.u64(Flow::end_time) / 1000);
  b[c++].append<uint64_t>(f.u64(Flow::duration) / 1000);
  b[c++].append<uint8_t>((uint8_t)f.u32(Flow::is_new_flow));
  b[c++].append<uint8_t>(getStatus((uint8_t)closeType, (uint8_t)k.u32(FlowKey::proto)));
  {
    std::vector<uint16_t> gids;
    for (auto v : f.ru[Flow::acl_gids]) gids.push_back((uint16_t)v);
    b[c++].append_arr<uint16_t>(gids);
  }
  b[c++].append<uint32_t>(s.u32(Peer::gpid));
  b[c++].append<uint32_t>(d.u32(Peer::gpid));
  b[c++].append<uint32_t>(s.u32(Peer::real_ip));
  b[c++].append<uint32_t>(d.u32(Peer::real_ip));
  b[c++].append<uint16_t>((uint16_t)s.u32(Peer::real_port));
  b[c++].append<uint16_t>((uint16_t)d.u32(Peer::real_port));
  b[c++].append<uint8_t>((uint8_t)f.u32(Flow::direction_score));
  b[c++].append_str(f.str(Flow::request_domain));
  // Metrics.Fill :817-884
  b[c++].append<uint64_t>(s.u64(Peer::packet_count));
  b[c++].append<uint64_t>(d.u64(Peer::packet_count));
  b[c++].append<uint64_t>(s.u64(Peer::byte_count));
  b[c++].append<uint64_t>(d.u64(Peer::byte_count));
  b[c++].append<uint64_t>(s.u64(Peer::l3_byte_count));
  b[c++].append<uint64_t>(d.u64(Peer::l3_byte_count));
  b[c++].append<uint64_t>(s.u64(Peer::l4_byte_count));
  b[c++].append<uint64_t>(d.u64(Peer::l4_byte_count));
  b[c++].append<uint64_t>(s.u64(Peer::total_packet_count));
  b[c++].append<uint64_t>(d.u64(Peer::total_packet_count));
  b[c++].append<uint64_t>(s.u64(Peer::total_byte_count));
  b[c++].append<uint64_t>(d.u64(Peer::total_byte_count));
  uint32_t L7Request = 0, L7Response = 0, L7ParseFailed = 0, RTT = 0, RTTClient = 0, RTTServer = 0,
           TLSRTT = 0, SRTSum = 0, ARTSum = 0, CITSum = 0, SRTCount = 0, ARTCount = 0, RRTCount = 0,
           CITCount = 0, SRTMax = 0, ARTMax = 0, RRTMax = 0, CITMax = 0, RetransTx = 0, RetransRx = 0,
           ZeroWinTx = 0, ZeroWinRx = 0, SynCount = 0, SynackCount = 0, RetransSyn = 0,
           RetransSynack = 0, L7ClientError = 0, L7ServerError = 0, L7ServerTimeout = 0, L7Error = 0,
           OooTx = 0, OooRx = 0, FinCount = 0;
  uint64_t RRTSum = 0;
  if (perf) {
    L7Request = l7.u32(L7Perf::request_count);
    L7Response = l7.u32(L7Perf::response_count);
    L7ClientError = l7.u32(L7Perf::err_client_count);
    L7ServerError = l7.u32(L7Perf::err_server_count);
    L7ServerTimeout = l7.u32(L7Perf::err_timeout);
    L7Error = L7ClientError + L7ServerError;
    L7ParseFailed = p.u32(PerfStats::l7_failed_count);
    RTT = tcp.u32(Tcp::rtt);
    RTTClient = tcp.u32(Tcp::rtt_client_max);
    RTTServer = tcp.u32(Tcp::rtt_server_max);
    TLSRTT = l7.u32(L7Perf::tls_rtt);
    SRTSum = tcp.u32(Tcp::srt_sum);
    SRTCount = tcp.u32(Tcp::srt_count);
    ARTSum = tcp.u32(Tcp::art_sum);
    ARTCount = tcp.u32(Tcp::art_count);
    RRTSum = l7.u64(L7Perf::rrt_sum);
    RRTCount = l7.u32(L7Perf::rrt_count);
    CITSum = tcp.u32(Tcp::cit_sum);
    CITCount = tcp.u32(Tcp::cit_count);
    SRTMax = tcp.u32(Tcp::srt_max);
    ARTMax = tcp.u32(Tcp::art_max);
    RRTMax = l7.u32(L7Perf::rrt_max);
    CITMax = tcp.u32(Tcp::cit_max);
    RetransTx = ctx.u32(Counts::retrans_count);  // CountsPeerTx pre-allocated: never nil
    ZeroWinTx = ctx.u32(Counts::zero_win_count);
    OooTx = ctx.u32(Counts::ooo_count);
    RetransRx = crx.u32(Counts::retrans_count);
    ZeroWinRx = crx.u32(Counts::zero_win_count);
    OooRx = crx.u32(Counts::ooo_count);
    SynCount = tcp.u32(Tcp::syn_count);
    SynackCount = tcp.u32(Tcp::synack_count);
    FinCount = tcp.u32(Tcp::fin_count);
    if (SynCount > 0) RetransSyn = SynCount - 1;
    if (SynackCount > 0) RetransSynack = SynackCount - 1;
  }
  b[c++].append<uint32_t>(L7Request);
  b[c++].append<uint32_t>(L7Response);
  b[c++].append<uint32_t>(L7ParseFailed);
  b[c++].append<double>((double)RTT);
  b[c++].append<double>((double)RTTClient);
  b[c++].append<double>((double)RTTServer);
  b[c++].append<double>((double)TLSRTT);
  b[c++].append<double>((double)SRTSum);
  b[c++].append<double>((double)ARTSum);
  b[c++].append<double>((double)RRTSum);  // uint64 -> float64, round-to-nearest-even
  b[c++].append<double>((double)CITSum);
  b[c++].append<uint64_t>(SRTCount);
  b[c++].append<uint64_t>(ARTCount);
  b[c++].append<uint64_t>(RRTCount);
  b[c++].append<uint64_t>(CITCount);
  b[c++].append<uint32_t>(SRTMax);
  b[c++].append<uint32_t>(ARTMax);
  b[c++].append<uint32_t>(RRTMax);
  b[c++].append<uint32_t>(CITMax);
  b[c++].append<uint32_t>(RetransTx);
  b[c++].append<uint32_t>(RetransRx);
  b[c++].append<uint32_t>(ZeroWinTx);
  b[c++].append<uint32_t>(ZeroWinRx);
  b[c++].append<uint32_t>(SynCount);
  b[c++].append<uint32_t>(SynackCount);
  b[c++].append<uint32_t>(RetransSyn);
  b[c++].append<uint32_t>(RetransSynack);
  b[c++].append<uint32_t>(L7ClientError);
  b[c++].append<uint32_t>(L7ServerError);
  b[c++].append<uint32_t>(L7ServerTimeout);
  b[c++].append<uint32_t>(L7Error);
  b[c++].append<uint32_t>(OooRx);
  b[c++].append<uint32_t>(OooTx);
  b[c++].append<uint32_t>(FinCount);
  b.nrows++;
}

}  // namespace oracle
