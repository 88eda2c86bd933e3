"""The l4_flow_log ROW a second time (SURVEY A4 + A8): TaggedFlowToL4FlowLog, DataLinkLayer / NetworkLayer /
TransportLayer / ApplicationLayer / FlowInfo / Metrics .Fill, getStatus, genID, Uint64SliceToString
(flow_log/log_data/l4_flow_log.go:457-953) and the blocks' AppendToColumnBlock (l4_column_block.go) restated in
Python from the Go source, fed by Google's protobuf runtime, against every column of the oracle's rows except the
KnowledgeGraph ones (second opinion: test_kg_crosscheck_cpu.py); province_0/1 come from a second restatement of the
netmask geo tree (libs/geo/netmask_tree.go) over the same range table."""
import numpy as np
import pytest

import pb_schema as P
from deepflow_b200 import abi
from deepflow_b200 import synth as S
from test_l7_row_crosscheck_cpu import KG_PREFIX, TAPSIDE, _records

CLIENT_ERR = {10, 9, 13, 18}        # ClientAckMiss, TCPClientRst, ClientSourcePortReuse, ClientEstablishReset
SERVER_ERR = {2, 3, 7, 15, 17, 19}  # TCPServerRst, Timeout, ServerSynMiss, ServerReset, ServerQueueLack, ServerEstablishReset


def get_status(close_type, proto):  # l4_flow_log.go:747-762
    if close_type in (1, 5, 20) or (proto != 6 and close_type == 3) or close_type in (11, 8):
        return 0  # STATUS_OK
    if close_type in CLIENT_ERR:
        return 4  # STATUS_CLIENT_ERROR
    if proto == 6 and close_type in SERVER_ERR:
        return 3  # STATUS_SERVER_ERROR
    return 2      # STATUS_TIMEOUT


def _ip6(b):
    return (bytes(b) + bytes(16))[:16]


def l4_row(m, counter, analyzer_id):
    f = m.flow
    k, a, b, t = f.flow_key, f.metrics_peer_src, f.metrics_peer_dst, f.Tunnel
    v6 = f.eth_type == 0x86DD
    o = {}
    time_s = (f.end_time // 1_000_000_000) & 0xFFFFFFFF
    o["_id"] = (time_s << 32) | ((analyzer_id & 0x3FF) << 22) | (counter & 0x3FFFFF)
    o["mac_0"], o["mac_1"], o["eth_type"], o["vlan"] = k.mac_src, k.mac_dst, f.eth_type & 0xFFFF, f.vlan & 0xFFFF
    o["ip4_0"], o["ip4_1"] = (0, 0) if v6 else (k.ip_src, k.ip_dst)
    o["ip6_0"], o["ip6_1"] = (_ip6(k.ip6_src), _ip6(k.ip6_dst)) if v6 else (bytes(16), bytes(16))
    o["is_ipv4"], o["protocol"] = 0 if v6 else 1, k.proto & 0xFF
    tun = t.tunnel_type != 0
    o["tunnel_tier"] = (t.tier & 0xFF) if tun else 0
    o["tunnel_type"] = (t.tunnel_type & 0xFFFF) if tun else 0
    o["tunnel_tx_id"], o["tunnel_rx_id"] = (t.tx_id, t.rx_id) if tun else (0, 0)
    o["tunnel_tx_ip4_0"], o["tunnel_tx_ip4_1"] = (t.tx_ip0, t.tx_ip1) if tun else (0, 0)
    o["tunnel_rx_ip4_0"], o["tunnel_rx_ip4_1"] = (t.rx_ip0, t.rx_ip1) if tun else (0, 0)
    for c in ("tunnel_tx_ip6_0", "tunnel_tx_ip6_1", "tunnel_rx_ip6_0", "tunnel_rx_ip6_1"):
        o[c] = bytes(16)
    o["tunnel_is_ipv4"] = 1 if tun else 0
    o["tunnel_tx_mac_0"], o["tunnel_tx_mac_1"] = (t.tx_mac0, t.tx_mac1) if tun else (0, 0)
    o["tunnel_rx_mac_0"], o["tunnel_rx_mac_1"] = (t.rx_mac0, t.rx_mac1) if tun else (0, 0)
    o["client_port"], o["server_port"] = k.port_src & 0xFFFF, k.port_dst & 0xFFFF
    o["tcp_flags_bit_0"], o["tcp_flags_bit_1"] = a.tcp_flags & 0xFFFF, b.tcp_flags & 0xFFFF
    o["syn_seq"], o["syn_ack_seq"] = f.syn_seq, f.synack_seq
    o["last_keepalive_seq"], o["last_keepalive_ack"] = f.last_keepalive_seq, f.last_keepalive_ack
    perf = f.has_perf_stats == 1
    p = f.perf_stats
    o["l7_protocol"] = (p.l7_protocol & 0xFF) if perf else 0
    o["time"] = time_s
    o["close_type"], o["signal_source"] = f.close_type & 0xFFFF, f.signal_source & 0xFFFF
    o["flow_id"] = f.flow_id
    o["aggregated_flow_ids"] = ",".join(str(x) for x in f.aggregated_flow_ids).encode()
    o["init_ipid"] = f.init_ipid
    o["capture_network_type_id"] = k.tap_type & 0xFF
    tp = k.tap_port
    o["capture_nic"], o["capture_nic_type"], o["nat_source"] = tp & 0xFFFFFFFF, (tp >> 60) & 0xFF, (tp >> 36) & 0xF
    o["observation_point"] = TAPSIDE[f.tap_side].encode()
    o["agent_id"] = k.vtap_id & 0xFFFF
    o["l2_end_0"], o["l2_end_1"] = int(a.is_l2_end == 1), int(b.is_l2_end == 1)
    o["l3_end_0"], o["l3_end_1"] = int(a.is_l3_end == 1), int(b.is_l3_end == 1)
    o["start_time"], o["end_time"], o["duration"] = f.start_time // 1000, f.end_time // 1000, f.duration // 1000
    o["is_new_flow"] = f.is_new_flow & 0xFF
    o["status"] = get_status(f.close_type & 0xFF, k.proto & 0xFF)  # datatype.CloseType is a uint8: CloseType(uint16) truncates
    o["acl_gids"] = [x & 0xFFFF for x in f.acl_gids]
    o["gprocess_id_0"], o["gprocess_id_1"] = a.gpid, b.gpid
    o["nat_real_ip4_0"], o["nat_real_ip4_1"] = a.real_ip, b.real_ip
    o["nat_real_port_0"], o["nat_real_port_1"] = a.real_port & 0xFFFF, b.real_port & 0xFFFF
    o["direction_score"], o["request_domain"] = f.direction_score & 0xFF, f.request_domain.encode()
    o["packet_tx"], o["packet_rx"], o["byte_tx"], o["byte_rx"] = a.packet_count, b.packet_count, a.byte_count, b.byte_count
    o["l3_byte_tx"], o["l3_byte_rx"], o["l4_byte_tx"], o["l4_byte_rx"] = a.l3_byte_count, b.l3_byte_count, a.l4_byte_count, b.l4_byte_count
    o["total_packet_tx"], o["total_packet_rx"] = a.total_packet_count, b.total_packet_count
    o["total_byte_tx"], o["total_byte_rx"] = a.total_byte_count, b.total_byte_count
    z = lambda v: v if perf else 0  # noqa: E731  (everything below is read only when has_perf_stats == 1)
    l7, tcp = p.l7, p.tcp
    o["l7_request"], o["l7_response"], o["l7_parse_failed"] = z(l7.request_count), z(l7.response_count), z(p.l7_failed_count)
    o["rtt"], o["rtt_client"], o["rtt_server"], o["tls_rtt"] = (float(z(x)) for x in (tcp.rtt, tcp.rtt_client_max, tcp.rtt_server_max, l7.tls_rtt))
    o["srt_sum"], o["art_sum"], o["rrt_sum"], o["cit_sum"] = (float(z(x)) for x in (tcp.srt_sum, tcp.art_sum, l7.rrt_sum, tcp.cit_sum))
    o["srt_count"], o["art_count"], o["rrt_count"], o["cit_count"] = z(tcp.srt_count), z(tcp.art_count), z(l7.rrt_count), z(tcp.cit_count)
    o["srt_max"], o["art_max"], o["rrt_max"], o["cit_max"] = z(tcp.srt_max), z(tcp.art_max), z(l7.rrt_max), z(tcp.cit_max)
    tx, rx = tcp.counts_peer_tx, tcp.counts_peer_rx  # absent sub-message reads as zeros, like the nil checks
    o["retrans_tx"], o["zero_win_tx"], o["ooo_tx"] = z(tx.retrans_count), z(tx.zero_win_count), z(tx.ooo_count)
    o["retrans_rx"], o["zero_win_rx"], o["ooo_rx"] = z(rx.retrans_count), z(rx.zero_win_count), z(rx.ooo_count)
    o["syn_count"], o["synack_count"], o["fin_count"] = z(tcp.syn_count), z(tcp.synack_count), z(tcp.fin_count)
    o["retrans_syn"] = o["syn_count"] - 1 if o["syn_count"] > 0 else 0
    o["retrans_synack"] = o["synack_count"] - 1 if o["synack_count"] > 0 else 0
    o["l7_client_error"], o["l7_server_error"], o["l7_server_timeout"] = z(l7.err_client_count), z(l7.err_server_count), z(l7.err_timeout)
    o["l7_error"] = (o["l7_client_error"] + o["l7_server_error"]) & 0xFFFFFFFF
    return o


class GeoTree:
    """geo.NewNetmaskGeoTree / netmaskTree.Query (libs/geo/netmask_tree.go:19-104) a second time."""
    MINL, MAXL = 16, 32
    NETMASK = [(~(0xFFFFFFFF >> i)) & 0xFFFFFFFF for i in range(33)]

    @classmethod
    def first_mask(cls, start, end):
        ml = cls.MAXL
        while ml > cls.MINL:
            if start & (1 << (cls.MAXL - ml)):
                break
            inv = lambda m: (~cls.NETMASK[m]) & 0xFFFFFFFF  # noqa: E731
            if ((start + inv(ml)) & 0xFFFFFFFF) >= end or ((start + inv(ml - 1)) & 0xFFFFFFFF) > end:
                break
            ml -= 1
        return ml

    def __init__(self, entries, names):
        self.max_mask = [0] * (1 << self.MINL)
        self.names = names
        for s, e, _r in entries:  # updateMask
            start = s
            while e >= start:
                ml = self.first_mask(start, e)
                key = start >> (self.MAXL - self.MINL)
                if ml > self.max_mask[key]:
                    self.max_mask[key] = ml
                start += 1 << (self.MAXL - ml)
                if start > 0xFFFFFFFF:
                    break
        self.cache = {}
        for s, e, r in entries:
            ip = s
            while ip <= e and ip > 0:
                m = self.max_mask[ip >> 16]
                self.cache[ip & self.NETMASK[m]] = r
                ip += 1 << (self.MAXL - m)
                if ip > 0xFFFFFFFF:
                    break

    def province(self, ip):
        m = self.max_mask[ip >> 16]
        r = self.cache.get(ip & self.NETMASK[m], 0)
        return self.names[r] if r < len(self.names) else "未知".encode()  # geo/utils.go:35-40


def _geo(synth):
    import ctypes as C
    e, n, names, nn = synth.geo()
    ent = C.cast(e, C.POINTER(abi.GeoEntry * n)).contents
    nm = C.cast(names, C.POINTER(C.c_char_p * nn)).contents
    return GeoTree([(x.ip_start, x.ip_end, x.region) for x in ent], [bytes(x) for x in nm])


@pytest.mark.parametrize("sizes,seed,flags", [([1500, 800], 5, 0), ([1200], 6, S.F_ALL_WIRE)])
def test_l4_rows_match_second_opinion(synth, built, sizes, seed, flags):
    from oracle_binding import Oracle
    from deepflow_b200.synth import lib
    analyzer = int(lib().synth_dict(synth._h).contents.analyzer_id)
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    cls = P.classes()["TaggedFlow"]
    skip = KG_PREFIX[:-2] + ("team_id",)
    geo = _geo(synth)
    try:
        counter, compared = 0, 0
        for k, n in enumerate(sizes):
            wl = synth.gen(abi.DFGPU_MSG_TAGGEDFLOW, n, seed=seed * 100 + k, flags=flags)
            ora.submit(wl)
            cols, nrows = ora.columns(abi.TABLE_L4_FLOW_LOG)
            recs = _records(wl)
            assert nrows == len(recs) == n
            rows = {c.name: c.rows() for c in cols if not c.name.startswith(skip) or c.name.startswith("gprocess_id")}
            problems = []
            for r, raw in enumerate(recs):
                m = cls()
                m.ParseFromString(raw)
                counter += 1
                want = l4_row(m, counter, analyzer)
                k4 = m.flow.flow_key  # Internet.Fill :530-533: looked up by the IPv4 fields whatever the family
                want["province_0"], want["province_1"] = geo.province(k4.ip_src), geo.province(k4.ip_dst)
                assert set(want) == set(rows), sorted(set(want) ^ set(rows))
                for name, w in want.items():
                    g = rows[name][r]
                    if isinstance(g, np.ndarray):
                        g = g.tolist()
                    if isinstance(w, float):
                        ok = np.float64(g).tobytes() == np.float64(w).tobytes()
                    else:
                        ok = g == w
                    compared += 1
                    if not ok:
                        problems.append(f"batch {k} row {r} {name}: oracle {g!r} != second opinion {w!r}")
                if len(problems) > 15:
                    break
            assert not problems, "\n".join(problems[:15])
        assert compared > 110 * sum(sizes)
    finally:
        ora.close()
