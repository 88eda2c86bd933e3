"""CPU: pin the wire format.  The synthetic generator's encoder and the oracle's
decoder are both checked against Google's protobuf implementation (an independent
third opinion), and against the field literal of the reference's own round-trip test
(server/libs/datatype/tagged_flow_test.go:28-122)."""
import numpy as np

import pb_schema
from deepflow_b200 import abi
from deepflow_b200 import synth as S


def split_records(wl):
    for fr in wl.frames:
        buf = wl.data[int(fr["offset"]):int(fr["offset"]) + int(fr["len"])].tobytes()
        off = 0
        while off < len(buf):
            n = int.from_bytes(buf[off:off + 4], "little")
            yield buf[off + 4:off + 4 + n]
            off += 4 + n


def one_frame(payload, msg_type, org=1, team=1, vtap=1):
    fr = np.zeros(1, abi.FRAME_DTYPE)
    fr["len"], fr["org_id"], fr["team_id"], fr["vtap_id"] = len(payload), org, team, vtap
    return S.Workload(np.frombuffer(payload, np.uint8).copy(), fr, None, msg_type)


def test_generator_output_parses_with_google_protobuf(synth, built):
    C = pb_schema.classes()
    for mt, cls in ((abi.DFGPU_MSG_TAGGEDFLOW, "TaggedFlow"), (abi.DFGPU_MSG_PROTOCOLLOG, "AppProtoLogsData")):
        # F_UNKNOWN_FIELDS uses groups, which python-protobuf also skips
        wl = synth.gen(mt, 3000, seed=11, flags=S.F_UNPACKED | S.F_DUP_FIELDS | S.F_OMIT_SUBMSG)
        n = 0
        for rec in split_records(wl):
            m = C[cls]()
            m.ParseFromString(rec)  # raises on malformed wire data
            n += 1
        assert n == wl.nrecords


def test_oracle_decode_matches_google_protobuf_l4(synth, built):
    from oracle_binding import Oracle
    C = pb_schema.classes()
    wl = synth.gen(abi.DFGPU_MSG_TAGGEDFLOW, 4000, seed=12, flags=S.F_UNPACKED | S.F_DUP_FIELDS | S.F_OMIT_SUBMSG)
    ora = Oracle()
    ora.load_synth(synth)
    ora.submit(wl)
    cols, n = ora.columns(abi.TABLE_L4_FLOW_LOG)
    col = {c.name: c for c in cols}
    msgs = []
    for rec in split_records(wl):
        m = C["TaggedFlow"]()
        m.ParseFromString(rec)
        msgs.append(m.flow)
    assert n == len(msgs)
    U64 = lambda a: np.array(a, dtype=np.uint64)
    assert np.array_equal(col["mac_0"].values, U64([f.flow_key.mac_src for f in msgs]))
    assert np.array_equal(col["mac_1"].values, U64([f.flow_key.mac_dst for f in msgs]))
    assert np.array_equal(col["flow_id"].values, U64([f.flow_id for f in msgs]))
    assert np.array_equal(col["byte_tx"].values, U64([f.metrics_peer_src.byte_count for f in msgs]))
    assert np.array_equal(col["total_packet_rx"].values, U64([f.metrics_peer_dst.total_packet_count for f in msgs]))
    assert np.array_equal(col["client_port"].values, np.array([f.flow_key.port_src & 0xffff for f in msgs], np.uint16))
    assert np.array_equal(col["l3_epc_id_0"].values.astype(np.int64),
                          np.array([e if (e > 0 or e in (-1, -2)) else e & 0xffff
                                    for e in (f.metrics_peer_src.l3_epc_id for f in msgs)], np.int64))
    assert np.array_equal(col["close_type"].values, np.array([f.close_type & 0xffff for f in msgs], np.uint16))
    assert col["request_domain"].strings() == [f.request_domain.encode() for f in msgs]
    assert col["acl_gids"].rows() == [[g & 0xffff for g in f.acl_gids] for f in msgs]
    assert col["aggregated_flow_ids"].strings() == [",".join(str(i) for i in f.aggregated_flow_ids).encode() for f in msgs]
    perf = [f.has_perf_stats == 1 for f in msgs]
    assert np.array_equal(col["rrt_sum"].values,
                          np.array([float(f.perf_stats.l7.rrt_sum) if p else 0.0 for f, p in zip(msgs, perf)]))
    assert np.array_equal(col["retrans_tx"].values,
                          np.array([f.perf_stats.tcp.counts_peer_tx.retrans_count if p else 0 for f, p in zip(msgs, perf)], np.uint32))


def test_oracle_decode_matches_google_protobuf_l7(synth, built):
    from oracle_binding import Oracle
    C = pb_schema.classes()
    wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 4000, seed=13, flags=S.F_UNPACKED | S.F_OMIT_SUBMSG)
    ora = Oracle()
    ora.load_synth(synth)
    ora.submit(wl)
    cols, n = ora.columns(abi.TABLE_L7_FLOW_LOG)
    col = {c.name: c for c in cols}
    msgs = []
    for rec in split_records(wl):
        m = C["AppProtoLogsData"]()
        m.ParseFromString(rec)
        msgs.append(m)
    assert n == len(msgs)
    assert np.array_equal(col["flow_id"].values, np.array([m.base.flow_id for m in msgs], np.uint64))
    assert col["request_domain"].strings() == [m.req.domain.encode() if m.HasField("req") else b"" for m in msgs]
    assert col["request_resource"].strings() == [m.req.resource.encode() if m.HasField("req") else b"" for m in msgs]
    assert col["span_id"].strings() == [m.trace_info.span_id.encode() if m.HasField("trace_info") else b"" for m in msgs]
    assert col["version"].strings() == [m.version.encode() for m in msgs]
    exp_names = []
    for m in msgs:
        names = []
        if m.HasField("ext_info"):
            e = m.ext_info
            if e.http_user_agent:
                names.append(b"http_user_agent")
            if e.http_referer:
                names.append(b"http_referer")
            if e.rpc_service:
                names.append(b"rpc_service")
            names += [x.encode() for x in e.attribute_names]
        exp_names.append(names)
    assert col["attribute_names"].rows() == exp_names
    assert col["metrics_values"].rows() == [list(m.ext_info.metrics_values) if m.HasField("ext_info") else [] for m in msgs]
    assert np.array_equal(col["response_duration"].values, np.array([m.base.head.rrt // 1000 for m in msgs], np.uint64))


def test_reference_tagged_flow_literal(synth, built):
    """The literal of TestTaggedFlowEncodeDecode (tagged_flow_test.go:29-103), encoded by
    Google protobuf, decoded + packed by the oracle: every carried field must land in
    its column (the reference test only checks encode == decode)."""
    from oracle_binding import Oracle
    C = pb_schema.classes()
    tf = C["TaggedFlow"]()
    f = tf.flow
    k = f.flow_key
    k.vtap_id, k.tap_type, k.tap_port = 200, 3, 201  # FromLocalMAC(0, 201)
    k.mac_src, k.mac_dst, k.ip_src, k.ip_dst = 20000000002, 20000000003, 204, 205
    k.ip6_src, k.ip6_dst = bytes([2] * 16), bytes([3] * 16)
    k.port_src, k.port_dst, k.proto = 206, 207, 6
    for p in (f.metrics_peer_src, f.metrics_peer_dst):
        p.byte_count, p.l3_byte_count, p.l4_byte_count, p.packet_count = 303, 310, 311, 304
        p.total_byte_count, p.total_packet_count, p.first, p.last = 305, 306, 307, 308
        p.l3_epc_id, p.tcp_flags, p.is_l2_end, p.is_l3_end, p.is_active_host = 312, 222, 1, 0, 1
    t = f.Tunnel
    t.tx_ip0, t.tx_ip1, t.rx_ip0, t.rx_ip1, t.tx_id, t.rx_id, t.tunnel_type = 100, 200, 200, 300, 102, 102, 1
    f.flow_id, f.start_time, f.end_time, f.duration = 401, 404, 405, 406
    f.eth_type, f.has_perf_stats, f.close_type = 0x0800, 1, 2
    f.perf_stats.l4_protocol = 1
    for c in (f.perf_stats.tcp.counts_peer_tx, f.perf_stats.tcp.counts_peer_rx):
        c.retrans_count, c.zero_win_count = 601, 602
    f.perf_stats.tcp.total_retrans_count = 508
    wl = one_frame(pb_schema.frame([tf.SerializeToString()]), abi.DFGPU_MSG_TAGGEDFLOW, vtap=200)
    ora = Oracle()
    ora.load_synth(synth)
    ora.submit(wl)
    cols, n = ora.columns(abi.TABLE_L4_FLOW_LOG)
    assert n == 1
    v = {c.name: c.rows()[0] for c in cols}
    assert (v["mac_0"], v["mac_1"], v["ip4_0"], v["ip4_1"]) == (20000000002, 20000000003, 204, 205)
    assert v["ip6_0"] == bytes(16) and v["is_ipv4"] == 1  # eth_type v4: ip6 fields are not copied
    assert (v["client_port"], v["server_port"], v["protocol"], v["agent_id"]) == (206, 207, 6, 200)
    assert (v["byte_tx"], v["l3_byte_rx"], v["l4_byte_tx"], v["packet_rx"]) == (303, 310, 311, 304)
    assert (v["total_byte_tx"], v["total_packet_rx"]) == (305, 306)
    assert (v["l3_epc_id_0"], v["l3_epc_id_1"]) == (312, 312)
    assert (v["tcp_flags_bit_0"], v["l2_end_0"], v["l3_end_1"]) == (222, 1, 0)
    assert (v["tunnel_type"], v["tunnel_tx_id"], v["tunnel_rx_ip4_1"], v["tunnel_is_ipv4"]) == (1, 102, 300, 1)
    assert (v["flow_id"], v["start_time"], v["end_time"], v["duration"], v["time"]) == (401, 0, 0, 0, 0)
    assert (v["close_type"], v["status"]) == (2, 3)  # TCPServerRst on TCP -> server error
    assert (v["retrans_tx"], v["retrans_rx"], v["zero_win_tx"]) == (601, 601, 602)
    assert (v["capture_nic"], v["capture_nic_type"], v["capture_network_type_id"]) == (201, 0, 3)
    assert v["_id"] == (0 << 32) | ((7 & 0x3ff) << 22) | 1  # analyzer_id 7 (synth), first record


# ---- droplet message headers + org / team normalisation (include/dfgpu_wire.h) ------------------
def _message(msg_type, payload, version=0x8000, encoder=0, team=7, org=3, agent=1234, frame_size=None):
    """BaseHeader.Encode + FlowHeader.Encode (droplet-message.go:177-180, 248-256): u32 BE frame size (headers
    included), u8 type | u16 LE version, u8 encoder, u32 LE team, u16 LE org, u16 reserved, u16 LE agent id,
    u8 reserved."""
    import struct
    flow = struct.pack("<HBIHHHB", version, encoder, team, org, 0, agent, 0)
    assert len(flow) == 14
    size = 5 + len(flow) + len(payload) if frame_size is None else frame_size
    return struct.pack(">IB", size, msg_type) + flow + payload


def _parse(msg):
    import ctypes as C
    import os
    from deepflow_b200 import build
    build.build_all()
    L = C.CDLL(os.path.join(os.path.dirname(build.__file__), "host", "libdfhost.so"))
    L.dfhost_parse_message.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_int64)]
    out = (C.c_int64 * 9)()
    L.dfhost_parse_message(msg, len(msg), out)
    keys = ("rc", "msg_type", "payload_off", "len", "vtap", "org", "team", "encoder", "invalid")
    return dict(zip(keys, list(out)))


def test_droplet_headers_and_org_team_rules():
    body = b"\x05\x00\x00\x00hello"
    r = _parse(_message(5, body))
    assert r == dict(rc=0, msg_type=5, payload_off=19, len=len(body), vtap=1234, org=3, team=7, encoder=0, invalid=0)
    # parseOrgIdTeamId (receiver.go:688-704): org 0 -> 1, team 0 -> 1, org > 1024 -> 1 and counted
    r = _parse(_message(4, body, team=0, org=0))
    assert (r["org"], r["team"], r["invalid"]) == (1, 1, 0)
    r = _parse(_message(3, body, org=1025))
    assert (r["org"], r["invalid"]) == (1, 1)
    assert _parse(_message(3, body, org=1024))["org"] == 1024
    # encoder flag travels to the caller (compressed frames are inflated before staging)
    assert _parse(_message(5, body, encoder=3))["encoder"] == 3
    # header of agents <= v6.5.8 (FlowHeaderOld): version != 0x8000 -> default org / team, vtap at offset 12
    import struct
    old = struct.pack(">IB", 5 + 14 + len(body), 4) + struct.pack("<IIIH", 20210101, 9, 9, 77) + body
    r = _parse(old)
    assert (r["rc"], r["vtap"], r["org"], r["team"], r["len"]) == (0, 77, 1, 1, len(body))
    # BaseHeader.Decode errors: frame size smaller than the headers, unknown message type
    assert _parse(_message(5, b"", frame_size=18))["rc"] == abi.DFGPU_EINVAL if hasattr(abi, "DFGPU_EINVAL") else _parse(_message(5, b"", frame_size=18))["rc"] == -22
    assert _parse(_message(200, body))["rc"] == -22
    # message types outside this path are not ours (ENOENT), a short buffer is invalid
    assert _parse(_message(6, body))["rc"] == -2
    assert _parse(_message(5, body)[:10] + b"")["rc"] == -22
