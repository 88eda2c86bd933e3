"""CPU: pin the flow_metrics oracle against what the reference offers:
  * column names / order / container types of all 9 tables against lists machine-extracted
    from the reference's ToInput functions (tests/golden/metrics_columns.json,
    tools/make_metrics_golden.py) and the tag Code constants of tag.go;
  * the meter-merge known answers of basic_meter_test.go:112-135 (TestMerge) and
    usage_meter_test.go:57-79 (TestVTAPMeterMerge);
  * the MiniTag literals of tag_test.go:259-352 (TestEncodeMiniTag) and the direction
    tables of tag_test.go:354-392;
  * decode + ReadFromPB against Google's protobuf runtime on the generator's Documents.
"""
import ctypes as C
import json
import os

import numpy as np

import pb_schema
from deepflow_b200 import abi
from deepflow_b200 import synth as S
from test_wire_cpu import split_records

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "metrics_columns.json")))
MSG_METRICS = 3
TABLES = ["NETWORK", "NETWORK_MAP", "APPLICATION", "APPLICATION_MAP", "TRAFFIC_POLICY",
          "NETWORK", "NETWORK_MAP", "APPLICATION", "APPLICATION_MAP"]
CH = {"ColUInt8": (abi.CH_UINT8, 1), "ColUInt16": (abi.CH_UINT16, 2), "ColUInt32": (abi.CH_UINT32, 4),
      "ColUInt64": (abi.CH_UINT64, 8), "ColInt32": (abi.CH_INT32, 4), "ColFloat64": (abi.CH_FLOAT64, 8),
      "ColIPv4": (abi.CH_IPV4, 4), "ColIPv6": (abi.CH_IPV6, 16), "ColDateTime": (abi.CH_DATETIME, 4),
      "ColStr": (abi.CH_STR, 0)}


def olib():
    import oracle_binding
    L = oracle_binding.lib()
    L.oracle_kat_metrics_columns.restype = C.c_char_p
    L.oracle_kat_metrics_columns.argtypes = [C.c_int]
    L.oracle_kat_table_id.argtypes = [C.c_uint64, C.c_int]
    L.oracle_kat_meter_merge.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
    L.oracle_kat_tag_read.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_int64)]
    L.oracle_kat_doc_read.argtypes = [C.c_char_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_int64),
                                      C.POINTER(C.c_uint64)]
    return L


def expected_columns(table):
    code = GOLD["table_codes"][TABLES[table]]
    bits = GOLD["code_bits"]
    cols = [c for c in GOLD["tag_columns"] if c["cond"] is None or code & bits[c["cond"]]]
    meter = {"NETWORK": "flow_meter_columns", "NETWORK_MAP": "flow_meter_columns", "APPLICATION": "app_meter_columns",
             "APPLICATION_MAP": "app_meter_columns", "TRAFFIC_POLICY": "usage_meter_columns"}[TABLES[table]]
    return cols + GOLD[meter]


def test_columns_match_reference_toinput(built):
    L = olib()
    for table in range(9):
        got = [l.split(":") for l in L.oracle_kat_metrics_columns(table).decode().strip().split("\n")]
        exp = expected_columns(table)
        assert [g[0] for g in got] == [e["name"] for e in exp], table
        for g, e in zip(got, exp):
            if e["type"] == "ColLowCardinality[string]":
                # observation_point: dictionary coded; app_service / app_instance: per-row strings
                assert int(g[1]) in (abi.CH_LC_STR, abi.CH_LC_STR_ROWS), (table, g)
            else:
                ch, es = CH[e["type"]]
                assert (int(g[1]), int(g[2])) == (ch, es), (table, g, e)


def test_survey_column_counts():
    # tag columns per table code, counted there from GenTagColumns (tag.go:933-1141)
    fixed = {"ColUInt8": 1, "ColUInt16": 2, "ColUInt32": 4, "ColInt32": 4, "ColIPv4": 4, "ColIPv6": 16,
             "ColDateTime": 4}
    want = {0: (32, 90, 0), 1: (57, 168, 1), 2: (37, 92, 3), 3: (62, 170, 4), 4: (6, 13, 0)}
    for table, (ncols, nbytes, nstr) in want.items():
        code = GOLD["table_codes"][TABLES[table]]
        cols = [c for c in GOLD["tag_columns"] if c["cond"] is None or code & GOLD["code_bits"][c["cond"]]]
        strs = [c for c in cols if c["type"] not in fixed]
        assert (len(cols), sum(fixed[c["type"]] for c in cols if c["type"] in fixed), len(strs)) == \
            (ncols, nbytes, nstr), table


def test_table_ids(built):
    L = olib()
    for i, name in enumerate(TABLES[:5]):
        code = GOLD["table_codes"][name]
        assert L.oracle_kat_table_id(code, 0) == i
        assert L.oracle_kat_table_id(code | GOLD["code_bits"]["MAC"] | GOLD["code_bits"]["MACPath"], 0) == i
        assert L.oracle_kat_table_id(code, 1) == (i + 5 if i != 4 else -1)
        assert L.oracle_kat_table_id(code & ~GOLD["code_bits"]["VTAPID"], 0) == -1
    assert L.oracle_kat_table_id(0, 0) == -1


def test_meter_merge_kats(built):
    L = olib()
    out = (C.c_uint64 * 80)()
    # TestMerge: every uint field 1, merged with itself
    n = L.oracle_kat_meter_merge(1, 1, 1, out)
    v = list(out)[:n]
    assert n == 65
    names = ("packet_tx packet_rx byte_tx byte_rx l3_byte_tx l3_byte_rx l4_byte_tx l4_byte_rx new_flow closed_flow "
             "l7_request l7_response syn_count synack_count rtt_max rtt_client_max rtt_server_max srt_max art_max "
             "rrt_max cit_max rtt_sum rtt_client_sum rtt_server_sum srt_sum art_sum rrt_sum cit_sum rtt_count "
             "rtt_client_count rtt_server_count srt_count art_count rrt_count cit_count retrans_tx retrans_rx "
             "zero_win_tx zero_win_rx retrans_syn retrans_synack client_rst_flow server_rst_flow server_syn_miss "
             "client_ack_miss client_half_close_flow server_half_close_flow client_source_port_reuse "
             "client_establish_reset server_reset server_queue_lack server_establish_reset tcp_timeout "
             "client_establish_fail server_establish_fail tcp_establish_fail tcp_transfer_fail tcp_rst_fail "
             "l7_client_error l7_server_error l7_timeout ooo_tx ooo_rx flow_load direction_score").split()
    got = dict(zip(names, v))
    for k, x in got.items():
        if k.endswith("_max") or k.endswith("_fail") or k in ("ooo_tx", "ooo_rx"):
            assert x == 1, k  # max / not merged / assigned (basic_meter.go:277-299, 556-577)
        elif k == "direction_score":
            assert x == 0, k
        else:
            assert x == 2, k  # Traffic / Performance / Anomaly sums: initMeter(2)
    # uint32 counters wrap like Go's (l7_request is a uint32 in Traffic)
    n = L.oracle_kat_meter_merge(1, 0xffffffff, 3, out)
    got = dict(zip(names, list(out)[:n]))
    assert got["l7_request"] == 2 and got["packet_tx"] == 0xffffffff + 3 and got["rtt_max"] == 0xffffffff
    assert got["ooo_tx"] == 3  # assigned from `other`
    # TestVTAPMeterMerge: fillMetrics(1) merged into fillMetrics(2) == fillMetrics(3)
    n = L.oracle_kat_meter_merge(4, 2, 1, out)
    assert list(out)[:n] == [9, 6, 291, 267, 288, 264, 285, 261]
    n = L.oracle_kat_meter_merge(5, 5, 7, out)
    assert list(out)[:n] == [12, 12, 0, 7, 12, 12, 12, 12, 12]


def _minitag(**kw):
    Cl = pb_schema.classes()
    t = Cl["MiniTag"]()
    code = kw.pop("code")
    for k, v in kw.items():
        setattr(t.field, k, v)
    t.code = code
    return t.SerializeToString()


def test_minitag_literals(built):
    L = olib()
    b = GOLD["code_bits"]
    out = (C.c_int64 * 32)()
    # sideMiniTag tag_test.go:263-279 (the deprecated TagType/TagValue code bits are gone from the Code enum)
    raw = _minitag(ip=bytes([10, 11, 1, 123]), is_ipv6=0, l3_epc_id=15, vtap_id=24, protocol=6, server_port=5324,
                   direction=1, tap_type=3, acl_gid=16,
                   code=b["IP"] | b["L3EpcID"] | b["VTAPID"] | b["Protocol"] | b["ServerPort"] | b["Direction"] |
                   b["TAPType"] | b["ACLGID"])
    assert L.oracle_kat_tag_read(raw, len(raw), out) == 17
    code, ip0, ip1, epc0, epc1, vtap, proto, port, role, tap_type, acl, is4 = list(out)[:12]
    assert ip0 == 0x0A0B017B and epc0 == 15 and vtap == 24 and proto == 6 and port == 5324 and role == 0
    assert tap_type == 3 and acl == 16 and is4 == 1 and ip1 == 0
    assert code == (b["IP"] | b["L3EpcID"] | b["VTAPID"] | b["Protocol"] | b["ServerPort"] | b["Direction"] |
                    b["TAPType"] | b["ACLGID"] | b["GPID"] | b["SignalSource"])  # tag.go:1218-1225
    assert out[16] == 5324  # tunnel_ip_id = server_port (:1239)
    # edgeMiniTag :296-314, then Direction = ServerToClient :336
    ip6a = bytes([0xfe, 0x80] + [0] * 13 + [1])
    ip6b = bytes([0xfe, 0x80] + [0] * 13 + [2])
    ecode = b["IPPath"] | b["L3EpcIDPath"] | b["VTAPID"] | b["Protocol"] | b["ServerPort"] | b["Direction"] | \
        b["TAPType"] | b["ACLGID"]
    for direction, role_want in ((1, 0), (2, 1)):
        raw = _minitag(ip=ip6a, ip1=ip6b, is_ipv6=1, l3_epc_id=15, l3_epc_id1=16, vtap_id=24, protocol=6,
                       server_port=5324, direction=direction, tap_type=3, acl_gid=16, pod_id=77, code=ecode)
        assert L.oracle_kat_tag_read(raw, len(raw), out) == 17
        v = list(out)[:17]
        assert v[0] == ecode | b["GPIDPath"] | b["SignalSource"]
        assert (v[3], v[4], v[8], v[11], v[12], v[13]) == (15, 16, role_want, 0, 1, 2)
        assert (v[14], v[15]) == ((0, 77) if role_want == 1 else (77, 0))  # pod_id side rule :1229-1235
    # TestDirectionEnum / ToRole: every *ToServer direction is role 0, *ToClient role 1, local 2, else 3
    for d, want in ((1, 0), (1 | 8, 0), (1 | 16, 0), (1 | 24, 0), (1 | 32, 0), (2, 1), (2 | 8, 1), (2 | 32, 1), (4, 2),
                    (0, 3), (3, 3)):
        raw = _minitag(ip=bytes(4), direction=d, code=b["IP"])
        L.oracle_kat_tag_read(raw, len(raw), out)
        assert out[8] == want, d


def test_doc_decode_matches_google_protobuf(synth, built):
    L = olib()
    Cl = pb_schema.classes()
    wl = synth.gen(MSG_METRICS, 6000, seed=21, flags=S.F_DUP_FIELDS | S.F_OMIT_SUBMSG | S.F_WIDE_VARINT)
    hdr = (C.c_uint32 * 3)()
    tag = (C.c_int64 * 32)()
    met = (C.c_uint64 * 80)()
    b = GOLD["code_bits"]
    n_flow = n_app = n_usage = 0
    for rec in split_records(wl):
        d = Cl["Document"]()
        d.ParseFromString(rec)
        k = L.oracle_kat_doc_read(rec, len(rec), hdr, tag, met)
        assert k >= 0
        f = d.tag.field
        assert (hdr[0], hdr[1], hdr[2]) == (d.timestamp, d.flags, d.meter.meter_id & 0xff)
        code = d.tag.code
        if code & b["IPPath"]:
            code |= b["GPIDPath"] | b["SignalSource"]
        elif code != GOLD["table_codes"]["TRAFFIC_POLICY"]:
            code |= b["GPID"] | b["SignalSource"]
        t = list(tag)
        assert t[0] == code - (1 << 64 if code >= 1 << 63 else 0)
        if f.is_ipv6 == 0:
            assert t[1] == int.from_bytes(f.ip[:4].ljust(4, b"\0"), "big")
        else:
            assert t[12] == f.ip[15]
        marsh = lambda e: e if (e > 0 or e in (-1, -2)) else e & 0xffff
        assert (t[3], t[4]) == (marsh(f.l3_epc_id), marsh(f.l3_epc_id1))
        assert (t[5], t[6], t[7]) == (f.vtap_id & 0xffff, f.protocol & 0xff, f.server_port & 0xffff)
        assert (t[17], t[18], t[19], t[20], t[21]) == (f.gpid, f.gpid1, f.signal_source & 0xffff, f.tap_side & 0xff,
                                                      f.l7_protocol & 0xff)
        assert (t[23], t[24], t[25], t[26]) == (f.tap_port & 0xffffffff, f.tap_port >> 60, (f.tap_port >> 36) & 0xf,
                                                (f.tap_port >> 32) & 0xf)
        m = list(met)[:k]
        if hdr[2] == 1:
            n_flow += 1
            tr, la, an = d.meter.flow.traffic, d.meter.flow.latency, d.meter.flow.anomaly
            assert m[0:4] == [tr.packet_tx, tr.packet_rx, tr.byte_tx, tr.byte_rx]
            assert m[10] == tr.l7_request and m[64] == tr.direction_score & 0xff
            assert m[14] == la.rtt_max and m[21] == la.rtt_sum and m[27] == la.cit_sum and m[34] == la.cit_count
            assert m[53] == (an.client_ack_miss + an.client_source_port_reuse + an.client_establish_reset) % 2**64
            assert m[61] == an.client_ooo and m[62] == an.server_ooo and m[63] == d.meter.flow.flow_load.load
        elif hdr[2] == 4:
            n_usage += 1
            u = d.meter.usage
            assert m == [u.packet_tx, u.packet_rx, u.byte_tx, u.byte_rx, u.l3_byte_tx, u.l3_byte_rx, u.l4_byte_tx,
                         u.l4_byte_rx]
        elif hdr[2] == 5:
            n_app += 1
            a = d.meter.app
            assert m == [a.traffic.request, a.traffic.response, a.traffic.direction_score & 0xff, a.latency.rrt_max,
                         a.latency.rrt_sum, a.latency.rrt_count, a.anomaly.client_error, a.anomaly.server_error,
                         a.anomaly.timeout]
    assert n_flow > 1000 and n_app > 1000 and n_usage > 100


def test_oracle_routing_and_reduce(synth, built):
    from oracle_binding import Oracle
    wl = synth.gen(MSG_METRICS, 30000, seed=22)
    ora = Oracle()
    ora.load_synth(synth)
    ora.submit(wl)
    c = ora.counters()
    rows = [c.rows[abi.TABLE_METRICS_BASE + i] for i in range(9)]
    assert c.count == wl.nrecords and c.error_count == 0
    assert sum(rows) + c.drop_count == c.count
    assert all(r > 0 for r in rows) and c.drop_count > 0
    # reduce network_map.1s to 1m buckets: sums are preserved, groups are unique
    t = abi.TABLE_METRICS_BASE + 6
    cols, n = ora.columns(t)
    assert ora.metrics_reduce(t, 60) == 0
    rcols, rn = ora.columns(abi.TABLE_METRICS_REDUCED)
    assert 0 < rn <= n and [x.name for x in rcols] == [x.name for x in cols]
    a = {x.name: x for x in cols}
    r = {x.name: x for x in rcols}
    for name in ("packet_tx", "byte_rx", "rtt_count", "l7_timeout", "flow_load", "retrans"):
        assert int(a[name].values.sum(dtype=np.uint64)) == int(r[name].values.sum(dtype=np.uint64)), name
    assert int(r["rtt_max"].values.max()) == int(a["rtt_max"].values.max())
    assert set(np.unique(r["time"].values) % 60) == {0}
