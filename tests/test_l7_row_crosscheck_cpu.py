"""The l7_flow_log ROW a second time (SURVEY A6 + A8): ProtoLogToL7FlowLog / L7Base.Fill / L7FlowLog.Fill /
fillL7FlowLog / fillTraceIds / fillExceptionDesc / ParseTraceIdIndex / genID
(flow_log/log_data/l7_flow_log.go:296-644, l4_flow_log.go:947-953, exception_desc.go, datatype/tap_port.go:75-115,
libs/utils/utils.go:266-302, hash.go:37-43) and L7Base / L7FlowLog.AppendToColumnBlock
(l7_column_block.go:167-208, 376-422) restated in Python from the Go source, fed by Google's protobuf runtime,
against every non-KnowledgeGraph column of the oracle's rows (the KnowledgeGraph columns have their own second
opinion in test_kg_crosscheck_cpu.py).  Data tables come from machine-extracted files: exception descriptions from
include/dfgpu_exception_tables.h (generated from exception_desc.go), protocol names from
tests/golden/l7_protocols.json."""
import json
import os
import re

import numpy as np
import pytest

import pb_schema as P
from deepflow_b200 import abi
from deepflow_b200 import synth as S

HERE = os.path.dirname(os.path.abspath(__file__))
TAPSIDE = {0: "rest", 1: "c", 2: "s", 4: "local", 9: "c-nd", 10: "s-nd", 17: "c-hv", 18: "s-hv", 25: "c-gw-hv",
           26: "s-gw-hv", 33: "c-gw", 34: "s-gw", 41: "c-p", 42: "s-p", 49: "c-app", 50: "s-app", 48: "app"}  # tag.go:183-233
M64 = (1 << 64) - 1


def _exception_tables():
    src = open(os.path.join(HERE, "..", "include", "dfgpu_exception_tables.h"), encoding="utf-8").read()
    out = {}
    for m in re.finditer(r"DFGPU_EXC_TABLE\((\w+), (\d+)\) = \{(.*?)\n\};", src, re.S):
        vals = [json.loads('"' + s + '"') for s in re.findall(r'/\*\s*\d+ \*/ "((?:[^"\\]|\\.)*)"', m.group(3))]
        assert len(vals) == int(m.group(2)), m.group(1)
        out[m.group(1).replace("dfgpu_exc_", "")] = vals
    unk = re.search(r'#define DFGPU_EXC_KAFKA_UNKNOWN_SERVER_ERROR "([^"]*)"', src).group(1)
    return out, unk


EXC, KAFKA_UNKNOWN = _exception_tables()


def djb(base, b):  # utils.DJBHash
    h = base
    for c in b:
        h = ((h << 5) + h + c) & M64
    return h


def exception_desc(proto, version, code):
    u16, i16 = code & 0xFFFF, ((code & 0xFFFF) ^ 0x8000) - 0x8000
    if proto in (20, 21):
        return EXC["http"][u16] if u16 < len(EXC["http"]) else ""
    if proto == 120:
        if u16 < len(EXC["dns"]):
            return EXC["dns"][u16]
        if 22 <= u16 <= 3840:
            return "available for assignment"
        if 3841 <= u16 <= 4095:
            return "Private Use"
        return "available for assignment" if 4096 <= u16 <= 65535 else ""
    if proto == 40:
        return EXC["dubbo"][u16] if u16 < len(EXC["dubbo"]) else ""
    if proto == 100:
        if i16 == -1:
            return KAFKA_UNKNOWN
        return EXC["kafka"][i16] if 0 < i16 < len(EXC["kafka"]) else ""
    if proto == 101:
        t = EXC["mqtt_v3"] if version != "5" else EXC["mqtt_v5"]
        return t[u16] if 0 < u16 < len(t) else ""
    return ""


class State:
    def __init__(self, analyzer_id):
        self.counter, self.last_trace_index, self.analyzer_id = 0, 0, analyzer_id


def l7_row(m, st, protocols, trace_disabled=False):
    b = m.base
    o = {}
    v6 = b.is_ipv6 == 1
    o["time"] = (b.end_time // 1_000_000_000) & 0xFFFFFFFF
    o["ip4_0"], o["ip4_1"] = (0, 0) if v6 else (b.ip_src, b.ip_dst)
    o["ip6_0"] = (bytes(b.ip6_src) + bytes(16))[:16] if v6 else bytes(16)
    o["ip6_1"] = (bytes(b.ip6_dst) + bytes(16))[:16] if v6 else bytes(16)
    o["is_ipv4"] = 0 if v6 else 1
    o["protocol"] = b.protocol & 0xFF
    o["client_port"], o["server_port"] = b.port_src & 0xFFFF, b.port_dst & 0xFFFF
    o["flow_id"] = b.flow_id
    o["capture_network_type_id"] = b.tap_type & 0xFF
    tp = b.tap_port
    o["capture_nic"], o["capture_nic_type"] = tp & 0xFFFFFFFF, (tp >> 60) & 0xFF
    o["nat_source"], o["tunnel_type"] = (tp >> 36) & 0xF, (tp >> 32) & 0xF
    o["signal_source"] = 4 if o["capture_nic_type"] == 8 else 3 if o["capture_nic_type"] == 7 else 0
    o["observation_point"] = TAPSIDE[b.tap_side].encode()
    o["agent_id"] = b.vtap_id & 0xFFFF
    o["req_tcp_seq"], o["resp_tcp_seq"] = b.req_tcp_seq, b.resp_tcp_seq
    o["start_time"], o["end_time"] = b.start_time // 1000, b.end_time // 1000
    o["biz_type"] = b.biz_type & 0xFF
    o["biz_code"], o["biz_scenario"] = m.biz_code.encode(), m.biz_scenario.encode()
    o["biz_response_code"] = m.biz_response_code.encode()
    o["process_id_0"], o["process_id_1"] = b.process_id_0, b.process_id_1  # int32 casts: compared modulo 2^32
    o["process_kname_0"], o["process_kname_1"] = b.process_kname_0.encode(), b.process_kname_1.encode()
    o["syscall_trace_id_request"], o["syscall_trace_id_response"] = b.syscall_trace_id_request, b.syscall_trace_id_response
    o["syscall_thread_0"], o["syscall_thread_1"] = b.syscall_trace_id_thread_0, b.syscall_trace_id_thread_1
    o["syscall_coroutine_0"], o["syscall_coroutine_1"] = b.syscall_coroutine_0, b.syscall_coroutine_1
    o["syscall_cap_seq_0"], o["syscall_cap_seq_1"] = b.syscall_cap_seq_0, b.syscall_cap_seq_1
    # ---- ProtoLogToL7FlowLog / genID
    st.counter = (st.counter + 1) & 0xFFFFFFFF
    o["_id"] = (o["time"] << 32) | ((st.analyzer_id & 0x3FF) << 22) | (st.counter & 0x3FFFFF)
    # ---- L7FlowLog.Fill
    typ = b.head.msg_type & 0xFF
    o["type"] = typ
    o["is_tls"], o["is_async"], o["is_reversed"] = int(m.flags & 1 != 0), int(m.flags & 2 != 0), int(m.flags & 4 != 0)
    proto = b.head.proto & 0xFF
    o["l7_protocol"] = proto
    has_ext, has_req, has_resp, has_trace = (m.HasField(f) for f in ("ext_info", "req", "resp", "trace_info"))
    if has_ext and m.ext_info.protocol_str != "":
        o["biz_protocol"] = m.ext_info.protocol_str.encode()
    else:
        names = protocols["names"].get(str(proto))
        o["biz_protocol"] = (names[1 if o["is_tls"] else 0] if names else protocols["default"]).encode()
    status = 5  # STATUS_UNKNOWN
    o["response_duration"] = b.head.rrt // 1000
    # ---- fillL7FlowLog
    o["version"] = m.version.encode()
    req_len, resp_len = m.req_len, m.resp_len
    o["sql_affected_rows"] = m.row_effect if m.row_effect != 0 else None
    o["direction_score"] = m.direction_score & 0xFF
    o["captured_request_byte"], o["captured_response_byte"] = m.captured_request_byte, m.captured_response_byte
    o["request_domain"] = o["request_resource"] = o["request_type"] = o["endpoint"] = b""
    o["request_length"] = o["response_length"] = o["response_code"] = o["request_id"] = None
    o["response_result"] = o["response_exception"] = b""
    response_code, request_id = 0, 0
    if has_req:
        o["request_domain"], o["request_resource"] = m.req.domain.encode(), m.req.resource.encode()
        o["request_type"], o["endpoint"] = m.req.req_type.encode(), m.req.endpoint.encode()
        if req_len != -1 and typ != 1:
            o["request_length"] = req_len
    if has_resp:
        status = m.resp.status & 0xFF
        if typ != 0:
            o["response_result"] = m.resp.result.encode()
            response_code = m.resp.code
            exc = m.resp.exception
            if exc == "" and status in (3, 4):
                exc = exception_desc(proto, m.version, m.resp.code)
            o["response_exception"] = exc.encode()
            if response_code != -32768:
                o["response_code"] = response_code
            if resp_len != -1:
                o["response_length"] = resp_len
    an, av, mn, mv = [], [], [], []
    o["app_service"] = o["x_request_id_0"] = o["x_request_id_1"] = o["http_proxy_client"] = b""
    if has_ext:
        e = m.ext_info
        request_id = e.request_id
        if request_id != 0:
            o["request_id"] = request_id
        o["app_service"], o["http_proxy_client"] = e.service_name.encode(), e.client_ip.encode()
        o["x_request_id_0"], o["x_request_id_1"] = e.x_request_id_0.encode(), e.x_request_id_1.encode()
        for name, val in (("http_user_agent", e.http_user_agent), ("http_referer", e.http_referer),
                          ("rpc_service", e.rpc_service)):
            if val != "":
                an.append(name.encode())
                av.append(val.encode())
        an += [x.encode() for x in e.attribute_names]
        av += [x.encode() for x in e.attribute_values]
        mn += [x.encode() for x in e.metrics_names]
        mv += list(e.metrics_values)
    o["attribute_names"], o["attribute_values"], o["metrics_names"], o["metrics_values"] = an, av, mn, mv
    trace_id = trace_id_2 = ""
    o["span_id"] = o["parent_span_id"] = b""
    if has_trace:
        t = m.trace_info
        o["span_id"], o["parent_span_id"] = t.span_id.encode(), t.parent_span_id.encode()
        if len(t.trace_ids) > 0:
            trace_id = t.trace_ids[0]
        if len(t.trace_ids) > 1:
            trace_id_2 = t.trace_ids[1]
        if trace_id == "":
            trace_id = t.trace_id
    o["trace_id"], o["_trace_id_2"] = trace_id.encode(), trace_id_2.encode()
    if trace_disabled:
        o["trace_id_index"] = 0
    elif trace_id == "":
        o["trace_id_index"] = (st.last_trace_index + 1) & M64
    else:
        st.last_trace_index = djb(17, trace_id.encode())
        o["trace_id_index"] = st.last_trace_index
    if proto == 100 and has_req:  # Kafka
        if response_code == 0 and m.req.req_type not in ("Fetch", "Produce", "JoinGroup", "LeaveGroup", "SyncGroup"):
            status = 2
            o["response_code"] = None
        o["request_id"] = request_id
    elif proto == 43:  # SofaRPC
        o["request_id"] = request_id
    o["response_status"] = status
    o["span_kind"] = None
    o["app_instance"], o["events"] = b"", b""
    return o


def _records(wl):
    out = []
    for f in wl.frames:
        p, end = int(f["offset"]), int(f["offset"]) + int(f["len"])
        while p < end:
            n = int.from_bytes(bytes(wl.data[p:p + 4]), "little")
            out.append(bytes(wl.data[p + 4:p + 4 + n]))
            p += 4 + n
    return out


KG_PREFIX = ("region_id", "az_id", "host_id", "l3_device", "pod_", "l3_epc_id", "epc_id", "subnet_id", "service_id",
             "auto_", "tag_source", "gprocess_id", "team_id")


@pytest.mark.parametrize("sizes,seed,flags", [([1500, 1500], 7, 0), ([1200], 8, S.F_ALL_WIRE)])
def test_l7_rows_match_second_opinion(synth, built, sizes, seed, flags):
    from oracle_binding import Oracle
    from deepflow_b200.synth import lib
    protocols = json.load(open(os.path.join(HERE, "golden", "l7_protocols.json")))
    analyzer = int(lib().synth_dict(synth._h).contents.analyzer_id)
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    st = State(analyzer)
    cls = P.classes()["AppProtoLogsData"]
    try:
        compared = 0
        for k, n in enumerate(sizes):
            wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, n, seed=seed * 100 + k, flags=flags)
            ora.submit(wl)
            cols, nrows = ora.columns(abi.TABLE_L7_FLOW_LOG)
            recs = _records(wl)
            assert nrows == len(recs) == n
            rows = {c.name: c.rows() for c in cols if not c.name.startswith(KG_PREFIX)}
            kinds = {c.name: c.ch_type for c in cols}
            problems = []
            for r, raw in enumerate(recs):
                m = cls()
                m.ParseFromString(raw)
                want = l7_row(m, st, protocols)
                assert set(want) == set(rows), sorted(set(want) ^ set(rows))
                for name, w in want.items():
                    g = rows[name][r]
                    if kinds[name] == abi.CH_INT32 and w is not None:
                        w = ((w & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000
                    if isinstance(g, np.ndarray):
                        g = g.tolist()
                    if name == "metrics_values":
                        g, w = [float(x) for x in g], [float(x) for x in w]
                        if np.array(g, np.float64).tobytes() == np.array(w, np.float64).tobytes():
                            continue
                    if g != w:
                        problems.append(f"batch {k} row {r} {name}: oracle {g!r} != second opinion {w!r}")
                    compared += 1
                if len(problems) > 15:
                    break
            assert not problems, "\n".join(problems[:15])
        assert compared > 80 * sum(sizes) * 0.9
    finally:
        ora.close()
