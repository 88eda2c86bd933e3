"""The wire schema the whole test stack stands on (tests/pb_schema.py: the descriptors Google's protobuf
runtime uses to cross-check the generator's encoder and the oracle's decoder, tests/test_wire_cpu.py) against
the field lists MACHINE-EXTRACTED from the reference's message/flow_log.proto and message/metric.proto
(tests/golden/proto_fields.json, made by tools/make_proto_golden.py): every field the path declares must exist
in the reference message with the same number, scalar type / sub-message and repeated-ness."""
import json
import os

import pb_schema as P

_T = {P.U32: "uint32", P.U64: "uint64", P.I32: "int32", P.STR: "string", P.BYT: "bytes", P.DBL: "double"}
# messages of the two files the path decodes (the rest of metric.proto — Profile, ProcEvent ... — is out of scope)
_FILE_OF = {"flow_log.proto": ["TaggedFlow", "Flow", "FlowKey", "FlowMetricsPeer", "TunnelField", "FlowPerfStats",
                               "TCPPerfStats", "TcpPerfCountsPeer", "L7PerfStats", "L7Request", "L7Response",
                               "TraceInfo", "ExtendedInfo", "AppProtoLogsData", "AppProtoLogsBaseInfo",
                               "AppProtoHead"],
            "metric.proto": ["MiniField", "MiniTag", "Meter", "Document", "FlowMeter", "Traffic", "Latency",
                             "Performance", "Anomaly", "FlowLoad", "UsageMeter", "AppMeter", "AppTraffic",
                             "AppLatency", "AppAnomaly"]}


def _golden():
    with open(os.path.join(os.path.dirname(__file__), "golden", "proto_fields.json")) as f:
        return json.load(f)


def test_every_declared_field_matches_the_reference_proto():
    g = _golden()
    where = {m: f for f, ms in _FILE_OF.items() for m in ms}
    problems = []
    for msg, fields in P.SCHEMA.items():
        assert msg in where, f"{msg}: not mapped to a reference .proto"
        ref = {r[1]: r for r in g[where[msg]][msg]}  # by field number
        for f in fields:
            name, num, typ = f[0], f[1], f[2]
            sub = f[3] if len(f) > 3 else None
            rep = bool(f[4]) if len(f) > 4 else False
            if num not in ref:
                problems.append(f"{msg}.{name} = {num}: no such field number in the reference")
                continue
            rname, _, rtype, rrep = ref[num]
            if rname.lower() != name.lower():
                problems.append(f"{msg} field {num}: name {name} != {rname}")
            want = sub if typ == P.MSG else _T[typ]
            if rtype.split(".")[-1] != want:
                problems.append(f"{msg}.{name}: type {want} != {rtype}")
            if (rrep == "repeated") != rep:
                problems.append(f"{msg}.{name}: repeated {rep} != {rrep!r}")
    assert not problems, "\n".join(problems)


def test_messages_on_the_path_are_declared_completely():
    """The decoders must know EVERY field of the messages they fill rows from (an unknown field is skipped
    silently): the reference's field numbers of these messages are all declared."""
    g = _golden()
    complete = list(P.SCHEMA)
    where = {m: f for f, ms in _FILE_OF.items() for m in ms}
    problems = []
    for msg in complete:
        have = {f[1] for f in P.SCHEMA[msg]}
        for rname, num, _t, _r in g[where[msg]][msg]:
            if num not in have:
                problems.append(f"{msg}: reference field {rname} = {num} not declared")
    assert not problems, "\n".join(problems)
