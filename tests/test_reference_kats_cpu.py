"""CPU: known-answer tests lifted from the reference's own unit tests, run against
the oracle's restatements (SURVEY §8c list)."""
import ctypes as C

import numpy as np


def lib():
    from oracle_binding import lib as L
    l = L()
    l.oracle_kat_marshal_int32.restype = C.c_int32
    l.oracle_kat_marshal_int32.argtypes = [C.c_int32]
    l.oracle_kat_tapside.restype = C.c_char_p
    l.oracle_kat_tapside.argtypes = [C.c_uint32]
    l.oracle_kat_djb.restype = C.c_uint64
    l.oracle_kat_djb.argtypes = [C.c_uint64, C.c_char_p, C.c_uint32]
    l.oracle_kat_first_mask_sequence.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
    l.oracle_kat_lru.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    l.oracle_kat_is_local_ip.argtypes = [C.c_int, C.c_uint32, C.c_char_p, C.c_uint32]
    l.oracle_kat_is_loopback.argtypes = [C.c_int, C.c_uint32, C.c_char_p, C.c_uint32]
    l.oracle_kat_province.restype = C.c_char_p
    l.oracle_kat_province.argtypes = [C.c_void_p, C.c_uint32]
    return l


def test_int32_special_id_roundtrip(built):
    """libs/flow-metrics/tag_test.go:45-56 TestInt16Unmarshal: int16(marshal(v)) == v."""
    l = lib()
    for i in range(-32768, 32768, 7):
        m = l.oracle_kat_marshal_int32(i)
        lo = m & 0xffff
        assert (lo - 65536 if lo >= 32768 else lo) == i  # int16(v) of the marshalled value
    assert l.oracle_kat_marshal_int32(-1) == -1 and l.oracle_kat_marshal_int32(-2) == -2
    assert l.oracle_kat_marshal_int32(-3) == 65533 and l.oracle_kat_marshal_int32(0) == 0


def test_tap_side_strings(built):
    """tag_test.go:371-392 TestDirectionToTAPSide enum values + tag.go:211-233 names."""
    l = lib()
    node, hv, gwhv, gw, proc, app = 8, 16, 24, 32, 40, 48
    exp = {0: b"rest", 1: b"c", 2: b"s", 4: b"local", 1 | node: b"c-nd", 2 | node: b"s-nd", 1 | hv: b"c-hv",
           2 | hv: b"s-hv", 1 | gwhv: b"c-gw-hv", 2 | gwhv: b"s-gw-hv", 1 | gw: b"c-gw", 2 | gw: b"s-gw",
           1 | proc: b"c-p", 2 | proc: b"s-p", 1 | app: b"c-app", 2 | app: b"s-app", app: b"app"}
    for v in range(51):
        assert l.oracle_kat_tapside(v) == exp.get(v, b"")


def test_get_first_mask(built):
    """libs/geo/netmask_tree_test.go:24-52 TestGetFirstMask."""
    l = lib()
    pre = (C.c_uint32 * 8)()
    ml = (C.c_uint8 * 8)()
    l.oracle_kat_first_mask_sequence(0x01000000, 0x010000fe, pre, ml, 8)
    assert list(pre) == [0x01000000, 0x01000080, 0x010000c0, 0x010000e0, 0x010000f0, 0x010000f8, 0x010000fc, 0x010000fe]
    assert list(ml) == [25, 26, 27, 28, 29, 30, 31, 32]
    l.oracle_kat_first_mask_sequence(0x01000000, 0x010000ff, pre, ml, 1)
    assert (pre[0], ml[0]) == (0x01000000, 24)


def test_lru_order_and_eviction(built):
    """libs/lru/lru_test.go:23-36 TestKeys (oldest -> newest) + capacity eviction and
    MoveToFront on AddOrGet (lru.go:44-60)."""
    l = lib()
    out = (C.c_int * 200)()
    n = l.oracle_kat_lru(100, 100, -1, out)
    assert n == 100 and list(out[:100]) == list(range(100))
    n = l.oracle_kat_lru(100, 130, -1, out)  # 30 oldest evicted
    assert n == 100 and list(out[:100]) == list(range(30, 130))
    n = l.oracle_kat_lru(100, 100, 5, out)  # touching 5 makes it the newest
    assert list(out[:100]) == [i for i in range(100) if i != 5] + [5]


def test_djb_hash(built):
    """libs/utils/hash.go:37-43: hash = hash*33 + c, base 17 for trace ids."""
    l = lib()

    def ref(base, s):
        h = base
        for c in s:
            h = ((h << 5) + h + c) & 0xFFFFFFFFFFFFFFFF
        return h

    for s in (b"", b"a", b"0123456789abcdef0123456789abcdef", bytes(range(256))):
        assert l.oracle_kat_djb(17, s, len(s)) == ref(17, s)


def test_go_ip_predicates(built):
    """Go net.IP semantics behind utils.IsLocalIP / IsLoopback (utils.go:116-130)."""
    l = lib()
    ip4 = lambda a, b, c, d: (a << 24) | (b << 16) | (c << 8) | d
    local4 = {ip4(0, 0, 0, 0), ip4(255, 255, 255, 255), ip4(127, 1, 2, 3), ip4(224, 0, 0, 1), ip4(239, 9, 9, 9),
              ip4(169, 254, 0, 7)}
    for v in local4:
        assert l.oracle_kat_is_local_ip(0, v, None, 0) == 1
    for v in (ip4(10, 0, 0, 1), ip4(8, 8, 8, 8), ip4(169, 253, 0, 1), ip4(223, 255, 255, 255), ip4(192, 168, 1, 1)):
        assert l.oracle_kat_is_local_ip(0, v, None, 0) == 0
    assert l.oracle_kat_is_loopback(0, ip4(127, 0, 0, 1), None, 0) == 1
    assert l.oracle_kat_is_loopback(0, ip4(128, 0, 0, 1), None, 0) == 0
    v6 = lambda s: bytes.fromhex(s)
    assert l.oracle_kat_is_loopback(1, 0, v6("00" * 15 + "01"), 16) == 1
    assert l.oracle_kat_is_local_ip(1, 0, v6("00" * 16), 16) == 1           # ::
    assert l.oracle_kat_is_local_ip(1, 0, v6("fe80" + "00" * 14), 16) == 1   # link local
    assert l.oracle_kat_is_local_ip(1, 0, v6("ff02" + "00" * 14), 16) == 1   # multicast
    assert l.oracle_kat_is_local_ip(1, 0, v6("fd00" + "11" * 14), 16) == 0   # ULA is global unicast in Go
    assert l.oracle_kat_is_local_ip(1, 0, v6("00" * 10 + "ffff" + "7f000001"), 16) == 1  # v4-mapped loopback
    assert l.oracle_kat_is_local_ip(1, 0, v6("00" * 10 + "ffff" + "0a000001"), 16) == 0
    assert l.oracle_kat_is_local_ip(1, 0, b"", 0) == 1                        # nil ip is not global unicast
    assert l.oracle_kat_is_loopback(1, 0, b"", 0) == 0


def test_geo_tianjin_golden(built):
    """libs/geo/netmask_tree_test.go:54-64 TestNetmaskGeoTree: 223.104.x -> "天津", on the
    slice of the reference's ip_info_mini.json committed under tests/golden/."""
    import json
    import os
    from deepflow_b200 import abi
    from oracle_binding import Oracle
    path = os.path.join(os.path.dirname(__file__), "golden", "geo_slice.json")
    g = json.load(open(path, encoding="utf-8"))
    names = [n.encode("utf-8") for n in g["region_names"]]
    ent = (abi.GeoEntry * len(g["entries"]))()
    for i, (s, e, r) in enumerate(g["entries"]):
        ent[i].ip_start, ent[i].ip_end, ent[i].region = s, e, r
    arr = (C.c_char_p * len(names))(*names)
    ora = Oracle()
    ora.geo_load((C.cast(ent, C.c_void_p), len(g["entries"]), C.cast(arr, C.c_void_p), len(names)))
    l = lib()
    assert l.oracle_kat_province(ora._h, 3748071168).decode("utf-8") == "天津"
    assert l.oracle_kat_province(ora._h, 1).decode("utf-8") == "未知"
    for ip, region in g["probes"]:
        assert l.oracle_kat_province(ora._h, ip).decode("utf-8") == region


def test_l7_protocol_strings_match_reference_switch(built):
    """L7Protocol.String(isTLS) for every uint8 value against the table machine-extracted from the reference's
    switch (datatype/flow.go:150-183 constants, :623-803 String; tests/golden/l7_protocols.json,
    tools/make_l7proto_golden.py) — on the oracle, and on the device-side table literal of l7_tables_host.h."""
    import json
    import os
    import re
    l = lib()
    l.oracle_kat_l7_protocol_string.restype = C.c_char_p
    l.oracle_kat_l7_protocol_string.argtypes = [C.c_uint32, C.c_int]
    here = os.path.dirname(os.path.abspath(__file__))
    g = json.load(open(os.path.join(here, "golden", "l7_protocols.json")))
    assert len(g["names"]) == 31 and g["default"] == "N/A"
    for p in range(256):
        plain, tls = g["names"].get(str(p), [g["default"], g["default"]])
        assert l.oracle_kat_l7_protocol_string(p, 0).decode() == plain, p
        assert l.oracle_kat_l7_protocol_string(p, 1).decode() == tls, p
    # the CUDA side keeps its own literal (name, name + "_TLS"; 121 = "TLS" both ways; the rest "N/A")
    src = open(os.path.join(here, "..", "deepflow_b200", "csrc", "l7_tables_host.h"), encoding="utf-8").read()
    lit = re.search(r"static const PN protos\[\] = \{(.*?)\};", src, re.S).group(1)
    dev = {int(a): b for a, b in re.findall(r"\{(\d+),\s*\"([^\"]*)\"\}", lit)}
    want = {int(k): v for k, v in g["names"].items()}
    assert want.pop(121) == ["TLS", "TLS"] and "tb.proto_cs[121][0] = tb.proto_cs[121][1] = CS_PROTO_TLS" in src
    assert dev == {k: v[0] for k, v in want.items()}
    assert all(v[1] == v[0] + "_TLS" for v in want.values())


def test_auto_instance_service_type_codes_match_trident_enums(built):
    """common.GetAutoInstance / GetAutoService / IsPodServiceIP (ingester/common/common.go:137-198) return
    trident.AutoServiceType / DeviceType codes; the values are MACHINE-EXTRACTED from message/trident.proto
    (tests/golden/proto_fields.json "trident.proto:enums") — the Go source names them by enum, and its inline
    comment for POD_SERVICE ("// 12") is stale: the proto says 11."""
    import json
    import os
    l = lib()
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proto_fields.json")))
    A, D = g["trident.proto:enums"]["AutoServiceType"], g["trident.proto:enums"]["DeviceType"]
    POD, POD_NODE = A["AUTO_SERVICE_TYPE_POD"], A["AUTO_SERVICE_TYPE_POD_NODE"]
    POD_SERVICE, POD_CLUSTER = A["AUTO_SERVICE_TYPE_POD_SERVICE"], A["AUTO_SERVICE_TYPE_POD_CLUSTER"]
    CUSTOM, PROCESS = A["AUTO_SERVICE_TYPE_CUSTOM_SERVICE"], A["AUTO_SERVICE_TYPE_PROCESS"]
    IP, INTERNET = A["AUTO_SERVICE_TYPE_IP"], A["AUTO_SERVICE_TYPE_INTERNET_IP"]
    assert (POD, POD_NODE, POD_SERVICE, POD_CLUSTER, CUSTOM, PROCESS, IP, INTERNET) == (10, 14, 11, 103, 104, 120, 255, 0)
    out = (C.c_uint32 * 2)()

    def inst(pod, gp, node, dev, subnet, l3type, epc):
        l.oracle_kat_auto_instance(pod, gp, node, dev, subnet, l3type, C.c_int32(epc), out)
        return out[0], out[1]

    def svc(custom, ps, pg, gp, pc, dev, subnet, l3type, pgtype, epc):
        l.oracle_kat_auto_service(custom, ps, pg, gp, pc, dev, subnet, l3type, pgtype, C.c_int32(epc), out)
        return out[0], out[1]

    # the priority chains of common.go:160-191, one step at a time
    assert inst(7, 8, 9, 10, 11, 1, 5) == (7, POD)
    assert inst(0, 8, 9, 10, 11, 1, 5) == (8, PROCESS)
    assert inst(0, 0, 9, 10, 11, 1, 5) == (9, POD_NODE)
    assert inst(0, 0, 0, 10, 11, 6, 5) == (10, 6)           # l3 device: its own device type
    assert inst(0, 0, 0, 0, 11, 6, -2) == (0, INTERNET)     # EPC_FROM_INTERNET
    assert inst(0, 0, 0, 0, 11, 6, 5) == (11, IP)
    assert svc(1, 2, 3, 4, 5, 6, 7, 1, 130, 5) == (1, CUSTOM)
    assert svc(0, 2, 3, 4, 5, 6, 7, 1, 130, 5) == (2, POD_SERVICE)
    assert svc(0, 0, 3, 4, 5, 6, 7, 1, 130, 5) == (3, 130)  # pod group: its own pod_group_type
    assert svc(0, 0, 0, 4, 5, 6, 7, 1, 130, 5) == (4, PROCESS)
    assert svc(0, 0, 0, 0, 5, 6, 7, 1, 130, 5) == (5, POD_CLUSTER)
    assert svc(0, 0, 0, 0, 0, 6, 7, 1, 130, 5) == (6, 1)
    assert svc(0, 0, 0, 0, 0, 0, 7, 1, 130, -2) == (0, INTERNET)
    assert svc(0, 0, 0, 0, 0, 0, 7, 1, 130, 5) == (7, IP)
    ps = D["DEVICE_TYPE_POD_SERVICE"]
    assert ps == 11 and (D["DEVICE_TYPE_VM"], D["DEVICE_TYPE_HOST_DEVICE"]) == (1, 6)  # custom service chost / host
    assert l.oracle_kat_is_pod_service_ip(ps, 0, 0) == 1 and l.oracle_kat_is_pod_service_ip(1, 0, 0) == 0
    assert l.oracle_kat_is_pod_service_ip(1, 3, 0) == 1 and l.oracle_kat_is_pod_service_ip(1, 0, 4) == 1
    # dictionary snapshot enums the builder switches on (include/dfgpu.h: dfgpu_service.type / protocol, dfgpu_cidr.type)
    assert g["trident.proto:enums"]["ServiceType"] == {"POD_SERVICE_NODE": 1, "POD_SERVICE_POD_GROUP": 2,
                                                         "POD_SERVICE_IP": 3, "LB_SERVICE": 4, "CUSTOM_SERVICE": 7}
    assert g["trident.proto:enums"]["ServiceProtocol"] == {"ANY": 0, "TCP_SERVICE": 1, "UDP_SERVICE": 2}
    assert g["trident.proto:enums"]["CidrType"] == {"WAN": 1, "LAN": 2}


def test_exception_tables_header_regenerates_from_reference():
    """include/dfgpu_exception_tables.h is read by the product, the oracle and the second-opinion tests alike:
    it must be exactly what tools/gen_exception_tables.py produces from the reference's exception_desc.go.
    Runs where the reference tree is mounted (the build container); skipped on the GPU box."""
    import os
    import subprocess
    import sys
    import pytest
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "server/ingester/flow_log/log_data/exception_desc.go")):
        pytest.skip("reference tree not present")
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_exception_tables.py"), ref],
                         capture_output=True, text=True, check=True).stdout
    committed = open(os.path.join(root, "include", "dfgpu_exception_tables.h"), encoding="utf-8").read()
    assert out == committed, "include/dfgpu_exception_tables.h differs from the generator's output"
