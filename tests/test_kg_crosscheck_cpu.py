"""SmartEncoding cross-check ("a second independent Python restatement of A5 on small inputs"):
the oracle's 38 KnowledgeGraph columns + gprocess ids of l7_flow_log and l4_flow_log against
tests/kg_second_opinion.py — a restatement written straight from the Go source, sharing no code with oracle/ —
on generated batches over the synthetic dictionary (WAN / LAN CIDRs, duplicate MACs, pods known only through
pod_ips, gprocess infos, agent addresses, pod / custom services, IPv6).  Inputs are decoded with Google's protobuf
runtime, not with the oracle's decoder."""
import ctypes as C

import numpy as np
import pytest

import kg_second_opinion as K
import pb_schema as P
from deepflow_b200 import abi

KG_COLS = ["region_id", "az_id", "host_id", "l3_device_type", "l3_device_id", "pod_node_id", "pod_ns_id",
           "pod_group_id", "pod_id", "pod_cluster_id", "l3_epc_id", "epc_id", "subnet_id", "service_id",
           "auto_instance_id", "auto_instance_type", "auto_service_id", "auto_service_type", "tag_source"]


def _records(wl):
    out = []
    for f in wl.frames:
        o, end = int(f["offset"]), int(f["offset"]) + int(f["len"])
        while o < end:
            n = int.from_bytes(bytes(wl.data[o:o + 4]), "little")
            out.append(bytes(wl.data[o + 4:o + 4 + n]))
            o += 4 + n
    return out


def _platform(synth):
    from deepflow_b200.synth import lib
    snap = lib().synth_dict(synth._h).contents
    return K.Platform(snap)


def _check(cols, n, want_rows, extra=()):
    by = {c.name: c for c in cols}
    problems = []
    for name in [f"{c}_{s}" for c in KG_COLS for s in (0, 1)] + list(extra):
        got = by[name].values
        for r in range(n):
            w = want_rows[r][name]
            g = int(got[r])
            if by[name].ch_type == abi.CH_INT32:
                w = K._s32(w)
            if g != w:
                problems.append(f"row {r} {name}: oracle {g} != second opinion {w}")
                if len(problems) > 20:
                    return problems
    return problems


@pytest.mark.parametrize("seed,nrec", [(11, 3000), (12, 1500)])
def test_l7_knowledge_graph_columns(synth, built, seed, nrec):
    from oracle_binding import Oracle
    pd = _platform(synth)
    wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, nrec, seed=seed)
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    try:
        ora.submit(wl)
        cols, n = ora.columns(abi.TABLE_L7_FLOW_LOG)
        recs = _records(wl)
        assert n == len(recs) == nrec
        cls = P.classes()["AppProtoLogsData"]
        want = []
        for raw in recs:
            m = cls()
            m.ParseFromString(raw)
            b = m.base
            vtap = b.vtap_id & 0xFFFF
            o = K.fill(pd, b.is_ipv6 == 1, (b.is_vip_interface_src == 1, b.is_vip_interface_dst == 1),
                       (b.l3_epc_id_src, b.l3_epc_id_dst), (b.ip_src, b.ip_dst), (b.ip6_src, b.ip6_dst),
                       (b.mac_src, b.mac_dst), (b.gpid_0, b.gpid_1), vtap, (b.pod_id_0, b.pod_id_1),
                       b.port_dst & 0xFFFF, b.tap_side, b.protocol & 0xFF)
            for s, pid, gp in ((0, b.process_id_0, b.gpid_0), (1, b.process_id_1, b.gpid_1)):
                o[f"gprocess_id_{s}"] = gp
                if pid != 0 and gp == 0:  # l7_flow_log.go:589-596
                    o[f"gprocess_id_{s}"] = pd.vtap_pid.get((vtap << 32) | pid, 0)
                    o[f"tag_source_{s}"] |= K.TS_PROCESSID
            want.append(o)
        problems = _check(cols, n, want, extra=("gprocess_id_0", "gprocess_id_1"))
        assert not problems, "\n".join(problems)
        # the batch must actually exercise the interesting paths
        ts = np.concatenate([c.values for c in cols if c.name in ("tag_source_0", "tag_source_1")])
        for bit in (K.TS_PODID, K.TS_MAC, K.TS_EPCIP, K.TS_PEER, K.TS_AGENT, K.TS_PROCESSID) + \
                ((K.TS_GPID,) if nrec >= 3000 else ()):
            assert (ts & bit).any(), f"tag_source bit {bit} never set"
        assert (next(c for c in cols if c.name == "auto_service_type_1").values == 104).any()  # custom services
    finally:
        ora.close()


def test_l4_knowledge_graph_columns(synth, built):
    from oracle_binding import Oracle
    pd = _platform(synth)
    wl = synth.gen(abi.DFGPU_MSG_TAGGEDFLOW, 1500, seed=21)
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    try:
        ora.submit(wl)
        cols, n = ora.columns(abi.TABLE_L4_FLOW_LOG)
        recs = _records(wl)
        assert n == len(recs) == 1500
        cls = P.classes()["TaggedFlow"]
        want = []
        for raw in recs:
            m = cls()
            m.ParseFromString(raw)
            f = m.flow
            k, a, b = f.flow_key, f.metrics_peer_src, f.metrics_peer_dst
            o = K.fill(pd, f.eth_type == 0x86DD, (a.is_vip_interface == 1, b.is_vip_interface == 1),
                       (K.marshal_epc(a.l3_epc_id), K.marshal_epc(b.l3_epc_id)), (k.ip_src, k.ip_dst),
                       (k.ip6_src, k.ip6_dst), (k.mac_src, k.mac_dst), (a.gpid, b.gpid), k.vtap_id & 0xFFFF, (0, 0),
                       k.port_dst & 0xFFFF, f.tap_side, k.proto & 0xFF)  # FillL4 :730-745
            want.append(o)
        problems = _check(cols, n, want)
        assert not problems, "\n".join(problems)
    finally:
        ora.close()


def test_l7_service_lookups_targeted(synth, built):
    """Generated traffic rarely lands on a pod-service key.  Records are re-addressed so that the server side
    hits every service table on purpose (pod-group+port, node cluster+port, cluster ip+port, the custom-service
    priority list with and without port / epc), re-serialised with Google's protobuf runtime, and compared again."""
    from oracle_binding import Oracle
    from deepflow_b200.synth import Workload
    pd = _platform(synth)
    base = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 1200, seed=31)
    cls = P.classes()["AppProtoLogsData"]
    msgs = []
    for raw in _records(base):
        m = cls()
        m.ParseFromString(raw)
        msgs.append(m)
    ipproto = {0: 47, 1: 6, 2: 17}  # ServiceProtocol -> an IP protocol that maps to it (47: neither TCP nor UDP)
    pods_by_group, node_infos = {}, []
    for pid, inf in pd.pod.items():
        pods_by_group.setdefault(inf.pod_group, pid)
    for (epc, ip), inf in pd.ip4.items():
        if inf.pod_id == 0 and inf.pod_node != 0 and epc > 0:
            node_infos.append((epc, ip, inf))
    plans = []
    for k in list(pd.ps_group_port)[:150]:
        g, sp, port = k >> 32, (k >> 16) & 0xFFFF, k & 0xFFFF
        if g in pods_by_group:
            plans.append(("group", pods_by_group[g], sp, port))
    for epc, ip, inf in node_infos[:150]:
        for k in pd.ps_cluster_port:
            if k >> 32 == inf.pod_cluster:
                plans.append(("node", epc, ip, (k >> 16) & 0xFFFF, k & 0xFFFF))
                break
    for sp in (0, 1, 2):
        for k in list(pd.ps_ip4[sp])[:60] + list(pd.cs_ip4_port[sp])[:60]:
            plans.append(("ip", K._s32(k >> 48 if (k >> 48) < 0x8000 else (k >> 48) - 0x10000), (k >> 16) & 0xFFFFFFFF, sp,
                          k & 0xFFFF))
    for k in list(pd.cs_ip4)[:60]:
        plans.append(("ip", K._s32(k >> 48 if (k >> 48) < 0x8000 else (k >> 48) - 0x10000), (k >> 16) & 0xFFFFFFFF, 1, 4321))
    for pid in list(pd.cs_pod)[:60]:
        plans.append(("pod", pid))
    assert len(plans) > 200, len(plans)
    used = []
    for m, plan in zip(msgs, plans):
        b = m.base
        b.is_ipv6, b.tap_side, b.mac_dst, b.gpid_1, b.process_id_1, b.is_vip_interface_dst = 0, 0, 0, 0, 0, 0
        if plan[0] == "group":
            b.pod_id_1, b.protocol, b.port_dst = plan[1], ipproto[plan[2]], plan[3]
        elif plan[0] == "node":
            b.pod_id_1, b.l3_epc_id_dst, b.ip_dst, b.protocol, b.port_dst = 0, plan[1], plan[2], ipproto[plan[3]], plan[4]
        elif plan[0] == "ip":
            b.pod_id_1, b.l3_epc_id_dst, b.ip_dst, b.protocol, b.port_dst = 0, max(plan[1], 1), plan[2], ipproto[plan[3]], plan[4]
        else:
            b.pod_id_1 = plan[1]
        used.append(m)
    buf = bytearray()
    for m in used:
        raw = m.SerializeToString()
        buf += len(raw).to_bytes(4, "little") + raw
    fr = base.frames[:1].copy()
    fr["offset"], fr["len"] = 0, len(buf)
    wl = Workload(np.frombuffer(bytes(buf), np.uint8).copy(), fr, None, base.msg_type)
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    try:
        ora.submit(wl)
        cols, n = ora.columns(abi.TABLE_L7_FLOW_LOG)
        assert n == len(used) and ora.counters().error_count == 0
        want = []
        for m in used:
            b = m.base
            vtap = b.vtap_id & 0xFFFF
            o = K.fill(pd, b.is_ipv6 == 1, (b.is_vip_interface_src == 1, b.is_vip_interface_dst == 1),
                       (b.l3_epc_id_src, b.l3_epc_id_dst), (b.ip_src, b.ip_dst), (b.ip6_src, b.ip6_dst),
                       (b.mac_src, b.mac_dst), (b.gpid_0, b.gpid_1), vtap, (b.pod_id_0, b.pod_id_1),
                       b.port_dst & 0xFFFF, b.tap_side, b.protocol & 0xFF)
            for s, pid, gp in ((0, b.process_id_0, b.gpid_0), (1, b.process_id_1, b.gpid_1)):
                if pid != 0 and gp == 0:
                    o[f"tag_source_{s}"] |= K.TS_PROCESSID
            want.append(o)
        problems = _check(cols, n, want)
        assert not problems, "\n".join(problems)
        by = {c.name: c for c in cols}
        assert int((by["service_id_1"].values != 0).sum()) > 100          # pod services found
        assert int((by["auto_service_type_1"].values == 104).sum()) > 30   # custom services found
        assert int((by["auto_service_type_1"].values == 11).sum()) > 60    # pod service as auto service
    finally:
        ora.close()


def test_metrics_document_expand_and_routing(synth, built):
    """flow_metrics twin (A10): Tag.ReadFromPB + DocumentExpand + Tag.TableID restated a second time in Python
    against the oracle — which of the 9 tables every Document goes to (or that it is dropped), and the
    SmartEncoding-derived tag columns of its row."""
    import json
    import os
    from oracle_binding import Oracle
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_columns.json")))
    bits, tables = dict(g["code_bits"]), dict(g["table_codes"])
    pd = _platform(synth)
    wl = synth.gen(abi.DFGPU_MSG_METRICS, 4000, seed=41)
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    try:
        ora.submit(wl)
        cls = P.classes()["Document"]
        meter_of_table = {0: 1, 1: 1, 2: 5, 3: 5, 4: 4}  # network* -> Flow, application* -> App, traffic_policy -> Usage
        rows = {i: [] for i in range(9)}
        dropped = 0
        for raw in _records(wl):
            m = cls()
            m.ParseFromString(raw)
            t = K.read_tag(m.tag.field, m.tag.code, bits, tables)
            mid = m.meter.meter_id
            if not K.doc_expand(pd, t, mid, bits, pd.region_id & 0xFFFF):
                dropped += 1
                continue
            ti = K.table_index(t["code"], (m.flags & 1) == 1, bits, tables)
            # inputs the reference leaves undefined: a meter the table does not store, a per-second
            # traffic_policy document -> counted as dropped by the library
            if ti is None or ti > 8 or meter_of_table[ti % 5] != mid:
                dropped += 1
                continue
            rows[ti].append(t)
        single = {"region_id": "region", "host_id": "host", "l3_device_id": "dev_id", "l3_device_type": "dev_type",
                  "subnet_id": "subnet", "pod_node_id": "pod_node", "pod_ns_id": "pod_ns", "az_id": "az",
                  "pod_group_id": "pod_group", "pod_id": "pod", "pod_cluster_id": "pod_cluster", "service_id": "service",
                  "auto_instance_id": "ai_id", "auto_instance_type": "ai_type", "auto_service_id": "as_id",
                  "auto_service_type": "as_type", "tag_source": "ts", "l3_epc_id": "epc", "gprocess_id": "gpid"}
        problems, compared = [], 0
        for ti in range(9):
            cols, n = ora.columns(abi.TABLE_METRICS_BASE + ti)
            if n != len(rows[ti]):
                problems.append(f"table {ti}: oracle {n} rows != second opinion {len(rows[ti])}")
                continue
            by = {c.name: c for c in cols}
            for name, key in single.items():
                for cname, k in ((name, key), (name + "_0", key), (name + "_1", key + "1")):
                    if cname not in by:
                        continue
                    got = by[cname].values
                    for r, t in enumerate(rows[ti]):
                        w = t[k]
                        if by[cname].ch_type == abi.CH_INT32:
                            w = K._s32(w)
                        compared += 1
                        if int(got[r]) != w:
                            problems.append(f"table {ti} row {r} {cname}: oracle {int(got[r])} != second opinion {w}")
                            if len(problems) > 20:
                                break
            if "server_port" in by:
                for r, t in enumerate(rows[ti]):
                    if int(by["server_port"].values[r]) != t["server_port"]:
                        problems.append(f"table {ti} row {r} server_port")
        assert not problems, "\n".join(problems[:25])
        assert compared > 50_000 and all(len(rows[i]) > 50 for i in range(9) if i != 9)
        assert dropped == ora.counters().drop_count and dropped > 100
    finally:
        ora.close()


def test_metrics_full_rows(synth, built):
    """EVERY column of every row of the 9 flow_metrics tables (tag columns in GenTagColumns order + the meter's
    columns) from the second restatement: Tag.AppendToColumnBlock, the three meters' ReadFromPB + AppendToColumnBlock."""
    import json
    import os
    from oracle_binding import Oracle
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_columns.json")))
    bits, tables = dict(g["code_bits"]), dict(g["table_codes"])
    pd = _platform(synth)
    wl = synth.gen(abi.DFGPU_MSG_METRICS, 3000, seed=43)
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    try:
        ora.submit(wl)
        cls = P.classes()["Document"]
        meter_of_table = {0: 1, 1: 1, 2: 5, 3: 5, 4: 4}
        rows = {i: [] for i in range(9)}
        svc_cache, svc_rows = {}, []
        for fr in wl.frames:
            o, end = int(fr["offset"]), int(fr["offset"]) + int(fr["len"])
            while o < end:
                n = int.from_bytes(bytes(wl.data[o:o + 4]), "little")
                raw = bytes(wl.data[o + 4:o + 4 + n])
                o += 4 + n
                m = cls()
                m.ParseFromString(raw)
                t = K.read_tag(m.tag.field, m.tag.code, bits, tables)
                K.read_tag_rest(t, m.tag.field, int(fr["team_id"]) & 0xFFFF)
                mid = m.meter.meter_id
                if not K.doc_expand(pd, t, mid, bits, pd.region_id & 0xFFFF):
                    continue
                ti = K.table_index(t["code"], (m.flags & 1) == 1, bits, tables)
                if ti is None or ti > 8 or meter_of_table[ti % 5] != mid:
                    continue
                # appServiceTagWrite (unmarshaller.go:309-327): only the two 1m application tables, through an lru
                # with the 1800 s flush timeout (app_service_tag_writer.go:92-113); below capacity the lru is a dict
                if ti in (2, 3) and (t["app_service"] or t["app_instance"]):
                    key = (ti, t["app_service"], t["app_instance"], 1, t["team_id"])
                    old_time = svc_cache.get(key)
                    if old_time is None or ((old_time + 1800) & 0xFFFFFFFF) < m.timestamp:
                        svc_cache[key] = m.timestamp
                        svc_rows.append((m.timestamp, b"application" if ti == 2 else b"application_map", t["app_service"],
                                         t["app_instance"], t["team_id"]))
                row = K.tag_columns(t, m.timestamp, bits)
                meter = K.flow_meter_columns(m.meter.flow) if mid == 1 else \
                    K.app_meter_columns(m.meter.app) if mid == 5 else K.usage_meter_columns(m.meter.usage)
                assert not (set(row) & set(meter)) or mid == 5  # application tables: direction_score only in the meter
                row.update(meter)
                rows[ti].append(row)
        problems, compared = [], 0
        for ti in range(9):
            cols, n = ora.columns(abi.TABLE_METRICS_BASE + ti)
            assert n == len(rows[ti]), (ti, n, len(rows[ti]))
            names = [c.name for c in cols]
            assert len(set(names)) == len(names)
            got = {c.name: c.rows() for c in cols}
            for r, want in enumerate(rows[ti]):
                if r == 0:
                    assert set(want) == set(names), (ti, sorted(set(want) ^ set(names)))
                for name, w in want.items():
                    gv = got[name][r]
                    if isinstance(w, float):
                        ok = np.float64(gv).tobytes() == np.float64(w).tobytes()
                    else:
                        if isinstance(w, int) and w >= 1 << 31 and next(c for c in cols if c.name == name).ch_type == abi.CH_INT32:
                            w -= 1 << 32
                        if name.startswith("l3_epc_id"):
                            w = K._s32(w)
                        ok = gv == w
                    compared += 1
                    if not ok:
                        problems.append(f"table {ti} row {r} {name}: oracle {gv!r} != second opinion {w!r}")
                if len(problems) > 15:
                    break
        assert not problems, "\n".join(problems[:15])
        assert compared > 100_000
        cols, n = ora.columns(abi.TABLE_METRICS_APP_SERVICE)
        got = list(zip(*[c.rows() for c in cols])) if n else []
        assert [c.name for c in cols] == ["time", "table", "app_service", "app_instance", "team_id"]
        assert got == svc_rows and len(svc_rows) > 50, (len(got), len(svc_rows))
    finally:
        ora.close()


def _merge_into(acc, other):
    """Meter.ConcurrentMerge on the protobuf messages themselves (field widths = the wire schema's):
    wrapping sums, `_max` / direction_score keep the maximum, Anomaly.client_ooo / server_ooo are ASSIGNED
    (basic_meter.go:94-113, 277-315, 417-424, 556-577, 707-709; app_meter.go:153-265; usage_meter.go)."""
    from google.protobuf.descriptor import FieldDescriptor as FD
    for fd in acc.DESCRIPTOR.fields:
        a, b = getattr(acc, fd.name), getattr(other, fd.name)
        if fd.message_type is not None:
            _merge_into(a, b)
            continue
        mask = (1 << 32) - 1 if fd.type == FD.TYPE_UINT32 else (1 << 64) - 1
        if fd.name.endswith("_max"):
            v = max(a, b)
        elif fd.name == "direction_score":
            v = max(a & 0xFF, b & 0xFF)
        elif fd.name in ("client_ooo", "server_ooo"):
            v = b
        else:
            v = (a + b) & mask
        setattr(acc, fd.name, v)


@pytest.mark.parametrize("bucket", [60, 86400])
def test_metrics_reduce_second_opinion(synth, built, bucket):
    """The reduce operator (A9): rows of one table grouped by (time bucket, every tag column) in first-appearance
    order, meters folded with ConcurrentMerge in input order, the derived *_fail columns kept from the first row."""
    import copy
    import json
    import os
    from oracle_binding import Oracle
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "metrics_columns.json")))
    bits, tables = dict(g["code_bits"]), dict(g["table_codes"])
    pd = _platform(synth)
    from deepflow_b200.synth import Workload
    base = synth.gen(abi.DFGPU_MSG_METRICS, 2500, seed=47)
    # the same documents three times in one submit: every group gets >= 3 rows (first / last / sum / max rules)
    data = np.concatenate([base.data] * 3)
    frames = np.concatenate([base.frames] * 3).copy()
    for r in range(3):
        frames["offset"][r * len(base.frames):(r + 1) * len(base.frames)] += r * base.data.nbytes
    wl = Workload(data, frames, None, base.msg_type)
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    try:
        ora.submit(wl)
        cls = P.classes()["Document"]
        meter_of_table = {0: 1, 1: 1, 2: 5, 3: 5, 4: 4}
        rows = {i: [] for i in range(9)}
        for fr in wl.frames:
            o, end = int(fr["offset"]), int(fr["offset"]) + int(fr["len"])
            while o < end:
                n = int.from_bytes(bytes(wl.data[o:o + 4]), "little")
                m = cls()
                m.ParseFromString(bytes(wl.data[o + 4:o + 4 + n]))
                o += 4 + n
                t = K.read_tag(m.tag.field, m.tag.code, bits, tables)
                K.read_tag_rest(t, m.tag.field, int(fr["team_id"]) & 0xFFFF)
                mid = m.meter.meter_id
                if not K.doc_expand(pd, t, mid, bits, pd.region_id & 0xFFFF):
                    continue
                ti = K.table_index(t["code"], (m.flags & 1) == 1, bits, tables)
                if ti is None or ti > 8 or meter_of_table[ti % 5] != mid:
                    continue
                rows[ti].append((K.tag_columns(t, m.timestamp, bits), m.meter))
        merged_rows = compared = 0
        for ti in range(9):
            groups, order = {}, []
            for tags, meter in rows[ti]:
                tags = dict(tags, time=tags["time"] // bucket * bucket)
                key = tuple(sorted((k, v) for k, v in tags.items()))
                if key not in groups:
                    groups[key] = [tags, copy.deepcopy(meter), meter]  # tags, accumulator, first row's meter
                    order.append(key)
                else:
                    _merge_into(groups[key][1], meter)
                    merged_rows += 1
            assert ora.metrics_reduce(abi.TABLE_METRICS_BASE + ti, bucket) == 0
            cols, n = ora.columns(abi.TABLE_METRICS_REDUCED)
            assert n == len(order), (ti, n, len(order))
            got = {c.name: c.rows() for c in cols}
            for r, key in enumerate(order):
                tags, acc, first = groups[key]
                mid = meter_of_table[ti % 5]
                if mid == 1:
                    want = K.flow_meter_columns(acc.flow)
                    f1 = K.flow_meter_columns(first.flow)
                    for c in ("client_establish_fail", "server_establish_fail", "tcp_establish_fail", "tcp_transfer_fail",
                              "tcp_rst_fail"):
                        want[c] = f1[c]  # set by ReadFromPB only; ConcurrentMerge does not touch them
                elif mid == 5:
                    want = K.app_meter_columns(acc.app)
                else:
                    want = K.usage_meter_columns(acc.usage)
                want = dict(tags, **want)
                for name, w in want.items():
                    gv = got[name][r]
                    if isinstance(w, float):
                        ok = np.float64(gv).tobytes() == np.float64(w).tobytes()
                    else:
                        if name.startswith("l3_epc_id"):
                            w = K._s32(w)
                        ok = gv == w
                    compared += 1
                    assert ok, f"table {ti} group {r} {name}: oracle {gv!r} != second opinion {w!r}"
        assert compared > 100_000 and merged_rows > 3000, (compared, merged_rows)
    finally:
        ora.close()
