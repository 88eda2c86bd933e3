"""flow_tag discovery a second time: L7FlowLog.GenerateNewFlowTags
(flow_log/log_data/l7_flow_log.go:646-745), FlowTag.AppendToColumnBlock (flow_tag/flow_tag_column_block.go:79-103)
and AppServiceTagWriter.Write (flow_tag/app_service_tag_writer.go:92-113, call site decoder.go:464-472) restated
in Python with plain dicts as the lru caches (below capacity an lru IS a dict), fed with the oracle's own
l7_flow_log rows, against the oracle's custom_field / custom_field_value / app_service tables — row for row, in
order, across consecutive batches (the caches persist)."""
import pytest

from deepflow_b200 import abi

TIMEOUT = 1800  # FlowTagCacheFlushTimeout, ingester/config/config.go:59
EXTRA = [b"app_service", b"endpoint", b"app_instance"]


def _u32(v):
    return v & 0xFFFFFFFF


class SecondOpinion:
    def __init__(self):
        self.value_cache, self.field_cache, self.svc_cache = {}, {}, {}

    @staticmethod
    def _add_or_get(cache, key, time):
        """lru AddOrGet + the caller's refresh: True when the row must be emitted."""
        old = cache.get(key)
        if old is None:
            cache[key] = time
            return True
        if _u32(old + TIMEOUT) >= time:
            return False
        cache[key] = time
        return True

    def row(self, r, fields, values, services):
        """r: one l7_flow_log row as a dict; appends the emitted rows."""
        epcs, nss = [r["l3_epc_id_0"], r["l3_epc_id_1"]], [r["pod_ns_id_0"], r["pod_ns_id_1"]]
        sides = 1 if epcs[0] == epcs[1] and nss[0] == nss[1] else 2
        time = _u32(r["end_time"] // 1_000_000)
        names = list(r["attribute_names"]) + EXTRA
        vals = list(r["attribute_values"]) + [r["app_service"], r["endpoint"], r["app_instance"]]
        n = min(len(names), len(vals))
        for s in range(sides):
            info = (b"l7_flow_log", epcs[s], nss[s], 1, r["team_id"])  # table, vpc, pod_ns, org, team
            for i in range(n):
                if not self._add_or_get(self.value_cache, info + (0, names[i], vals[i]), time):
                    continue
                values.append((time, b"l7_flow_log", epcs[s], nss[s], b"tag", names[i], b"string", r["team_id"], vals[i], 1))
                if i >= len(r["attribute_names"]):
                    continue
                if not self._add_or_get(self.field_cache, info + (0, names[i], b""), time):
                    continue
                fields.append((time, b"l7_flow_log", epcs[s], nss[s], b"tag", names[i], b"string", r["team_id"]))
            for name in r["metrics_names"]:
                if not self._add_or_get(self.field_cache, info + (1, name, b""), time):
                    continue
                fields.append((time, b"l7_flow_log", epcs[s], nss[s], b"metrics", name, b"float", r["team_id"]))
        if r["app_service"] or r["app_instance"]:
            if self._add_or_get(self.svc_cache, (b"l7_flow_log", r["app_service"], r["app_instance"], 1, r["team_id"]), r["time"]):
                services.append((r["time"], b"l7_flow_log", r["app_service"], r["app_instance"], r["team_id"]))


def _rows(cols, names):
    by = {c.name: c.rows() for c in cols if c.name in names}
    n = len(next(iter(by.values()))) if by else 0
    return [{k: by[k][i] for k in names} for i in range(n)]


@pytest.mark.parametrize("sizes,seed", [([4000, 4000, 2000], 3), ([1, 700], 9)])
def test_flow_tag_tables_match_second_opinion(synth, built, sizes, seed):
    from oracle_binding import Oracle
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = 1 << 21
    ora = Oracle(cfg)
    ora.load_synth(synth)
    so = SecondOpinion()
    need = ["l3_epc_id_0", "l3_epc_id_1", "pod_ns_id_0", "pod_ns_id_1", "end_time", "time", "team_id",
            "attribute_names", "attribute_values", "metrics_names", "app_service", "endpoint", "app_instance"]
    try:
        emitted = [0, 0, 0]
        for k, n in enumerate(sizes):
            ora.submit(synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, n, seed=seed * 100 + k))
            cols, nrows = ora.columns(abi.TABLE_L7_FLOW_LOG)
            fields, values, services = [], [], []
            for r in _rows(cols, need):
                so.row(r, fields, values, services)
            for tid, want, names in (
                    (abi.TABLE_L7_CUSTOM_FIELD, fields,
                     ["time", "table", "vpc_id", "pod_ns_id", "field_type", "field_name", "field_value_type", "team_id"]),
                    (abi.TABLE_L7_CUSTOM_FIELD_VALUE, values,
                     ["time", "table", "vpc_id", "pod_ns_id", "field_type", "field_name", "field_value_type", "team_id",
                      "field_value", "count"]),
                    (abi.TABLE_L7_APP_SERVICE, services, ["time", "table", "app_service", "app_instance", "team_id"])):
                tcols, tn = ora.columns(tid)
                got = [tuple(d[c] for c in names) for d in _rows(tcols, names)] if tn else []
                assert tn == len(want), f"batch {k} table {tid}: oracle {tn} rows != second opinion {len(want)}"
                for i, (g, w) in enumerate(zip(got, want)):
                    assert g == w, f"batch {k} table {tid} row {i}: oracle {g} != second opinion {w}"
                emitted[(abi.TABLE_L7_CUSTOM_FIELD, abi.TABLE_L7_CUSTOM_FIELD_VALUE, abi.TABLE_L7_APP_SERVICE).index(tid)] += tn
        if sum(sizes) > 5000:
            assert all(e > 10 for e in emitted), emitted
    finally:
        ora.close()
