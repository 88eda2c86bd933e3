"""GPU parity: L7 AppProtoLogsData path through the C ABI against
the CPU oracle — all 119 l7_flow_log columns, the two flow_tag tables and the
app_service tag table, bit exact, across consecutive batches (the flow_tag set, the
_id counter and lastTraceIdIndex persist between submits)."""
import pytest

from deepflow_b200 import abi
from deepflow_b200 import synth as S

pytestmark = pytest.mark.gpu

TABLES = [abi.TABLE_L7_FLOW_LOG, abi.TABLE_L7_CUSTOM_FIELD, abi.TABLE_L7_CUSTOM_FIELD_VALUE,
          abi.TABLE_L7_APP_SERVICE]


def fresh(synth, cfg=None):
    from deepflow_b200.native import Engine
    from oracle_binding import Oracle
    eng = Engine(0, cfg)
    eng.load_synth(synth)
    ora = Oracle(cfg)
    ora.load_synth(synth)
    return eng, ora


def run_both(eng, ora, wl):
    from parity import compare_counters, compare_tables
    ora.submit(wl)
    b = eng.submit(wl)
    problems = compare_tables(b, ora, TABLES)
    problems += compare_counters(b.counters(), ora.counters(),
                                 keys=("raw_count", "count", "error_count", "l7_http_count", "l7_dns_count",
                                       "l7_sql_count", "l7_nosql_count", "l7_rpc_count", "l7_mq_count",
                                       "new_field_count", "new_field_value_count"))
    if eng.id_counter.value != ora.id_counter.value:
        problems.append(f"id counter {eng.id_counter.value} != {ora.id_counter.value}")
    if eng.trace_idx.value != ora.trace_idx.value:
        problems.append(f"lastTraceIdIndex {eng.trace_idx.value} != {ora.trace_idx.value}")
    b.release()
    return problems


@pytest.mark.parametrize("sizes,seed,flags", [
    ([1], 1, 0), ([31, 33], 2, 0), ([5000, 5000, 5000], 3, 0), ([60000], 4, 0),
    ([20000, 20000], 5, S.F_ALL_WIRE), ([20000, 3000], 6, S.F_ALL_WIRE | S.F_CORRUPT),
])
def test_l7_parity(synth, built, sizes, seed, flags):
    # the device tag set never evicts; keep the reference's LRUs below their capacity so
    # that eviction (SURVEY §7 "hard parts") does not enter the comparison
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = 1 << 21
    eng, ora = fresh(synth, cfg)
    try:
        for k, n in enumerate(sizes):
            wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, n, seed=seed * 100 + k, flags=flags)
            problems = run_both(eng, ora, wl)
            assert not problems, f"batch {k}:\n" + "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()


def test_l7_trace_index_disabled(synth, built):
    cfg = abi.Cfg()
    cfg.trace_id_index_disabled = 1
    eng, ora = fresh(synth, cfg)
    try:
        wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 2000, seed=77)
        problems = run_both(eng, ora, wl)
        assert not problems, "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()


def test_l7_same_batch_twice_emits_no_new_tags(synth, built):
    """Second pass over identical records: every flow_tag key is cached and fresh."""
    eng, ora = fresh(synth)
    try:
        wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 4000, seed=12)
        assert not run_both(eng, ora, wl)
        ora.submit(wl)
        b = eng.submit(wl)
        c = b.counters()
        assert c.new_field_count == 0 and c.new_field_value_count == 0
        assert ora.counters().new_field_value_count == 0
        b.release()
    finally:
        eng.close()
        ora.close()
