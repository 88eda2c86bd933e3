"""GPU soak tests at the edges of the device-resident state.

1. flow_tag set far beyond its capacity: 10x more distinct keys than dfgpu_cfg.flow_tag_cache_max_size,
   over many batches.  The reference bounds its lru by evicting the least recently used key
   (flow_tag_writer.go:161-210, lru.NewU128LRU); the device set evicts oldest-first when it is more than
   half full (tagset.cu).  The two are not row-for-row equal once they evict, so the test
   checks what the ingest contract needs: no batch fails, the main table stays bit exact, every key the
   reference writes is written by the device at least once (ReplacingMergeTree makes extra rows harmless,
   missing ones are not), and the set never grows past its capacity.
2. A 1 M pod dictionary (BASELINE.json configs[4]) through L7 parity: the SmartEncoding tables at the size
   the high-cardinality configuration uses.
"""
import pytest

from deepflow_b200 import abi
from deepflow_b200 import synth as S

pytestmark = pytest.mark.gpu

MAIN = [abi.TABLE_L7_FLOW_LOG]
TAG_TABLES = [abi.TABLE_L7_CUSTOM_FIELD, abi.TABLE_L7_CUSTOM_FIELD_VALUE]


def tag_rows(cols):
    """Row tuples of a flow_tag table without the time column (the write time is not part of the key)."""
    keep = [c for c in cols if c.name != "time"]
    per_col = [c.rows() for c in keep]
    return set(zip(*[[bytes(v) if isinstance(v, (bytes, bytearray)) else v for v in col] for col in per_col])) if keep else set()


def test_flow_tag_set_10x_capacity(synth, built):
    from deepflow_b200.native import Engine
    from oracle_binding import Oracle
    from parity import compare_tables
    cap = 1 << 12
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = cap
    eng = Engine(0, cfg)
    eng.load_synth(synth)
    ora = Oracle(cfg)
    ora.load_synth(synth)
    dev_keys = {t: set() for t in TAG_TABLES}
    ora_keys = {t: set() for t in TAG_TABLES}
    evicted = unrecorded = 0
    try:
        for k in range(12):
            wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 6000, seed=9000 + k, flags=S.F_HIGH_CARD)
            ora.submit(wl)
            b = eng.submit(wl)
            c = b.counters()
            assert c.error_count == 0 and c.count == wl.nrecords, f"batch {k}: {c.as_dict()}"
            problems = compare_tables(b, ora, MAIN)
            assert not problems, f"batch {k}:\n" + "\n".join(problems[:10])
            for t in TAG_TABLES:
                dev_keys[t] |= tag_rows(b.columns(t)[0])
                ora_keys[t] |= tag_rows(ora.columns(t)[0])
            evicted = max(evicted, c.flow_tag_evicted)
            unrecorded += c.flow_tag_unrecorded
            b.release()
        n_distinct = sum(len(v) for v in ora_keys.values())
        assert n_distinct >= 10 * cap, f"workload too small for the test: {n_distinct} distinct keys, capacity {cap}"
        assert evicted > 0, "the set was driven 10x past its capacity and never evicted"
        for t in TAG_TABLES:
            missing = ora_keys[t] - dev_keys[t]
            assert not missing, f"table {t}: {len(missing)} keys the reference wrote were never written, e.g. {list(missing)[:3]}"
    finally:
        eng.close()
        ora.close()


def test_l7_parity_1m_pod_dictionary(built):
    from deepflow_b200.native import Engine
    from oracle_binding import Oracle
    from parity import compare_counters, compare_tables
    big = S.Synth(n_pods=1_000_000)
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = 1 << 21
    eng = Engine(0, cfg)
    eng.load_synth(big)
    ora = Oracle(cfg)
    ora.load_synth(big)
    try:
        for k, (n, flags) in enumerate(((30000, 0), (30000, S.F_HIGH_CARD))):
            wl = big.gen(abi.DFGPU_MSG_PROTOCOLLOG, n, seed=700 + k, flags=flags)
            ora.submit(wl)
            b = eng.submit(wl)
            problems = compare_tables(b, ora, MAIN + TAG_TABLES + [abi.TABLE_L7_APP_SERVICE])
            problems += compare_counters(b.counters(), ora.counters(),
                                         keys=("raw_count", "count", "error_count", "new_field_count", "new_field_value_count"))
            b.release()
            assert not problems, f"batch {k}:\n" + "\n".join(problems[:20])
        wl = big.gen(abi.DFGPU_MSG_TAGGEDFLOW, 30000, seed=710)
        ora.submit(wl)
        b = eng.submit(wl)
        problems = compare_tables(b, ora, [abi.TABLE_L4_FLOW_LOG])
        b.release()
        assert not problems, "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()
