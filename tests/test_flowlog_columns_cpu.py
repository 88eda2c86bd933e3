"""The l4_flow_log / l7_flow_log column blocks: names, ORDER and ch-go container kinds of every
column against the lists MACHINE-EXTRACTED from the reference's own ToInput functions
(tests/golden/flowlog_columns.json, tools/make_flowlog_golden.py; l4_column_block.go:822-833,
l7_column_block.go:114-208,304-345).  Checked on the oracle here; the GPU library is compared with the oracle
column by column in the -m gpu tests (same order, same kinds, same bytes)."""
import json
import os

import pytest

from deepflow_b200 import abi

_KIND = {  # ch-go container -> dfgpu_ch_type kinds that may carry it
    "ColUInt8": {abi.CH_UINT8}, "ColUInt16": {abi.CH_UINT16}, "ColUInt32": {abi.CH_UINT32},
    "ColUInt64": {abi.CH_UINT64}, "ColInt32": {abi.CH_INT32}, "ColInt64": {abi.CH_INT64},
    "ColFloat64": {abi.CH_FLOAT64}, "ColDateTime": {abi.CH_DATETIME}, "ColDateTime64": {abi.CH_DATETIME64},
    "ColIPv4": {abi.CH_IPV4}, "ColIPv6": {abi.CH_IPV6}, "ColStr": {abi.CH_STR},
    # dictionary-coded on the device where the value set is closed, else handed over as per-row values
    "ColLowCardinality[string]": {abi.CH_LC_STR, abi.CH_LC_STR_ROWS},
    # the struct type does not say whether the element is LowCardinality (proto.NewArray(...LowCardinality()))
    "ColArr[string]": {abi.CH_ARR_STR, abi.CH_ARR_LC_STR},
    "ColArr[uint16]": {abi.CH_ARR_UINT16}, "ColArr[float64]": {abi.CH_ARR_FLOAT64},
    "ColNullable[uint8]": {abi.CH_NULLABLE_UINT8}, "ColNullable[int32]": {abi.CH_NULLABLE_INT32},
    "ColNullable[int64]": {abi.CH_NULLABLE_INT64}, "ColNullable[uint64]": {abi.CH_NULLABLE_UINT64},
}


@pytest.mark.parametrize("table,tid,msg,ncols", [
    ("l4_flow_log", abi.TABLE_L4_FLOW_LOG, abi.DFGPU_MSG_TAGGEDFLOW, 154),
    ("l7_flow_log", abi.TABLE_L7_FLOW_LOG, abi.DFGPU_MSG_PROTOCOLLOG, 119),
    # flow_tag_column_block.go:50-67, app_service_tag_column_block.go
    ("flow_tag.custom_field", abi.TABLE_L7_CUSTOM_FIELD, abi.DFGPU_MSG_PROTOCOLLOG, 8),
    ("flow_tag.custom_field_value", abi.TABLE_L7_CUSTOM_FIELD_VALUE, abi.DFGPU_MSG_PROTOCOLLOG, 10),
    ("flow_tag.app_service", abi.TABLE_L7_APP_SERVICE, abi.DFGPU_MSG_PROTOCOLLOG, 5),
])
def test_column_block_matches_reference_toinput(synth, built, table, tid, msg, ncols):
    from oracle_binding import Oracle
    with open(os.path.join(os.path.dirname(__file__), "golden", "flowlog_columns.json")) as f:
        want = json.load(f)[table]
    assert len(want) == ncols
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    try:
        ora.submit(synth.gen(msg, 16, seed=3))
        cols, n = ora.columns(tid)
        assert n == 16 or (table.startswith("flow_tag") and n > 0)
        assert [c.name for c in cols] == [w[0] for w in want]
        bad = [(c.name, w[1], c.ch_type) for c, w in zip(cols, want) if c.ch_type not in _KIND[w[1]]]
        assert not bad, bad
    finally:
        ora.close()
