"""GPU parity of dfgpu_cfg.block_lz4 (per-column compressed frames on the device).
Every buffer of every l7_flow_log / l4_flow_log column comes back as a stream of ClickHouse compressed frames
(lz4.cu).  abi.Column decodes the frames — each LZ4 block by two independent decoders (pyarrow's LZ4 codec and a
plain one in abi.py) — and the decoded containers must be byte for byte what the uncompressed mode and the
oracle produce."""
import numpy as np
import pytest

from deepflow_b200 import abi
from deepflow_b200 import synth as S

pytestmark = pytest.mark.gpu


def _engines(synth, native):
    from deepflow_b200.native import Engine
    from oracle_binding import Oracle
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = 1 << 21
    cfg.native_strings = native
    cfg.block_lz4 = 1
    eng = Engine(0, cfg)
    eng.load_synth(synth)
    ora = Oracle(cfg)
    ora.load_synth(synth)
    return eng, ora


@pytest.mark.parametrize("msg,table,sizes,seed,flags", [
    (abi.DFGPU_MSG_PROTOCOLLOG, abi.TABLE_L7_FLOW_LOG, [1, 33], 1, 0),
    (abi.DFGPU_MSG_PROTOCOLLOG, abi.TABLE_L7_FLOW_LOG, [20000, 5000], 2, 0),
    (abi.DFGPU_MSG_PROTOCOLLOG, abi.TABLE_L7_FLOW_LOG, [20000], 3, S.F_ALL_WIRE | S.F_CORRUPT),
    (abi.DFGPU_MSG_TAGGEDFLOW, abi.TABLE_L4_FLOW_LOG, [1, 4097, 20000], 4, 0),
])
def test_lz4_containers_equal_oracle(synth, built, msg, table, sizes, seed, flags):
    """Container mode (Buf + Positions), every buffer compressed: decoded == oracle."""
    from parity import compare_tables
    eng, ora = _engines(synth, 0)
    try:
        for k, n in enumerate(sizes):
            wl = synth.gen(msg, n, seed=seed * 100 + k, flags=flags)
            ora.submit(wl)
            b = eng.submit(wl)
            cols, rows = b.columns(table)
            assert rows == 0 or all(c.compressed_bytes is not None for c in cols), "columns are not compressed"
            problems = compare_tables(b, ora, [table])
            b.release()
            assert not problems, f"batch {k}:\n" + "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()


def test_lz4_native_strings_and_ratio(synth, built):
    """Native string streams compressed; and the frames are smaller than what they carry."""
    eng, ora = _engines(synth, 1)
    try:
        wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 30000, seed=77)
        ora.submit(wl)
        b = eng.submit(wl)
        gcols, gn = b.columns(abi.TABLE_L7_FLOW_LOG)
        ocols, on = ora.columns(abi.TABLE_L7_FLOW_LOG)
        assert gn == on and len(gcols) == len(ocols)
        raw = comp = 0
        for g, o in zip(gcols, ocols):
            if g.ch_type in (abi.CH_STR, abi.CH_LC_STR_ROWS, abi.CH_ARR_STR, abi.CH_ARR_LC_STR):
                assert g.native and g.native_bytes() == o.native_bytes(), g.name
                if g.arr_offsets is not None:
                    assert np.array_equal(g.arr_offsets, o.arr_offsets), g.name
            else:
                assert abi.columns_equal(g, o) is None, abi.columns_equal(g, o)
            comp += g.compressed_bytes
            for a in (g.values, g.buf, g.arr_offsets, g.nulls):
                if a is not None:
                    raw += a.nbytes if g.ch_type != abi.CH_LC_STR or a is not g.values else g.nrows * g.elem_size
        b.release()
        assert comp < 0.75 * raw, f"frames {comp} B for {raw} B of column data"
    finally:
        eng.close()
        ora.close()
