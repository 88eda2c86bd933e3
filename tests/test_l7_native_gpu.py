"""GPU parity of dfgpu_cfg.native_strings (first step): the String and Array(String)
columns of l7_flow_log leave the GPU in ClickHouse native column encoding — per value uvarint(len) + bytes,
what ch-go's ColStr.EncodeColumn appends for every Position — instead of Buf + Position arrays.
Expected bytes = that encoding applied to the ORACLE's containers; every other column, the flow_tag tables,
the counters and the persistent state must be exactly what the container mode produces."""
import pytest

from deepflow_b200 import abi
from deepflow_b200 import synth as S

pytestmark = pytest.mark.gpu

STR_TYPES = (abi.CH_STR, abi.CH_LC_STR_ROWS, abi.CH_ARR_STR, abi.CH_ARR_LC_STR)
TAG_TABLES = [abi.TABLE_L7_CUSTOM_FIELD, abi.TABLE_L7_CUSTOM_FIELD_VALUE, abi.TABLE_L7_APP_SERVICE]


def _compare(eng, ora, wl):
    from parity import compare_counters, compare_tables
    ora.submit(wl)
    b = eng.submit(wl)
    problems = compare_tables(b, ora, TAG_TABLES)
    gcols, gn = b.columns(abi.TABLE_L7_FLOW_LOG)
    ocols, on = ora.columns(abi.TABLE_L7_FLOW_LOG)
    if gn != on or len(gcols) != len(ocols):
        problems.append(f"l7_flow_log shape: {gn}x{len(gcols)} != {on}x{len(ocols)}")
    else:
        n_native = 0
        for g, o in zip(gcols, ocols):
            if g.ch_type in STR_TYPES:
                if gn and not g.native:
                    problems.append(f"{g.name}: not native encoded")
                    continue
                n_native += 1
                if g.name != o.name or g.ch_type != o.ch_type or g.nrows != o.nrows:
                    problems.append(f"{g.name}: header differs from {o.name}")
                elif g.native_bytes() != o.native_bytes():
                    problems.append(f"{g.name}: native stream differs ({len(g.native_bytes())} vs "
                                    f"{len(o.native_bytes())} bytes)")
                elif g.arr_offsets is not None and (g.arr_offsets != o.arr_offsets).any():
                    problems.append(f"{g.name}: array offsets differ")
            else:
                msg = abi.columns_equal(g, o)
                if msg:
                    problems.append(msg)
        if gn and n_native < 26:
            problems.append(f"only {n_native} native string columns")
    problems += compare_counters(b.counters(), ora.counters(),
                                 keys=("raw_count", "count", "error_count", "new_field_count",
                                       "new_field_value_count"))
    if eng.id_counter.value != ora.id_counter.value:
        problems.append(f"id counter {eng.id_counter.value} != {ora.id_counter.value}")
    if eng.trace_idx.value != ora.trace_idx.value:
        problems.append(f"lastTraceIdIndex {eng.trace_idx.value} != {ora.trace_idx.value}")
    b.release()
    return problems


@pytest.mark.parametrize("sizes,seed,flags", [
    ([1], 1, 0), ([33, 5000], 2, 0), ([20000, 20000], 5, S.F_ALL_WIRE), ([20000, 3000], 6, S.F_ALL_WIRE | S.F_CORRUPT),
])
def test_l7_native_strings(synth, built, sizes, seed, flags):
    from test_l7_gpu import fresh
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = 1 << 21
    cfg.native_strings = 1
    eng, ora = fresh(synth, cfg)
    try:
        for k, n in enumerate(sizes):
            wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, n, seed=seed * 100 + k, flags=flags)
            problems = _compare(eng, ora, wl)
            assert not problems, f"batch {k}:\n" + "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()


def test_l7_native_long_values(synth, built):
    """Values of 128+ and 16384+ bytes: 2- and 3-byte length prefixes, per-row strings and array elements."""
    import numpy as np
    from test_framing_gpu import _records, _varint
    from test_l7_gpu import fresh
    from deepflow_b200.synth import Workload
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = 1 << 21
    cfg.native_strings = 1
    eng, ora = fresh(synth, cfg)
    try:
        base = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 300, seed=99)
        rng = np.random.default_rng(3)
        buf = bytearray()
        for j, (_, p) in enumerate(_records(base)):
            # append an ext_info occurrence (AppProtoLogsData.ext_info = field 15) carrying one long attribute
            # name/value pair (ExtendedInfo.attribute_names = 16, attribute_values = 17): sub-messages merge
            n = [0, 130, 200, 16384, 20000][j % 5]
            if n:
                name = bytes(rng.integers(97, 123, n, dtype=np.uint8))
                val = bytes(rng.integers(97, 123, n + 7, dtype=np.uint8))
                ext = _varint(16 << 3 | 2) + _varint(len(name)) + name + _varint(17 << 3 | 2) + _varint(len(val)) + val
                p = p + _varint(15 << 3 | 2) + _varint(len(ext)) + ext
            buf += len(p).to_bytes(4, "little") + p
        fr = base.frames[:1].copy()
        fr["offset"], fr["len"] = 0, len(buf)
        wl = Workload(np.frombuffer(bytes(buf), np.uint8).copy(), fr, None, base.msg_type)
        problems = _compare(eng, ora, wl)
        assert not problems, "\n".join(problems[:20])
        assert ora.counters().error_count == 0
    finally:
        eng.close()
        ora.close()
