"""The test-side restatement of ch-go's ColStr.EncodeColumn (abi.Column.native_bytes: per value
uvarint(len) + bytes — the ClickHouse native String encoding, SerializationString) against hand-written
vectors and its own decoder, on the oracle's l7_flow_log containers.  It is the expected value of
tests/test_l7_native_gpu.py."""
import numpy as np

from deepflow_b200 import abi


def _col(strings):
    c = abi.Column.__new__(abi.Column)
    c.name, c.ch_type, c.elem_size, c.nrows = "s", abi.CH_STR, 0, len(strings)
    c.values = c.arr_offsets = c.nulls = c.dict = None
    c.native = False
    buf, pos, o = bytearray(), [], 0
    for s in strings:
        buf += s
        pos.append((o, o + len(s)))
        o += len(s)
    c.buf = np.frombuffer(bytes(buf), np.uint8)
    c.pos = np.array(pos, np.int64).reshape(-1, 2)
    return c


def test_known_vectors():
    assert _col([b""]).native_bytes() == b"\x00"
    assert _col([b"a", b"", b"bc"]).native_bytes() == b"\x01a\x00\x02bc"
    s127, s128, s300 = b"x" * 127, b"y" * 128, b"z" * 300
    assert _col([s127]).native_bytes() == b"\x7f" + s127
    assert _col([s128]).native_bytes() == b"\x80\x01" + s128          # 128 = 0x80 0x01
    assert _col([s300]).native_bytes() == b"\xac\x02" + s300          # 300 = 0xAC 0x02
    big = b"q" * 16384
    assert _col([big]).native_bytes()[:3] == b"\x80\x80\x01"          # 16384 = 0x80 0x80 0x01


def test_round_trip_on_oracle_columns(synth, built):
    from oracle_binding import Oracle
    ora = Oracle(abi.Cfg())
    ora.load_synth(synth)
    try:
        ora.submit(synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 3000, seed=5))
        cols, n = ora.columns(abi.TABLE_L7_FLOW_LOG)
        seen = 0
        for c in cols:
            if c.ch_type not in (abi.CH_STR, abi.CH_LC_STR_ROWS, abi.CH_ARR_STR, abi.CH_ARR_LC_STR):
                continue
            enc = c.native_bytes()
            d = abi.Column.__new__(abi.Column)
            d.native, d.buf = True, np.frombuffer(enc, np.uint8)
            assert d.strings() == c.strings(), c.name
            assert len(enc) == int(c.buf.nbytes) + sum(1 if len(s) < 128 else 2 if len(s) < 16384 else 3
                                                       for s in c.strings()), c.name
            seen += 1
        assert seen == 26 and n == 3000
    finally:
        ora.close()


def test_lz4_block_decoder_against_pyarrow():
    """abi.lz4_block_decode (the plain decoder the GPU frame tests use as their second opinion) on blocks produced by
    pyarrow's LZ4 compressor: literals only, long literal / match length extensions, overlapping copies."""
    import os
    import numpy as np
    import pyarrow as pa
    from deepflow_b200 import abi
    rng = np.random.default_rng(7)
    cases = [b"a", b"abcd" * 3, bytes(4096), os.urandom(300), b"0123456789abcdef" * 256,
             (b"x" * 300 + os.urandom(20)) * 5, bytes(rng.integers(0, 4, 4096, dtype=np.uint8)),
             np.arange(512, dtype=np.uint64).tobytes()]
    for data in cases:
        blk = pa.compress(data, codec="lz4_raw", asbytes=True)
        assert abi.lz4_block_decode(blk, len(data)) == data
    import pytest
    with pytest.raises(ValueError):
        abi.lz4_block_decode(pa.compress(b"hello hello hello hello", codec="lz4_raw", asbytes=True), 5)


def test_frame_stream_reader_rejects_malformed():
    import ctypes as C
    import pyarrow as pa
    import pytest
    from deepflow_b200 import abi

    def stream(method, payload, usz, checksum=bytes(16)):
        return checksum + bytes([method]) + (len(payload) + 9).to_bytes(4, "little") + usz.to_bytes(4, "little") + payload

    data = b"flow_log " * 100
    good = stream(0x82, pa.compress(data, codec="lz4_raw", asbytes=True), len(data)) + stream(0x02, b"tail", 4)
    buf = (C.c_uint8 * len(good)).from_buffer_copy(good)
    assert abi.decode_frames(C.addressof(buf), len(good)) == data + b"tail"
    for bad in (stream(0x90, b"zzzz", 4), stream(0x02, b"abc", 4), stream(0x02, b"abcd", 4, checksum=b"\x01" * 16),
                stream(0x82, b"\x00" * 10, 4)):
        b2 = (C.c_uint8 * len(bad)).from_buffer_copy(bad)
        with pytest.raises(ValueError):
            abi.decode_frames(C.addressof(b2), len(bad))
