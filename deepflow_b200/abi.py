"""ctypes mirror of include/dfgpu.h (the C ABI).  Shared by the product binding
(`native.py`), the synthetic generator binding (`synth.py`) and the test-only
oracle binding (`tests/oracle_binding.py`)."""
import ctypes as C

import numpy as np

DFGPU_MSG_METRICS = 3
DFGPU_MSG_TAGGEDFLOW = 4
DFGPU_MSG_PROTOCOLLOG = 5

TABLE_L4_FLOW_LOG = 0
TABLE_L7_FLOW_LOG = 1
TABLE_L7_CUSTOM_FIELD = 2
TABLE_L7_CUSTOM_FIELD_VALUE = 3
TABLE_L7_APP_SERVICE = 4
TABLE_METRICS_BASE = 16
TABLE_METRICS_APP_SERVICE = 25
TABLE_METRICS_REDUCED = 31
TABLE_MAX = 32

CH_UINT8, CH_UINT16, CH_UINT32, CH_UINT64 = 1, 2, 3, 4
CH_INT32, CH_INT64, CH_FLOAT64, CH_IPV4, CH_IPV6 = 5, 6, 7, 8, 9
CH_DATETIME, CH_DATETIME64, CH_STR, CH_LC_STR = 10, 11, 12, 13
CH_ARR_UINT16, CH_ARR_STR, CH_ARR_LC_STR, CH_ARR_FLOAT64 = 14, 15, 16, 17
CH_NULLABLE_UINT8, CH_NULLABLE_INT32, CH_NULLABLE_INT64, CH_NULLABLE_UINT64 = 18, 19, 20, 21
CH_LC_STR_ROWS = 22

_FIXED_DTYPE = {
    CH_UINT8: np.uint8, CH_UINT16: np.uint16, CH_UINT32: np.uint32, CH_UINT64: np.uint64,
    CH_INT32: np.int32, CH_INT64: np.int64, CH_FLOAT64: np.float64, CH_IPV4: np.uint32,
    CH_DATETIME: np.uint32, CH_DATETIME64: np.int64,
    CH_NULLABLE_UINT8: np.uint8, CH_NULLABLE_INT32: np.int32, CH_NULLABLE_INT64: np.int64,
    CH_NULLABLE_UINT64: np.uint64, CH_ARR_UINT16: np.uint16, CH_ARR_FLOAT64: np.float64,
}


class Pos(C.Structure):
    _fields_ = [("start", C.c_int64), ("end", C.c_int64)]


class Col(C.Structure):
    _fields_ = [
        ("name", C.c_char_p), ("ch_type", C.c_int32), ("elem_size", C.c_int32),
        ("nrows", C.c_uint64), ("data", C.c_void_p), ("data_bytes", C.c_uint64),
        ("pos", C.c_void_p), ("arr_offsets", C.c_void_p), ("nelems", C.c_uint64),
        ("nulls", C.c_void_p), ("dict_data", C.c_void_p), ("dict_pos", C.c_void_p),
        ("dict_n", C.c_uint64),
        # dfgpu_cfg.block_lz4 (include/dfgpu.h): the buffers are streams of ClickHouse compressed frames
        ("compressed", C.c_uint32), ("reserved0", C.c_uint32), ("data_raw_bytes", C.c_uint64),
        ("pos_bytes", C.c_uint64), ("arr_offsets_bytes", C.c_uint64), ("nulls_bytes", C.c_uint64),
    ]


class Frame(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("len", C.c_uint32), ("vtap_id", C.c_uint16),
                ("org_id", C.c_uint16), ("team_id", C.c_uint32)]


FRAME_DTYPE = np.dtype([("offset", "<u8"), ("len", "<u4"), ("vtap_id", "<u2"), ("org_id", "<u2"),
                        ("team_id", "<u4")], align=True)
assert FRAME_DTYPE.itemsize == C.sizeof(Frame) == 24


class Ip(C.Structure):
    _fields_ = [("is_v6", C.c_uint8), ("valid", C.c_uint8), ("pad", C.c_uint8 * 2),
                ("ip4", C.c_uint32), ("ip6", C.c_uint8 * 16)]


class DictSnapshot(C.Structure):
    _fields_ = [
        ("version", C.c_uint64), ("region_id", C.c_uint32), ("analyzer_id", C.c_uint32),
        ("interfaces", C.c_void_p), ("n_interfaces", C.c_uint64),
        ("ip_resources", C.c_void_p), ("n_ip_resources", C.c_uint64),
        ("cidrs", C.c_void_p), ("n_cidrs", C.c_uint64),
        ("gprocesses", C.c_void_p), ("n_gprocesses", C.c_uint64),
        ("vtap_ips", C.c_void_p), ("n_vtap_ips", C.c_uint64),
        ("pod_ips", C.c_void_p), ("n_pod_ips", C.c_uint64),
        ("services", C.c_void_p), ("n_services", C.c_uint64),
        ("u32_pool", C.c_void_p), ("n_u32_pool", C.c_uint64),
        ("ip_pool", C.c_void_p), ("n_ip_pool", C.c_uint64),
    ]


class GeoEntry(C.Structure):
    _fields_ = [("ip_start", C.c_uint32), ("ip_end", C.c_uint32), ("region", C.c_uint8),
                ("pad", C.c_uint8 * 3)]


class Cfg(C.Structure):
    _fields_ = [("max_batch_bytes", C.c_uint64), ("max_batch_records", C.c_uint32),
                ("flow_tag_cache_flush_timeout", C.c_uint32), ("flow_tag_cache_max_size", C.c_uint32),
                ("trace_id_index_disabled", C.c_uint32), ("native_strings", C.c_uint32),
                ("rollup_tables", C.c_uint32), ("rollup_max_groups", C.c_uint32), ("block_lz4", C.c_uint32),
                ("reserved", C.c_uint32 * 3)]


class RollupRows(C.Structure):
    _fields_ = [("nrows", C.c_uint64), ("tag_cols", C.c_void_p), ("ntag", C.c_int32), ("raw", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("raw_count", C.c_int64), ("count", C.c_int64), ("error_count", C.c_int64),
                ("drop_count", C.c_int64), ("l7_http_count", C.c_int64), ("l7_dns_count", C.c_int64),
                ("l7_sql_count", C.c_int64), ("l7_nosql_count", C.c_int64), ("l7_rpc_count", C.c_int64),
                ("l7_mq_count", C.c_int64), ("new_field_count", C.c_int64),
                ("new_field_value_count", C.c_int64), ("rows", C.c_int64 * TABLE_MAX),
                ("device_us", C.c_double), ("flow_tag_unrecorded", C.c_int64), ("flow_tag_evicted", C.c_int64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_
             if k not in ("rows", "device_us", "flow_tag_unrecorded", "flow_tag_evicted")}
        d["rows"] = {i: self.rows[i] for i in range(TABLE_MAX) if self.rows[i]}
        return d


def _np_from(ptr, nbytes, dtype):
    if not ptr or nbytes == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def lz4_block_decode(src: bytes, raw_size: int) -> bytes:
    """Plain LZ4 block decoder (the format of lz4_Block_format.md), written for the tests: independent of the
    compressor in lz4.cu, checks every bound."""
    out = bytearray()
    i, n = 0, len(src)
    while True:
        if i >= n:
            raise ValueError("lz4: truncated block")
        tok = src[i]
        i += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = src[i]
                i += 1
                lit += b
                if b != 255:
                    break
        if i + lit > n:
            raise ValueError("lz4: literals past the end")
        out += src[i:i + lit]
        i += lit
        if i == n:
            break  # the last sequence has no match
        off = src[i] | src[i + 1] << 8
        i += 2
        if off == 0 or off > len(out):
            raise ValueError(f"lz4: offset {off} outside the {len(out)} bytes produced")
        ml = tok & 15
        if ml == 15:
            while True:
                b = src[i]
                i += 1
                ml += b
                if b != 255:
                    break
        ml += 4
        start = len(out) - off
        if off >= ml:
            out += out[start:start + ml]
        else:
            for k in range(ml):  # overlapping copy
                out.append(out[start + k])
    if len(out) != raw_size:
        raise ValueError(f"lz4: block decodes to {len(out)} bytes, header says {raw_size}")
    return bytes(out)


def decode_frames(ptr, nbytes):
    """A stream of ClickHouse compressed frames (lz4.cu) -> the bytes it carries.  Every frame is decoded
    twice, by pyarrow's LZ4 codec and by the decoder above; the checksum slot must still be zero (the
    host fills it)."""
    if not ptr or not nbytes:
        return b""
    import pyarrow as pa
    raw = bytes((C.c_uint8 * nbytes).from_address(ptr))
    out, i = [], 0
    while i < nbytes:
        if nbytes - i < 25:
            raise ValueError("frame stream: truncated header")
        if raw[i:i + 16] != bytes(16):
            raise ValueError("frame stream: checksum slot not zero")
        method = raw[i + 16]
        csz = int.from_bytes(raw[i + 17:i + 21], "little")
        usz = int.from_bytes(raw[i + 21:i + 25], "little")
        payload = raw[i + 25:i + 16 + csz]
        if len(payload) != csz - 9 or usz == 0 or usz > 4096:
            raise ValueError(f"frame stream: bad sizes {csz} / {usz}")
        if method == 0x82:
            a = lz4_block_decode(payload, usz)
            b = pa.decompress(payload, decompressed_size=usz, codec="lz4_raw", asbytes=True)
            if a != b:
                raise ValueError("frame stream: the two LZ4 decoders disagree")
            if len(payload) >= usz:
                raise ValueError("frame stream: LZ4 frame not smaller than its data (should have been stored)")
            out.append(a)
        elif method == 0x02:
            if len(payload) != usz:
                raise ValueError("frame stream: stored frame size mismatch")
            out.append(payload)
        else:
            raise ValueError(f"frame stream: unknown method {method:#x}")
        i += 16 + csz
    return b"".join(out)


class _Plain:
    """dfgpu_col fields with the frame streams of a compressed column already decoded."""
    pass


def _decompressed(c: Col):
    p = _Plain()
    for k, _ in Col._fields_:
        setattr(p, k, getattr(c, k))
    p._keep = []
    for field, nbytes in (("data", c.data_bytes), ("pos", c.pos_bytes), ("arr_offsets", c.arr_offsets_bytes),
                          ("nulls", c.nulls_bytes)):
        ptr = getattr(c, field)
        if not ptr:
            continue
        raw = decode_frames(ptr, int(nbytes))
        buf = (C.c_uint8 * max(1, len(raw))).from_buffer_copy(raw or b"\0")
        p._keep.append(buf)
        setattr(p, field, C.addressof(buf))
        if field == "data":
            if len(raw) != c.data_raw_bytes:
                raise ValueError(f"{c.name.decode()}: {len(raw)} bytes decoded, data_raw_bytes {c.data_raw_bytes}")
            p.data_bytes = len(raw)
    p.compressed_bytes = int(c.data_bytes + c.pos_bytes + c.arr_offsets_bytes + c.nulls_bytes)
    return p


class Column:
    """Python-side copy of one dfgpu_col (numpy arrays; safe after batch release)."""

    def __init__(self, c: Col):
        self.compressed_bytes = None
        if c.compressed:
            c = _decompressed(c)
            self.compressed_bytes = c.compressed_bytes
        self.name = c.name.decode()
        self.ch_type = c.ch_type
        self.elem_size = c.elem_size
        self.nrows = int(c.nrows)
        t = c.ch_type
        self.values = self.buf = self.pos = self.arr_offsets = self.nulls = None
        self.dict = None
        # dfgpu_cfg.native_strings: buf is the ClickHouse native encoding (uvarint length + bytes per value)
        self.native = t in (CH_STR, CH_LC_STR_ROWS, CH_ARR_STR, CH_ARR_LC_STR) and not c.pos and self.nrows > 0
        if t in (CH_STR, CH_LC_STR_ROWS):
            self.buf = _np_from(c.data, c.data_bytes, np.uint8)
            self.pos = _np_from(c.pos, self.nrows * 16, np.int64).reshape(-1, 2)
        elif t == CH_LC_STR:
            kd = {1: np.uint8, 2: np.uint16, 4: np.uint32}[c.elem_size]
            self.values = _np_from(c.data, self.nrows * c.elem_size, kd).astype(np.int64)
            dpos = _np_from(c.dict_pos, int(c.dict_n) * 16, np.int64).reshape(-1, 2)
            dlen = int(dpos[:, 1].max()) if len(dpos) else 0
            dbuf = _np_from(c.dict_data, dlen, np.uint8).tobytes()
            self.dict = [dbuf[s:e] for s, e in dpos]
        elif t == CH_IPV6:
            self.values = _np_from(c.data, self.nrows * 16, np.uint8).reshape(-1, 16)
        elif t in (CH_ARR_STR, CH_ARR_LC_STR):
            self.arr_offsets = _np_from(c.arr_offsets, self.nrows * 8, np.uint64)
            self.buf = _np_from(c.data, c.data_bytes, np.uint8)
            self.pos = _np_from(c.pos, int(c.nelems) * 16, np.int64).reshape(-1, 2)
        elif t in (CH_ARR_UINT16, CH_ARR_FLOAT64):
            dt = _FIXED_DTYPE[t]
            self.arr_offsets = _np_from(c.arr_offsets, self.nrows * 8, np.uint64)
            self.values = _np_from(c.data, int(c.nelems) * np.dtype(dt).itemsize, dt)
        else:
            dt = _FIXED_DTYPE[t]
            self.values = _np_from(c.data, self.nrows * np.dtype(dt).itemsize, dt)
            if t in (CH_NULLABLE_UINT8, CH_NULLABLE_INT32, CH_NULLABLE_INT64, CH_NULLABLE_UINT64):
                self.nulls = _np_from(c.nulls, self.nrows, np.uint8)

    # -- value-level view used by the parity tests ---------------------------------
    def strings(self):
        b = self.buf.tobytes()
        if self.native:  # decode the native stream
            out, i = [], 0
            while i < len(b):
                n = shift = 0
                while True:
                    c = b[i]
                    i += 1
                    n |= (c & 0x7F) << shift
                    shift += 7
                    if c < 0x80:
                        break
                out.append(b[i:i + n])
                i += n
            return out
        return [b[s:e] for s, e in self.pos]

    def native_bytes(self):
        """The ClickHouse native encoding of a String / Array(String) column's values
        (ch-go ColStr.EncodeColumn: per value uvarint(len) + bytes)."""
        if self.native:
            return self.buf.tobytes()
        out = bytearray()
        b = self.buf.tobytes()
        for s, e in self.pos:
            n = int(e - s)
            while n >= 0x80:
                out.append((n & 0x7F) | 0x80)
                n >>= 7
            out.append(n)
            out += b[s:e]
        return bytes(out)

    def rows(self):
        """Row-wise python values (slow; tests only)."""
        t = self.ch_type
        if t in (CH_STR, CH_LC_STR_ROWS):
            return self.strings()
        if t == CH_LC_STR:
            return [self.dict[k] for k in self.values]
        if t in (CH_ARR_STR, CH_ARR_LC_STR):
            s = self.strings()
            out, prev = [], 0
            for e in self.arr_offsets:
                out.append(s[prev:int(e)])
                prev = int(e)
            return out
        if t in (CH_ARR_UINT16, CH_ARR_FLOAT64):
            out, prev = [], 0
            for e in self.arr_offsets:
                out.append(self.values[prev:int(e)].tolist())
                prev = int(e)
            return out
        if t == CH_IPV6:
            return [bytes(r) for r in self.values]
        if self.nulls is not None:
            return [None if n else v for v, n in zip(self.values.tolist(), self.nulls.tolist())]
        return self.values.tolist()


def columns_equal(a: Column, b: Column):
    """Bit-exact container comparison; returns None or a message."""
    if a.name != b.name:
        return f"name {a.name} != {b.name}"
    if a.ch_type != b.ch_type:
        return f"{a.name}: ch_type {a.ch_type} != {b.ch_type}"
    if a.nrows != b.nrows:
        return f"{a.name}: nrows {a.nrows} != {b.nrows}"
    t = a.ch_type
    if t == CH_LC_STR:
        # dictionary must be in first-appearance order on both sides and decode equal
        if a.dict != b.dict:
            return f"{a.name}: LowCardinality dictionary differs: {a.dict[:8]} vs {b.dict[:8]}"
        if not np.array_equal(a.values, b.values):
            i = int(np.nonzero(a.values != b.values)[0][0])
            return f"{a.name}: key differs at row {i}: {a.values[i]} vs {b.values[i]}"
        return None
    for attr in ("values", "buf", "pos", "arr_offsets", "nulls"):
        x, y = getattr(a, attr), getattr(b, attr)
        if (x is None) != (y is None):
            return f"{a.name}: {attr} presence differs"
        if x is None:
            continue
        if x.shape != y.shape:
            return f"{a.name}: {attr} shape {x.shape} != {y.shape}"
        if x.dtype.kind == "f":
            same = np.array_equal(x.view(np.uint64), y.view(np.uint64))
        else:
            same = np.array_equal(x, y)
        if not same:
            idx = np.argwhere(x != y)
            i = idx[0] if len(idx) else None
            return f"{a.name}: {attr} differs first at {i}: {x[tuple(i)] if i is not None else '?'} vs {y[tuple(i)] if i is not None else '?'}"
    return None


def frame_bytes(fn, handle, table):
    """(bytes of compressed frames over all columns, rows) of a dfgpu_cfg.block_lz4 table, without decoding."""
    ncols = C.c_int(0)
    nrows = C.c_uint64(0)
    rc = fn(handle, table, None, C.byref(ncols), C.byref(nrows))
    if rc != 0:
        raise RuntimeError(f"columns(table={table}) failed: {rc}")
    arr = (Col * ncols.value)()
    rc = fn(handle, table, arr, C.byref(ncols), C.byref(nrows))
    if rc != 0:
        raise RuntimeError(f"columns(table={table}) failed: {rc}")
    tot = 0
    for i in range(ncols.value):
        c = arr[i]
        if c.compressed:
            tot += int(c.data_bytes + c.pos_bytes + c.arr_offsets_bytes + c.nulls_bytes)
    return tot, int(nrows.value)


def fetch_column(fn, handle, table, index):
    """One column of a table (a copy), without touching the others."""
    ncols = C.c_int(0)
    nrows = C.c_uint64(0)
    rc = fn(handle, table, None, C.byref(ncols), C.byref(nrows))
    if rc != 0:
        raise RuntimeError(f"columns(table={table}) failed: {rc}")
    arr = (Col * ncols.value)()
    rc = fn(handle, table, arr, C.byref(ncols), C.byref(nrows))
    if rc != 0:
        raise RuntimeError(f"columns(table={table}) failed: {rc}")
    return Column(arr[index]), int(nrows.value)


def fetch_columns(fn, handle, table):
    """fn(handle, table, cols*, ncols*, nrows*) -> list[Column] (copies)."""
    ncols = C.c_int(0)
    nrows = C.c_uint64(0)
    rc = fn(handle, table, None, C.byref(ncols), C.byref(nrows))
    if rc != 0:
        raise RuntimeError(f"columns(table={table}) failed: {rc}")
    arr = (Col * ncols.value)()
    rc = fn(handle, table, arr, C.byref(ncols), C.byref(nrows))
    if rc != 0:
        raise RuntimeError(f"columns(table={table}) failed: {rc}")
    return [Column(arr[i]) for i in range(ncols.value)], int(nrows.value)
