"""The L7 decode core of the product (pb_tok.cuh tokenizer + l7_row.cuh row builder — the source the CUDA
kernel l7_pass.cu instantiates), compiled for the HOST by tests/harness and compared with the oracle:
every fixed-width column the row builder produces, the nullable flags, all 22 string columns, the array
element counts and the per-record decode verdict, on clean and edge-case wire input.  Test infrastructure
only: libdfgpu.so has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from deepflow_b200 import abi
from deepflow_b200 import synth as S

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_CS = 0x80000000

# columns l7_build_row does not produce (scans / SmartEncoding kernel / flow_tag stage write them)
NOT_IN_ROW = {"trace_id_index", "observation_point", "team_id", "gprocess_id_0", "gprocess_id_1"}


@pytest.fixture(scope="module")
def harness():
    d = os.path.join(HERE, "harness")
    so = os.path.join(d, "libl7harness.so")
    src = os.path.join(d, "l7_harness.cc")
    deps = [src] + [os.path.join(ROOT, "deepflow_b200", "csrc", f) for f in
                    ("pb_tok.cuh", "l7_tok.h", "l7_row.cuh", "l7_tables_host.h", "schema_l7.h", "schema_l4.h")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, src,
                               "-I", os.path.join(ROOT, "include")])
    L = C.CDLL(so)
    L.harness_l7_cs_blob.restype = C.c_void_p
    L.harness_l7_cs_blob.argtypes = [C.POINTER(C.c_uint64)]
    L.harness_l7_ncols.restype = C.c_int
    return L


def host_framing(wl):
    """SimpleDecoder.ReadPB chain (simple_codec.go:229-248) on the host: record offsets / lengths per frame."""
    data = wl.data
    offs, lens, frame_of = [], [], []
    for fi, f in enumerate(wl.frames):
        o, end = int(f["offset"]), int(f["offset"]) + int(f["len"])
        while o < end:
            if end - o < 4:
                break
            n = int.from_bytes(data[o:o + 4].tobytes(), "little")
            o += 4
            if n > end - o:
                break
            offs.append(o)
            lens.append(n)
            frame_of.append(fi)
            o += n
    return np.array(offs, np.uint32), np.array(lens, np.uint32), np.array(frame_of, np.int64)


def run_harness(L, wl, ocols, nat=0, analyzer_id=0, id_counter=0):
    offs, lens, frame_of = host_framing(wl)
    n = len(offs)
    data = np.zeros(wl.nbytes + 64, np.uint8)  # padded: the window prefetches one block ahead
    data[:wl.nbytes] = wl.data
    assert data.ctypes.data % 8 == 0
    ncols = L.harness_l7_ncols()
    fix_bufs, null_bufs = [None] * ncols, [None] * ncols
    fix_ptrs = (C.c_void_p * ncols)()
    null_ptrs = (C.c_void_p * ncols)()
    for i, c in enumerate(ocols):
        if c.values is not None and c.ch_type != abi.CH_LC_STR and c.arr_offsets is None:
            es = 16 if c.ch_type == abi.CH_IPV6 else c.values.dtype.itemsize
            fix_bufs[i] = np.zeros(max(1, n) * es, np.uint8)
            fix_ptrs[i] = fix_bufs[i].ctypes.data
        if c.nulls is not None:
            null_bufs[i] = np.zeros(max(1, n), np.uint8)
            null_ptrs[i] = null_bufs[i].ctypes.data
    ok = np.zeros(max(1, n), np.uint8)
    str_off = np.zeros((22, max(1, n)), np.uint32)
    str_len = np.zeros((22, max(1, n)), np.uint32)
    aux_off = np.zeros((4, max(1, n)), np.uint32)
    aux_len = np.zeros((4, max(1, n)), np.uint32)
    arr = np.zeros((7, max(1, n)), np.uint32)
    info = np.zeros((4, max(1, n)), np.uint32)
    thash = np.zeros(max(1, n), np.uint64)
    kg = np.zeros((max(1, n), 16), np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.harness_l7(p(data), C.c_uint64(wl.nbytes), p(offs), p(lens), C.c_uint32(n), C.c_uint32(nat),
                      C.c_uint32(analyzer_id), C.c_uint32(id_counter), p(ok), fix_ptrs, null_ptrs, p(str_off),
                      p(str_len), p(aux_off), p(aux_len), p(arr), p(info), p(thash), p(kg))
    assert rc == 0
    return dict(n=n, ok=ok, fix=fix_bufs, nulls=null_bufs, str_off=str_off, str_len=str_len, arr=arr, info=info,
                thash=thash, data=data, frame_of=frame_of)


STREAM_COLS = ["biz_code", "biz_scenario", "biz_response_code", "process_kname_0", "process_kname_1", "biz_protocol",
               "version", "request_type", "request_domain", "request_resource", "endpoint", "response_exception",
               "response_result", "http_proxy_client", "x_request_id_0", "x_request_id_1", "trace_id", "_trace_id_2",
               "span_id", "parent_span_id", "app_service", "app_instance"]


@pytest.mark.parametrize("n,seed,flags", [(3000, 11, 0), (3000, 12, S.F_ALL_WIRE), (1500, 13, S.F_UNKNOWN_FIELDS | S.F_WIDE_VARINT),
                                           (1500, 14, S.F_DUP_FIELDS | S.F_OMIT_SUBMSG | S.F_UNPACKED)])
def test_l7_core_matches_oracle(synth, built, harness, n, seed, flags):
    from oracle_binding import Oracle
    ora = Oracle()
    ora.load_synth(synth)
    wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, n, seed=seed, flags=flags)
    ora.submit(wl)
    ocols, orows = ora.columns(abi.TABLE_L7_FLOW_LOG)
    # the oracle's analyzer id comes from the dictionary; _id is checked with it masked out below
    h = run_harness(harness, wl, ocols)
    assert h["n"] == wl.nrecords == orows and h["ok"][:h["n"]].all()
    nb = C.c_uint64()
    blob_p = harness.harness_l7_cs_blob(C.byref(nb))
    blob = bytes((C.c_uint8 * nb.value).from_address(blob_p))
    data = h["data"].tobytes()
    checked = 0
    for i, c in enumerate(ocols):
        if c.name in NOT_IN_ROW or i < 39:  # the 39 SmartEncoding columns come first (l7_kg.cu)
            continue
        if h["fix"][i] is not None:
            got = h["fix"][i]
            if c.ch_type == abi.CH_IPV6:
                exp = c.values.reshape(-1).view(np.uint8)
            else:
                exp = c.values.view(np.uint8).reshape(-1)
            if c.name == "_id":  # genID: time << 32 | analyzer << 22 | counter; analyzer id 0 in the harness
                g = got.view(np.uint64)[:orows] & ~np.uint64(0x3ff << 22)
                e = c.values.astype(np.uint64) & ~np.uint64(0x3ff << 22)
                assert np.array_equal(g, e), c.name
            else:
                assert np.array_equal(got[:len(exp)], exp), f"{c.name}: first diff at {np.argmax(got[:len(exp)] != exp) // max(1, c.values.dtype.itemsize)}"
            checked += 1
        if c.nulls is not None:
            assert np.array_equal(h["nulls"][i][:orows], c.nulls), c.name + " nulls"
    assert checked >= 45, checked
    by_name = {c.name: c for c in ocols}
    for s, name in enumerate(STREAM_COLS):
        exp = by_name[name].strings()
        for r in range(orows):
            off, ln = int(h["str_off"][s, r]), int(h["str_len"][s, r])
            got = blob[(off & ~REF_CS):(off & ~REF_CS) + ln] if off & REF_CS else data[off:off + ln]
            assert got == exp[r], (name, r, got, exp[r])
    # array element counts per row (attribute_names / values, metrics_names / values)
    for k, name in ((0, "attribute_names"), (2, "attribute_values"), (4, "metrics_names"), (6, "metrics_values")):
        ends = by_name[name].arr_offsets.astype(np.int64)
        cnt = np.diff(np.concatenate([[0], ends]))
        assert np.array_equal(h["arr"][k, :orows].astype(np.int64), cnt), name
    for k, name in ((1, "attribute_names"), (3, "attribute_values"), (5, "metrics_names")):
        c = by_name[name]
        ends = c.arr_offsets.astype(np.int64)
        lens = (c.pos[:, 1] - c.pos[:, 0]).astype(np.int64)
        cs = np.concatenate([[0], np.cumsum(lens)])
        per_row = cs[ends] - cs[np.concatenate([[0], ends[:-1]])]
        assert np.array_equal(h["arr"][k, :orows].astype(np.int64), per_row), name + " bytes"
    ora.close()


def test_l7_core_decode_verdicts(synth, built, harness):
    """Corrupt wire input: a record the tokenizer rejects ends its frame (decoder.go:225-229); the rows the
    oracle keeps per frame = the records before the first one the host-compiled tokenizer rejects."""
    from oracle_binding import Oracle
    ora = Oracle()
    ora.load_synth(synth)
    wl = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 20000, seed=21, flags=S.F_ALL_WIRE | S.F_CORRUPT, max_frame_bytes=4096)
    ora.submit(wl)
    ocols, orows = ora.columns(abi.TABLE_L7_FLOW_LOG)
    h = run_harness(harness, wl, ocols)
    ok, frame_of = h["ok"][:h["n"]].astype(bool), h["frame_of"]
    kept = 0
    bad_frames = 0
    for f in range(len(wl.frames)):
        m = ok[frame_of == f]
        bad = np.nonzero(~m)[0]
        kept += int(bad[0]) if len(bad) else len(m)
        bad_frames += 1 if len(bad) else 0
    assert bad_frames > 0
    assert kept == orows
    ora.close()
