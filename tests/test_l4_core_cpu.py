"""The L4 decode core of the product (pb_tok.cuh tokenizer + l4_tok.h table + l4_row.cuh row builder — the source
the CUDA kernel l4.cu instantiates), compiled for the HOST by tests/harness and compared with the oracle: every
fixed-width column the row builder produces, request_domain, the acl_gids / aggregated_flow_ids values and the
per-record decode verdict, on clean and edge-case wire input.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from deepflow_b200 import abi
from deepflow_b200 import synth as S
from test_l7_core_cpu import host_framing

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NOT_IN_ROW = {"province_0", "province_1", "observation_point", "team_id", "aggregated_flow_ids", "acl_gids", "request_domain"}


@pytest.fixture(scope="module")
def harness():
    d = os.path.join(HERE, "harness")
    so = os.path.join(d, "libl4harness.so")
    src = os.path.join(d, "l4_harness.cc")
    deps = [src] + [os.path.join(ROOT, "deepflow_b200", "csrc", f) for f in ("pb_tok.cuh", "l4_tok.h", "l4_row.cuh", "schema_l4.h")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, src, "-I", os.path.join(ROOT, "include")])
    L = C.CDLL(so)
    L.harness_l4_ncols.restype = C.c_int
    return L


def run_harness(L, wl, ocols):
    offs, lens, frame_of = host_framing(wl)
    n = len(offs)
    data = np.zeros(wl.nbytes + 64, np.uint8)
    data[:wl.nbytes] = wl.data
    ncols = L.harness_l4_ncols()
    fix_bufs = [None] * ncols
    fix_ptrs = (C.c_void_p * ncols)()
    for i, c in enumerate(ocols):
        if c.values is not None and c.ch_type != abi.CH_LC_STR and c.arr_offsets is None:
            es = 16 if c.ch_type == abi.CH_IPV6 else c.values.dtype.itemsize
            fix_bufs[i] = np.zeros(max(1, n) * es, np.uint8)
            fix_ptrs[i] = fix_bufs[i].ctypes.data
    ok = np.zeros(max(1, n), np.uint8)
    info = np.zeros((8, max(1, n)), np.uint32)
    elems = np.zeros((max(1, n), 16), np.uint64)
    kg = np.zeros((max(1, n), 16), np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = L.harness_l4(p(data), C.c_uint64(wl.nbytes), p(offs), p(lens), C.c_uint32(n), C.c_uint32(0), C.c_uint32(0), p(ok),
                      fix_ptrs, p(info), p(elems), p(kg))
    assert rc == 0
    return dict(n=n, ok=ok, fix=fix_bufs, info=info, elems=elems, data=data, frame_of=frame_of)


@pytest.mark.parametrize("n,seed,flags", [(3000, 11, 0), (3000, 12, S.F_ALL_WIRE), (1500, 13, S.F_UNKNOWN_FIELDS | S.F_WIDE_VARINT),
                                           (1500, 14, S.F_DUP_FIELDS | S.F_OMIT_SUBMSG | S.F_UNPACKED)])
def test_l4_core_matches_oracle(synth, built, harness, n, seed, flags):
    from oracle_binding import Oracle
    ora = Oracle()
    ora.load_synth(synth)
    wl = synth.gen(abi.DFGPU_MSG_TAGGEDFLOW, n, seed=seed, flags=flags)
    ora.submit(wl)
    ocols, orows = ora.columns(abi.TABLE_L4_FLOW_LOG)
    h = run_harness(harness, wl, ocols)
    assert h["n"] == wl.nrecords == orows and h["ok"][:h["n"]].all()
    names = [c.name for c in ocols]
    kg_first = names.index("region_id_0")
    checked = 0
    for i, c in enumerate(ocols):
        if c.name in NOT_IN_ROW or kg_first <= i < kg_first + 39:  # SmartEncoding columns: the KG kernel
            continue
        if h["fix"][i] is None:
            continue
        got = h["fix"][i]
        exp = c.values.reshape(-1).view(np.uint8) if c.ch_type == abi.CH_IPV6 else c.values.view(np.uint8).reshape(-1)
        if c.name == "_id":
            g = got.view(np.uint64)[:orows] & ~np.uint64(0x3ff << 22)
            e = c.values.astype(np.uint64) & ~np.uint64(0x3ff << 22)
            assert np.array_equal(g, e), c.name
        else:
            assert np.array_equal(got[:len(exp)], exp), f"{c.name}: first diff at row {np.argmax(got[:len(exp)] != exp) // max(1, c.values.dtype.itemsize)}"
        checked += 1
    assert checked >= 100, checked
    by = {c.name: c for c in ocols}
    data = h["data"].tobytes()
    doms = by["request_domain"].strings()
    for r in range(orows):
        off, ln = int(h["info"][3, r]), int(h["info"][4, r])
        assert data[off:off + ln] == doms[r], r
    # acl_gids values and the decimal join of aggregated_flow_ids, from the captured values
    acl = by["acl_gids"]
    ends = acl.arr_offsets.astype(np.int64)
    cnt = np.diff(np.concatenate([[0], ends]))
    assert np.array_equal(h["info"][5, :orows].astype(np.int64), cnt)
    aggs = by["aggregated_flow_ids"].strings()
    for r in range(orows):
        na = int(h["info"][6, r])
        assert int(h["info"][7, r]) == len(aggs[r]), r
        if na <= 8:
            assert ",".join(str(int(v)) for v in h["elems"][r, 8:8 + na]).encode() == aggs[r], r
        k = int(cnt[r])
        if k <= 8:
            s = int(ends[r]) - k
            assert [int(v) & 0xffff for v in h["elems"][r, 0:k]] == acl.values[s:s + k].tolist(), r
    ora.close()


def test_l4_core_decode_verdicts(synth, built, harness):
    from oracle_binding import Oracle
    ora = Oracle()
    ora.load_synth(synth)
    wl = synth.gen(abi.DFGPU_MSG_TAGGEDFLOW, 20000, seed=21, flags=S.F_ALL_WIRE | S.F_CORRUPT, max_frame_bytes=4096)
    ora.submit(wl)
    ocols, orows = ora.columns(abi.TABLE_L4_FLOW_LOG)
    h = run_harness(harness, wl, ocols)
    ok, frame_of = h["ok"][:h["n"]].astype(bool), h["frame_of"]
    kept = bad_frames = 0
    for f in range(len(wl.frames)):
        m = ok[frame_of == f]
        bad = np.nonzero(~m)[0]
        kept += int(bad[0]) if len(bad) else len(m)
        bad_frames += 1 if len(bad) else 0
    assert bad_frames > 0 and kept == orows
    ora.close()
