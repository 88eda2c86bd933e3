"""GPU: the C++ mirror of the decoder goroutine (deepflow_b200/host) — RecvBuffers in through
the droplet queue, column blocks out to a sink — against the CPU oracle fed the same frames
in the same Gets() batches.  Checks the boundary contract of SURVEY §8(b): every RecvBuffer
released exactly once, flush ticks forwarded, counters as decoder.Counter, blocks identical."""
import ctypes as C

import numpy as np
import pytest

from deepflow_b200 import abi
from deepflow_b200 import synth as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H(built):
    from deepflow_b200 import build
    L = C.CDLL(build.build_host())
    vp = C.c_void_p
    L.dfhost_decoder_new.restype = vp
    L.dfhost_decoder_new.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(abi.Cfg), C.c_int]
    L.dfhost_decoder_free.argtypes = [vp]
    L.dfhost_decoder_dict_load.argtypes = [vp, C.c_uint16, C.POINTER(abi.DictSnapshot)]
    L.dfhost_decoder_geo_load.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32]
    L.dfhost_decoder_feed.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_int]
    L.dfhost_decoder_drain.argtypes = [vp]
    L.dfhost_decoder_stats.argtypes = [vp, C.POINTER(C.c_int64)]
    L.dfhost_decoder_batch.restype = vp
    L.dfhost_decoder_batch.argtypes = [vp, C.c_int]
    return L


class BorrowedBatch:
    """A dfgpu_batch owned by the C++ sink."""

    def __init__(self, h):
        self._h = h

    def columns(self, table):
        from deepflow_b200.native import lib
        return abi.fetch_columns(lib().dfgpu_batch_columns, self._h, table)


@pytest.mark.parametrize("msg,tables", [
    (abi.DFGPU_MSG_TAGGEDFLOW, [abi.TABLE_L4_FLOW_LOG]),
    (abi.DFGPU_MSG_PROTOCOLLOG, [abi.TABLE_L7_FLOW_LOG, abi.TABLE_L7_CUSTOM_FIELD, abi.TABLE_L7_CUSTOM_FIELD_VALUE,
                                 abi.TABLE_L7_APP_SERVICE]),
    (abi.DFGPU_MSG_METRICS, [abi.TABLE_METRICS_BASE + i for i in range(9)] + [abi.TABLE_METRICS_APP_SERVICE]),
])
def test_decoder_loop_matches_oracle(synth, H, msg, tables):
    from oracle_binding import Oracle
    from parity import compare_tables
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = 1 << 21
    wl = synth.gen(msg, 30000, seed=41, flags=S.F_CORRUPT, max_frame_bytes=2048)  # a few thousand frames
    nf = len(wl.frames)
    assert nf > 1024  # more than one Gets() batch
    d = H.dfhost_decoder_new(0, msg, 0, C.byref(cfg), 1 << 14)
    assert d, "dfgpu_init failed (no CPU fallback)"
    ora = Oracle(cfg)
    ora.load_synth(synth)
    try:
        assert H.dfhost_decoder_dict_load(d, 1, synth.dict_snapshot()) == 0
        e, n, names, nn = synth.geo()
        assert H.dfhost_decoder_geo_load(d, e, n, names, nn) == 0
        H.dfhost_decoder_feed(d, wl.data.ctypes.data, wl.frames.ctypes.data, 0, nf, 1)
        nbatches = H.dfhost_decoder_drain(d)
        st = (C.c_int64 * (16 + abi.TABLE_MAX))()
        H.dfhost_decoder_stats(d, st)
        raw, count, err, drop = st[0], st[1], st[2], st[3]
        submits, flushes, acquired, released, nblocks, sink_flushes = st[10], st[11], st[12], st[13], st[14], st[15]
        assert nbatches == (nf + 1 + 1023) // 1024 and submits == nblocks == (nf + 1023) // 1024
        assert raw == nf and acquired == nf and released == nf  # each RecvBuffer released exactly once
        assert flushes == 1 and sink_flushes == 1
        # the oracle sees the same Gets() batches: 1024 frames per submit
        tot_count = tot_err = tot_drop = 0
        for k in range(nblocks):
            lo, hi = k * 1024, min(nf, (k + 1) * 1024)
            sub = S.Workload(wl.data, wl.frames[lo:hi].copy(), None, msg)
            ora.submit(sub)
            oc = ora.counters()
            tot_count += oc.count
            tot_err += oc.error_count
            tot_drop += oc.drop_count
            problems = compare_tables(BorrowedBatch(H.dfhost_decoder_batch(d, k)), ora, tables)
            assert not problems, f"block {k}:\n" + "\n".join(problems[:10])
        assert (count, err, drop) == (tot_count, tot_err, tot_drop)
    finally:
        H.dfhost_decoder_free(d)
        ora.close()


def test_queue_overwrite_releases_buffers(synth, H):
    """A full queue overwrites the oldest RecvBuffers and releases them (queue.go:130-166):
    nothing leaks, and only what was still queued gets decoded."""
    cfg = abi.Cfg()
    wl = synth.gen(abi.DFGPU_MSG_TAGGEDFLOW, 4000, seed=42, max_frame_bytes=2048)
    nf = len(wl.frames)
    d = H.dfhost_decoder_new(0, abi.DFGPU_MSG_TAGGEDFLOW, 0, C.byref(cfg), 64)
    assert d
    try:
        assert H.dfhost_decoder_dict_load(d, 1, synth.dict_snapshot()) == 0
        e, n, names, nn = synth.geo()
        assert H.dfhost_decoder_geo_load(d, e, n, names, nn) == 0
        assert nf > 200
        H.dfhost_decoder_feed(d, wl.data.ctypes.data, wl.frames.ctypes.data, 0, nf, 0)
        H.dfhost_decoder_drain(d)
        st = (C.c_int64 * (16 + abi.TABLE_MAX))()
        H.dfhost_decoder_stats(d, st)
        assert st[0] == 64 and st[12] == nf and st[13] == nf  # 64 decoded, all nf released
    finally:
        H.dfhost_decoder_free(d)
