"""ctypes binding of libdfgpu.so — the Python mirror of the C ABI in
include/dfgpu.h (the same calls the Go/cgo shim makes).

There is NO CPU fallback: if the CUDA library is missing or no device is
visible, every entry point raises."""
import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class DfgpuError(RuntimeError):
    pass


def lib():
    global _LIB
    if _LIB is None:
        path = os.environ.get("DFGPU_LIB") or os.path.join(_HERE, "libdfgpu.so")  # DFGPU_LIB: ablation builds
        if not os.path.exists(path):
            raise DfgpuError(f"{path} is not built — run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(path)
        vp = C.c_void_p
        L.dfgpu_init.argtypes = [C.c_int, C.POINTER(abi.Cfg), C.POINTER(vp)]
        L.dfgpu_close.argtypes = [vp]
        L.dfgpu_last_error.restype = C.c_char_p
        L.dfgpu_last_error.argtypes = [vp]
        L.dfgpu_dict_load.argtypes = [vp, C.c_uint16, C.POINTER(abi.DictSnapshot)]
        L.dfgpu_geo_load.argtypes = [vp, vp, C.c_uint64, vp, C.c_uint32]
        L.dfgpu_submit.argtypes = [vp, C.c_int, vp, C.c_size_t, vp, C.c_int, C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_uint64), C.POINTER(vp)]
        L.dfgpu_stage.argtypes = [vp, C.c_int, vp, C.c_size_t, vp, C.c_int]
        L.dfgpu_run.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.dfgpu_fetch.argtypes = [vp, C.POINTER(vp)]
        L.dfgpu_sync.argtypes = [vp]
        L.dfgpu_stream.restype = vp
        L.dfgpu_stream.argtypes = [vp]
        L.dfgpu_launch_count.restype = C.c_uint64
        L.dfgpu_launch_count.argtypes = [vp]
        L.dfgpu_batch_columns.argtypes = [vp, C.c_int, C.POINTER(abi.Col), C.POINTER(C.c_int),
                                          C.POINTER(C.c_uint64)]
        L.dfgpu_batch_counters.argtypes = [vp, C.POINTER(abi.Counters)]
        L.dfgpu_batch_release.argtypes = [vp]
        L.dfgpu_batch_release.restype = None
        L.dfgpu_metrics_reduce.argtypes = [vp, C.c_int, C.c_uint32, C.POINTER(vp)]
        L.dfgpu_rollup_add.argtypes = [vp, C.c_int, C.c_uint32]
        L.dfgpu_rollup_flush.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.dfgpu_rollup_add_rows.argtypes = [vp, C.c_int, C.c_uint32, vp]
        L.dfgpu_rollup_describe.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), vp, vp]
        L.dfgpu_batch_tag_fingerprints.argtypes = [vp, C.c_int, C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)]
        L.dfgpu_tags_import.argtypes = [vp, vp, vp, C.c_uint64]
        L.dfgpu_flush.argtypes = [vp]
        L.dfgpu_tags_reset.argtypes = [vp]
        L.dfgpu_profile_enable.argtypes = [vp, C.c_int]
        L.dfgpu_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.dfgpu_profile_stages.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        _LIB = L
    return _LIB


class Batch:
    def __init__(self, handle):
        self._h = handle

    def columns(self, table):
        return abi.fetch_columns(lib().dfgpu_batch_columns, self._h, table)

    def frame_bytes(self, table):
        """(compressed frame bytes, rows) of a block_lz4 table — sizes only, nothing is decoded."""
        return abi.frame_bytes(lib().dfgpu_batch_columns, self._h, table)

    def column(self, table, index):
        """One column (copied), e.g. the `time` column of a flow_tag table."""
        return abi.fetch_column(lib().dfgpu_batch_columns, self._h, table, index)

    def counters(self):
        c = abi.Counters()
        lib().dfgpu_batch_counters(self._h, C.byref(c))
        return c

    def tag_fingerprints(self, table):
        """(n, 2) uint64 copy of the lru-key fingerprints of a flow_tag table's rows."""
        import numpy as np
        p = C.POINTER(C.c_uint64)()
        n = C.c_uint64(0)
        rc = lib().dfgpu_batch_tag_fingerprints(self._h, table, C.byref(p), C.byref(n))
        if rc != 0:
            raise DfgpuError(f"dfgpu_batch_tag_fingerprints(table={table}) failed: {rc}")
        if n.value == 0:
            return np.zeros((0, 2), np.uint64)
        return np.ctypeslib.as_array(p, shape=(int(n.value), 2)).copy()

    def release(self):
        if self._h:
            lib().dfgpu_batch_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Engine:
    """One dfgpu_ctx == one decoder goroutine's worth of state."""

    def __init__(self, device=0, cfg=None):
        self._h = C.c_void_p()
        rc = lib().dfgpu_init(device, C.byref(cfg) if cfg is not None else None, C.byref(self._h))
        if rc != 0:
            raise DfgpuError(f"dfgpu_init(device={device}) failed: {rc} (no CUDA device? there is no CPU fallback)")
        self.id_counter = C.c_uint32(0)
        self.trace_idx = C.c_uint64(0)

    def _check(self, rc, what):
        if rc != 0:
            raise DfgpuError(f"{what} failed: {rc}: {lib().dfgpu_last_error(self._h).decode()}")

    def close(self):
        if self._h:
            lib().dfgpu_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dict_load(self, org, snap_ptr):
        self._check(lib().dfgpu_dict_load(self._h, org, snap_ptr), "dfgpu_dict_load")

    def geo_load(self, geo):
        e, n, names, nn = geo
        self._check(lib().dfgpu_geo_load(self._h, e, n, names, nn), "dfgpu_geo_load")

    def load_synth(self, synth, org=1):
        self.dict_load(org, synth.dict_snapshot())
        self.geo_load(synth.geo())

    def submit(self, wl):
        """End to end: host buffers in, host column blocks out."""
        out = C.c_void_p()
        rc = lib().dfgpu_submit(self._h, wl.msg_type, wl.data.ctypes.data, wl.data.nbytes,
                                wl.frames.ctypes.data, len(wl.frames), C.byref(self.id_counter),
                                C.byref(self.trace_idx), C.byref(out))
        self._check(rc, "dfgpu_submit")
        return Batch(out)

    def stage_ptr(self, msg_type, data_ptr, nbytes, frames_ptr, nframes):
        self._check(lib().dfgpu_stage(self._h, msg_type, data_ptr, nbytes, frames_ptr, nframes), "dfgpu_stage")

    def stage(self, wl):
        self.stage_ptr(wl.msg_type, wl.data.ctypes.data, wl.data.nbytes, wl.frames.ctypes.data, len(wl.frames))

    def run(self, with_counters=False):
        if with_counters:
            rc = lib().dfgpu_run(self._h, C.byref(self.id_counter), C.byref(self.trace_idx))
        else:
            rc = lib().dfgpu_run(self._h, None, None)
        self._check(rc, "dfgpu_run")

    def fetch(self):
        out = C.c_void_p()
        self._check(lib().dfgpu_fetch(self._h, C.byref(out)), "dfgpu_fetch")
        return Batch(out)

    def sync(self):
        self._check(lib().dfgpu_sync(self._h), "dfgpu_sync")

    def tags_import(self, fps, times):
        """Insert keys other ranks emitted: fps (n, 2) uint64, times (n,) uint32 (numpy, host)."""
        import numpy as np
        fps = np.ascontiguousarray(fps, np.uint64)
        times = np.ascontiguousarray(times, np.uint32)
        n = len(times)
        if n:
            self._check(lib().dfgpu_tags_import(self._h, fps.ctypes.data, times.ctypes.data, n), "dfgpu_tags_import")

    def tags_import_ptr(self, fps_ptr, times_ptr, n):
        """Same, from raw pointers — host or device memory ((n, 2) uint64 fingerprints, (n,) uint32 times)."""
        if n:
            self._check(lib().dfgpu_tags_import(self._h, fps_ptr, times_ptr, n), "dfgpu_tags_import")

    def metrics_reduce(self, table, bucket_seconds=60):
        """Reduce-by-key over one table of the last metrics batch (still resident in HBM)."""
        out = C.c_void_p()
        self._check(lib().dfgpu_metrics_reduce(self._h, table, bucket_seconds, C.byref(out)), "dfgpu_metrics_reduce")
        return Batch(out)

    def tags_reset(self):
        self._check(lib().dfgpu_tags_reset(self._h), "dfgpu_tags_reset")

    def rollup_add(self, table, bucket_seconds=60):
        """Fold the rows of `table` of the last metrics submit into the table's rollup accumulator."""
        self._check(lib().dfgpu_rollup_add(self._h, table, bucket_seconds), "dfgpu_rollup_add")

    def rollup_add_rows(self, table, bucket_seconds, tag_col_ptrs, raw_ptr, nrows):
        """Fold decoded rows already resident in HBM (device pointers): BASELINE config 3."""
        cols = (C.c_void_p * len(tag_col_ptrs))(*tag_col_ptrs)
        rows = abi.RollupRows(nrows, C.cast(cols, C.c_void_p), len(tag_col_ptrs), raw_ptr)
        self._check(lib().dfgpu_rollup_add_rows(self._h, table, bucket_seconds, C.byref(rows)), "dfgpu_rollup_add_rows")

    def rollup_describe(self, table):
        """(elem sizes per tag column, kinds per tag column, raw meter fields per row)"""
        ntag, nf = C.c_int(), C.c_int()
        elem, kind = (C.c_uint8 * 128)(), (C.c_uint8 * 128)()
        self._check(lib().dfgpu_rollup_describe(self._h, table, C.byref(ntag), C.byref(nf), elem, kind), "dfgpu_rollup_describe")
        return list(elem[:ntag.value]), list(kind[:ntag.value]), nf.value

    def rollup_flush(self, table):
        out = C.c_void_p()
        self._check(lib().dfgpu_rollup_flush(self._h, table, C.byref(out)), "dfgpu_rollup_flush")
        return Batch(out)

    def profile_enable(self, on=True):
        self._check(lib().dfgpu_profile_enable(self._h, 1 if on else 0), "dfgpu_profile_enable")

    def profile_read(self):
        ms, n, b = C.c_double(), C.c_uint64(), C.c_uint64()
        self._check(lib().dfgpu_profile_read(self._h, C.byref(ms), C.byref(n), C.byref(b)), "dfgpu_profile_read")
        return ms.value, int(n.value), int(b.value)

    def profile_stages(self):
        """ms per stage summed over the profiled runs: framing, prepare, decode, scans, pack (strings), SmartEncoding,
        flow_tag."""
        ms = (C.c_double * 7)()
        n = C.c_uint64()
        self._check(lib().dfgpu_profile_stages(self._h, ms, C.byref(n)), "dfgpu_profile_stages")
        names = ("framing", "prepare", "decode", "scans", "pack", "smartencoding", "flow_tag")
        return {k: ms[i] for i, k in enumerate(names)}, int(n.value)

    @property
    def stream(self):
        return lib().dfgpu_stream(self._h)

    @property
    def launch_count(self):
        return int(lib().dfgpu_launch_count(self._h))
