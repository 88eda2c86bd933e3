"""TEST-ONLY ctypes binding of oracle/liboracle.so (the CPU restatement of the
reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu baseline may
import this."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from deepflow_b200 import abi  # noqa: E402
from deepflow_b200.build import build_oracle  # noqa: E402

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.oracle_new.restype = C.c_void_p
        L.oracle_new.argtypes = [C.POINTER(abi.Cfg)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_dict_load.argtypes = [C.c_void_p, C.c_uint16, C.POINTER(abi.DictSnapshot)]
        L.oracle_geo_load.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
        L.oracle_submit.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.oracle_columns.argtypes = [C.c_void_p, C.c_int, C.POINTER(abi.Col), C.POINTER(C.c_int),
                                     C.POINTER(C.c_uint64)]
        L.oracle_counters.argtypes = [C.c_void_p, C.POINTER(abi.Counters)]
        L.oracle_metrics_reduce.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.oracle_rollup_add.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        L.oracle_rollup_flush.argtypes = [C.c_void_p, C.c_int]
        _LIB = L
    return _LIB


class Oracle:
    def __init__(self, cfg=None):
        self._h = lib().oracle_new(C.byref(cfg) if cfg is not None else None)
        self.id_counter = C.c_uint32(0)
        self.trace_idx = C.c_uint64(0)

    def close(self):
        if self._h:
            lib().oracle_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dict_load(self, org, snap_ptr):
        rc = lib().oracle_dict_load(self._h, org, snap_ptr)
        assert rc == 0, rc

    def geo_load(self, geo):
        e, n, names, nn = geo
        rc = lib().oracle_geo_load(self._h, e, n, names, nn)
        assert rc == 0, rc

    def load_synth(self, synth, org=1):
        self.dict_load(org, synth.dict_snapshot())
        self.geo_load(synth.geo())

    def submit(self, wl):
        rc = lib().oracle_submit(self._h, wl.msg_type, wl.data.ctypes.data, wl.data.nbytes,
                                 wl.frames.ctypes.data, len(wl.frames), C.byref(self.id_counter),
                                 C.byref(self.trace_idx))
        if rc != 0:
            raise RuntimeError(f"oracle_submit failed: {rc}")

    def columns(self, table):
        return abi.fetch_columns(lib().oracle_columns, self._h, table)

    def counters(self):
        c = abi.Counters()
        lib().oracle_counters(self._h, C.byref(c))
        return c

    def rollup_add(self, table, bucket_seconds=60):
        return lib().oracle_rollup_add(self._h, table, bucket_seconds)

    def rollup_flush(self, table):
        return lib().oracle_rollup_flush(self._h, table)

    def metrics_reduce(self, table, bucket_seconds=60):
        return lib().oracle_metrics_reduce(self._h, table, bucket_seconds)
