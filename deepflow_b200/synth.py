"""ctypes binding of the synthetic workload generator (deepflow_b200/synth/synth.cc,
SURVEY §8d).  Input tooling for tests and bench; not on the hot path."""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

F_UNKNOWN_FIELDS, F_UNPACKED, F_DUP_FIELDS, F_OMIT_SUBMSG, F_CORRUPT, F_WIDE_VARINT, F_HIGH_CARD = 1, 2, 4, 8, 16, 32, 64
F_ALL_WIRE = F_UNKNOWN_FIELDS | F_UNPACKED | F_DUP_FIELDS | F_OMIT_SUBMSG | F_WIDE_VARINT


class _Frames(C.Structure):
    _fields_ = [("data", C.c_void_p), ("nbytes", C.c_uint64), ("frames", C.c_void_p),
                ("nframes", C.c_uint32), ("pad", C.c_uint32), ("nrecords", C.c_uint64)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "synth", "libdfsynth.so")
        if not os.path.exists(path):
            from .build import build_synth
            build_synth()
        L = C.CDLL(path)
        L.synth_new.restype = C.c_void_p
        L.synth_new.argtypes = [C.c_uint64, C.c_uint32]
        L.synth_free.argtypes = [C.c_void_p]
        L.synth_dict.restype = C.POINTER(abi.DictSnapshot)
        L.synth_dict.argtypes = [C.c_void_p]
        L.synth_geo.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64),
                                C.POINTER(C.c_void_p), C.POINTER(C.c_uint32)]
        L.synth_gen.restype = C.c_void_p
        L.synth_gen.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint16,
                                C.c_uint32, C.c_uint32]
        L.synth_batch_view.argtypes = [C.c_void_p, C.POINTER(_Frames)]
        L.synth_batch_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


class Workload:
    """A generated batch: `data` (uint8 numpy, owned copy) + `frames` (FRAME_DTYPE)."""

    def __init__(self, data, frames, nrecords, msg_type):
        self.data, self.frames, self.nrecords, self.msg_type = data, frames, nrecords, msg_type

    @property
    def nbytes(self):
        return int(self.data.nbytes)

    def select(self, mask):
        """Sub-workload made of the frames where mask is True (payload not copied)."""
        fr = self.frames[mask]
        return Workload(self.data, fr, None, self.msg_type)

    def compact(self):
        """Re-pack the selected frames contiguously (what the Go shim's staging does)."""
        total = int(self.frames["len"].sum())
        out = np.empty(total, np.uint8)
        fr = self.frames.copy()
        o = 0
        for i, f in enumerate(self.frames):
            n = int(f["len"])
            out[o:o + n] = self.data[int(f["offset"]):int(f["offset"]) + n]
            fr[i]["offset"] = o
            o += n
        return Workload(out, fr, self.nrecords, self.msg_type)


class Synth:
    def __init__(self, seed=20260921, n_pods=2000):
        self._h = lib().synth_new(seed, n_pods)
        self.seed, self.n_pods = seed, n_pods

    def close(self):
        if self._h:
            lib().synth_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dict_snapshot(self):
        """POINTER(DictSnapshot) valid while this Synth lives."""
        return lib().synth_dict(self._h)

    def geo(self):
        e, n, names, nn = C.c_void_p(), C.c_uint64(), C.c_void_p(), C.c_uint32()
        lib().synth_geo(self._h, C.byref(e), C.byref(n), C.byref(names), C.byref(nn))
        return e, n.value, names, nn.value

    def gen(self, msg_type, nrecords, seed=1, flags=0, max_frame_bytes=0, org_id=1, vtap_mod=0, vtap_rem=0):
        b = lib().synth_gen(self._h, msg_type, seed, nrecords, flags, max_frame_bytes, org_id, vtap_mod, vtap_rem)
        if not b:
            raise ValueError("unsupported msg_type for the generator")
        v = _Frames()
        lib().synth_batch_view(b, C.byref(v))
        data = np.frombuffer((C.c_uint8 * v.nbytes).from_address(v.data), np.uint8).copy() if v.nbytes else np.zeros(0, np.uint8)
        frames = np.frombuffer((C.c_uint8 * (v.nframes * 24)).from_address(v.frames), abi.FRAME_DTYPE).copy()
        nrec = int(v.nrecords)
        lib().synth_batch_free(b)
        return Workload(data, frames, nrec, msg_type)
