"""GPU parity for the framing kernel's staged chase (framing.cu k_frame_walk: the frame streams through a
shared-memory ring of 6 x 4 KB filled by bulk async copies) on the shapes its fast path has to get right:
records longer than the whole ring, frames at unaligned offsets with gaps between them, frames below and
above the small-frame cut-off, and a chain error after a long record
(SimpleDecoder.ReadPB / IsEnd, server/libs/codec/simple_codec.go:229-248,404; abandon-the-buffer rule
decoder.go:206-209)."""
import numpy as np
import pytest

from deepflow_b200 import abi
from deepflow_b200.synth import Workload

pytestmark = pytest.mark.gpu


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _records(wl):
    """[(frame index, payload bytes)] of a well-formed workload."""
    out = []
    data = wl.data
    for fi, f in enumerate(wl.frames):
        o, end = int(f["offset"]), int(f["offset"]) + int(f["len"])
        while o < end:
            n = int.from_bytes(bytes(data[o:o + 4]), "little")
            out.append((fi, bytes(data[o + 4:o + 4 + n])))
            o += 4 + n
    return out


def _reframe(wl, shapes, rng, junk_every=0, junk_bytes=(0, 0), bad_tail_frame=None):
    """Re-pack the records of `wl` into frames of shapes[i] records each (cycled), with 1..15 bytes of gap
    before every frame; every junk_every-th record gets an unknown length-delimited field (number 1000) of
    junk_bytes[0]..[1] random bytes appended, which the decoder must skip."""
    recs = _records(wl)
    buf = bytearray()
    frames = []
    k = 0
    si = 0
    while k < len(recs):
        take = shapes[si % len(shapes)]
        si += 1
        buf += bytes(rng.integers(0, 256, int(rng.integers(1, 16)), dtype=np.uint8))  # gap: unaligned start
        start = len(buf)
        src = recs[k][0]
        for j in range(k, min(len(recs), k + take)):
            p = recs[j][1]
            if junk_every and j % junk_every == 0:
                n = int(rng.integers(junk_bytes[0], junk_bytes[1] + 1))
                p = p + _varint(1000 << 3 | 2) + _varint(n) + bytes(rng.integers(0, 256, n, dtype=np.uint8))
            buf += len(p).to_bytes(4, "little") + p
        k += take
        if bad_tail_frame is not None and len(frames) == bad_tail_frame:
            buf += (1 << 30).to_bytes(4, "little") + b"xx"  # header pointing past the frame: abandon the rest
        fr = wl.frames[src].copy()
        fr["offset"], fr["len"] = start, len(buf) - start
        frames.append(fr)
    return Workload(np.frombuffer(bytes(buf), np.uint8).copy(), np.array(frames, dtype=wl.frames.dtype), None,
                    wl.msg_type)


@pytest.mark.parametrize("shapes,junk_every,junk_bytes,bad", [
    ([1, 2, 3, 40, 400], 0, (0, 0), None),              # frames from ~0.5 KB (direct loads) to ~170 KB (ring)
    ([50], 7, (30_000, 120_000), None),                 # records of 30-120 KB: longer than the 24 KB ring
    ([200], 5, (3_000, 9_000), None),                   # records spanning 1-3 ring windows
    ([30, 300], 11, (25_000, 60_000), 1),               # a chain error in a staged frame, after long records
])
def test_framing_shapes(synth, built, shapes, junk_every, junk_bytes, bad):
    from test_l7_gpu import fresh, run_both
    rng = np.random.default_rng(7)
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = 1 << 21
    cfg.max_batch_bytes = 64 << 20
    eng, ora = fresh(synth, cfg)
    try:
        base = synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 3000, seed=4711)
        wl = _reframe(base, shapes, rng, junk_every, junk_bytes, bad)
        problems = run_both(eng, ora, wl)
        assert not problems, "\n".join(problems[:20])
        c = ora.counters()
        if bad is None:
            assert c.error_count == 0 and c.count == 3000
        else:
            assert c.error_count == 1 and c.count == 3000  # the bad header follows the frame's last record
    finally:
        eng.close()
        ora.close()
