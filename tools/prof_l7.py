#!/usr/bin/env python
"""One L7 sub-batch (625 k records, the bench's unit of work) through stage+run, for profiling.

  python tools/prof_l7.py                    stage breakdown from CUDA events (ms per run)
  ncu --set full --import-source on --clock-control none -k regex:k_l7_pass --launch-skip 6 -c 2 \
      -o gpurun_out/prof python tools/prof_l7.py --runs 1          one capture of SIZE + EMIT
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from deepflow_b200 import abi  # noqa: E402
from deepflow_b200.native import Engine  # noqa: E402
from deepflow_b200.synth import Synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=625_000)
    ap.add_argument("--pods", type=int, default=100_000)
    ap.add_argument("--runs", type=int, default=8)
    ap.add_argument("--msg", type=int, default=abi.DFGPU_MSG_PROTOCOLLOG)
    ap.add_argument("--lz4", action="store_true", help="native strings + compressed frames; every run is fetched")
    a = ap.parse_args()
    synth = Synth(n_pods=a.pods)
    wl = synth.gen(a.msg, a.records, seed=1000)
    cfg = abi.Cfg()
    cfg.max_batch_bytes = wl.nbytes + (1 << 20)
    cfg.flow_tag_cache_max_size = int(os.environ.get("FT_CACHE", 1 << 19))
    if a.lz4:
        cfg.native_strings = 1
        cfg.block_lz4 = 1
    eng = Engine(0, cfg)
    eng.load_synth(synth)
    dev = torch.from_numpy(wl.data).cuda()
    for _ in range(3):  # warm-up: the flow_tag set reaches its steady state
        eng.stage_ptr(wl.msg_type, dev.data_ptr(), wl.nbytes, wl.frames.ctypes.data, len(wl.frames))
        eng.run()
    b = eng.fetch()  # any buffer growth / repair happens here, before the timed runs
    c = b.counters()
    print("warm-up: count", c.count, "errors", c.error_count)
    b.release()
    eng.profile_enable(True)
    import time
    for _ in range(a.runs):
        eng.stage_ptr(wl.msg_type, dev.data_ptr(), wl.nbytes, wl.frames.ctypes.data, len(wl.frames))
        eng.run()
        if a.lz4:
            eng.sync()
            t0 = time.perf_counter()
            bb = eng.fetch()
            t1 = time.perf_counter()
            cols, rows = bb.columns(abi.TABLE_L7_FLOW_LOG if a.msg == 5 else abi.TABLE_L4_FLOW_LOG) if _ == 0 else ([], 0)
            if cols:
                print("frame bytes per row", sum(c.compressed_bytes or 0 for c in cols) / max(1, rows))
            print(f"fetch (compress + D2H of frames) {1000 * (t1 - t0):.3f} ms")
            bb.release()
    eng.sync()
    st, n = eng.profile_stages()
    if n:
        tot = sum(st.values())
        print("records", wl.nrecords, "bytes", wl.nbytes, "runs", n)
        for k, v in st.items():
            print(f"  {k:10s} {v / n:8.3f} ms  {100 * v / tot:5.1f} %")
        print(f"  total      {tot / n:8.3f} ms  -> {wl.nrecords / (tot / n) / 1e3:.1f} M rec/s")
        ems, en, _ = eng.profile_read()  # the pack stage (strings + arrays kernels; L4: its second pass)
        if en:
            print(f"  pack kernels {ems / en:8.3f} ms")
    eng.close()


if __name__ == "__main__":
    main()
