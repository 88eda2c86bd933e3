#!/usr/bin/env python
"""flow_metrics path timing: Document batches through stage+run (device time,
CUDA events of the library) and through dfgpu_submit (host buffers), then the reduce-by-key of
the network_map.1s rows of a replayed batch (every group gets `--replay` rows)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from deepflow_b200 import abi  # noqa: E402
from deepflow_b200 import synth as S  # noqa: E402
from deepflow_b200.native import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=1_000_000)
    ap.add_argument("--pods", type=int, default=100_000)
    ap.add_argument("--replay", type=int, default=8)
    ap.add_argument("--runs", type=int, default=5)
    a = ap.parse_args()
    synth = S.Synth(n_pods=a.pods)
    wl = synth.gen(abi.DFGPU_MSG_METRICS, a.records, seed=500)
    data = np.concatenate([wl.data] * a.replay)
    frames = np.concatenate([wl.frames] * a.replay).copy()
    for r in range(a.replay):
        frames["offset"][r * len(wl.frames):(r + 1) * len(wl.frames)] += r * wl.data.nbytes
    big = S.Workload(data, frames, wl.nrecords * a.replay, abi.DFGPU_MSG_METRICS)
    cfg = abi.Cfg()
    cfg.max_batch_bytes = big.nbytes + (1 << 20)
    eng = Engine(0, cfg)
    eng.load_synth(synth)
    dev = torch.from_numpy(big.data).cuda()
    pinned = torch.empty(big.nbytes, dtype=torch.uint8, pin_memory=True)
    pinned.copy_(torch.from_numpy(big.data))
    for _ in range(2):
        eng.submit(big).release()
    # device resident
    t = []
    for _ in range(a.runs):
        eng.stage_ptr(big.msg_type, dev.data_ptr(), big.nbytes, big.frames.ctypes.data, len(big.frames))
        eng.run()
        b = eng.fetch()
        t.append(b.counters().device_us / 1e3)
        rows = {i: b.counters().rows[abi.TABLE_METRICS_BASE + i] for i in range(9)}
        drop = b.counters().drop_count
        b.release()
    dev_ms = float(np.median(t))
    # host buffers
    t0 = time.perf_counter()
    for _ in range(a.runs):
        big2 = S.Workload(pinned.numpy(), big.frames, big.nrecords, big.msg_type)
        eng.submit(big2).release()
    e2e_ms = (time.perf_counter() - t0) / a.runs * 1e3
    print(f"documents {big.nrecords}  wire {big.nbytes / big.nrecords:.1f} B/doc  rows {rows}  dropped {drop}")
    print(f"  device   {dev_ms:8.3f} ms  -> {big.nrecords / dev_ms / 1e3:.1f} M doc/s")
    print(f"  e2e      {e2e_ms:8.3f} ms  -> {big.nrecords / e2e_ms / 1e3:.1f} M doc/s")
    # reduce network_map.1s (table 6) to 1 m buckets
    eng.submit(big).release()
    tbl = abi.TABLE_METRICS_BASE + 6
    rt = []
    for _ in range(a.runs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rb = eng.metrics_reduce(tbl, 60)
        rt.append((time.perf_counter() - t0) * 1e3)
        _, k = rb.columns(abi.TABLE_METRICS_REDUCED)[0], rb.counters().rows[abi.TABLE_METRICS_REDUCED]
        rb.release()
    n = rows[6]
    ms = float(np.median(rt))
    # algorithmic bytes: every input row's key (network_map tag row 168 B + strings) and 65 raw meter
    # fields (421 B) once, every group once more on output
    alg = n * (168 + 421) + k * (168 + 509)
    print(f"  reduce   {n} rows -> {k} groups  {ms:8.3f} ms (wall, incl. scratch alloc + D2H of the result)"
          f"  {n / ms / 1e3:.1f} M rows/s  {alg / ms / 1e6:.1f} GB/s algorithmic")
    eng.close()


if __name__ == "__main__":
    main()
