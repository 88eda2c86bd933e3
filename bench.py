#!/usr/bin/env python3
"""bench.py — flow records/sec of the ingest hot path (decode -> SmartEncoding -> flow_tag / rollup ->
ClickHouse column pack) on B200, per BASELINE.json.

Headline (N=1: BASELINE.json configs[1], "C2"): 10 M synthetic L7 HTTP AppProtoLogsData records per GPU,
submitted the way the ingester's decoder sees them: sub-batches of <= 1024 agent frames (decoder.go:56
BUFFER_SIZE x 256 KiB frames).  One "step" = one pass over all sub-batches.  Under torchrun every rank owns
the vtap_id % N == rank shard (weak scaling); per-record work has no collective — the only exchange is the
all-gather of newly emitted flow_tag keys after every e2e step (deepflow_b200/sharded.py).

  value     device-resident: inputs already in HBM, kernels only; every decoder context of the GPU runs its
            share of the sub-batches on its own stream; CUDA events on a timing stream that waits for all
  e2e       through the C ABI with HOST buffers: pinned H2D of every sub-batch, all kernels, D2H of every
            column block inside the timed region
  roofline  the dominant kernel stage, timed live in a leg where ONE decoder runs alone (CUDA events of the
            library on its stream around every stage): algorithmic bytes / stage time vs the measured HBM
            peak; `path` = all stages over SURVEY 8(d)'s bytes; `traffic` = DRAM bytes of one launch from
            the committed ncu capture (profiles/)
  cpu_baseline   the CPU oracle (a C++ restatement of the Go path) on a bounded sample

Further legs in the same JSON line (`legs`), each with value / e2e / roofline / cpu_baseline:
  c4_mixed     BASELINE configs[3]: 50 % L4 / 35 % L7 / 15 % metric records, one decoder pool per message type
  c3_rollup    BASELINE configs[2]: 100 M decoded network_map.1s rows -> 1 m groups (N=1 only)
  c5_highcard  BASELINE configs[4]: P = 1 M dictionary, high-cardinality attributes, cold flow_tag set every
               step: emitted keys, merge bytes and seconds
  parity       column hashes of one full-size sub-batch per decoder context (3 contexts concurrently, the
               bench's dictionary) against the oracle on the same sub-batch (N=1 only)

--impl reference times the CPU restatement with all host threads instead (the Go reference cannot be built in
this image: no Go toolchain).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from deepflow_b200 import abi  # noqa: E402

L4, L7, MET = abi.DFGPU_MSG_TAGGEDFLOW, abi.DFGPU_MSG_PROTOCOLLOG, abi.DFGPU_MSG_METRICS
METRIC = "flow records/sec ingest+SmartEncode (L7 decode + tag-inject + flow_tag + column pack)"
MAIN_TABLE = {L4: abi.TABLE_L4_FLOW_LOG, L7: abi.TABLE_L7_FLOW_LOG}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dfgpu", choices=["dfgpu", "reference"])
    ap.add_argument("--records", type=int, default=10_000_000, help="records per GPU per step (C2, C4)")
    ap.add_argument("--sub-batch", type=int, default=625_000, help="records per dfgpu_submit")
    ap.add_argument("--pods", type=int, default=100_000, help="dictionary size P (SURVEY 8d)")
    ap.add_argument("--decoders", type=int, default=3, help="contexts per GPU (DecoderQueueCount)")
    ap.add_argument("--legs", default="c2,c4,c3,c5,parity,modes", help="comma list; c2 always runs; modes = e2e in the other output modes")
    ap.add_argument("--e2e-mode", default="lz4", choices=["containers", "native", "lz4"],
                    help="output mode of the headline e2e: ch-go containers (Buf + Positions), native string streams, "
                         "or native strings + LZ4 frames compressed on the GPU (dfgpu_cfg.native_strings / block_lz4)")
    ap.add_argument("--c3-rows", type=int, default=100_000_000)
    ap.add_argument("--c3-groups", type=int, default=1_000_000)
    ap.add_argument("--c5-records", type=int, default=5_000_000)
    ap.add_argument("--c5-pods", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def gen_sub_batches(synth, msg, n_total, sub, seed0, world, rank, flags=0, threads=16):
    """Generate ceil(n_total/sub) sub-batches in parallel threads (ctypes drops the GIL)."""
    sizes = [min(sub, n_total - i) for i in range(0, n_total, sub)]
    out = [None] * len(sizes)
    it = iter(range(len(sizes)))
    lock = threading.Lock()

    def loop():
        while True:
            with lock:
                i = next(it, None)
            if i is None:
                return
            out[i] = synth.gen(msg, sizes[i], seed=seed0 + i, flags=flags, vtap_mod=world if world > 1 else 0, vtap_rem=rank)

    ts = [threading.Thread(target=loop) for _ in range(min(threads, len(sizes)))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"

    def __init__(self, index):
        self.index, self.p = index, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, busy, reasons = [], [], [], set()
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                util = float(f[7])
            except ValueError:
                continue
            if util >= 50:  # "under load": the sampler also spans the host-side generation between the legs
                busy.append(sm[-1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(busy)) if busy else (float(np.median(sm)) if sm else None),
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "samples_under_load": len(busy)}


def bind_near_gpu(index):
    """Run this rank's host threads (and, by first touch, place its pinned staging buffers) on the CPUs local to
    its GPU's PCIe root, like a NUMA-aware ingester deployment would.  Returns what it did."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c < ncpu]
        if cpus and len(cpus) < ncpu:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {index}"
        return "no narrower affinity reported"
    except Exception as ex:  # noqa: BLE001
        return f"unbound ({type(ex).__name__})"


def host_info():
    """What the CPU arm ran on: affinity, cgroup quota, load — the baseline is only interpretable with these."""
    info = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        info["cgroup_cpu_max"] = "unlimited" if q[0] == "max" else round(int(q[0]) / int(q[1]), 2)
    except Exception:  # noqa: BLE001
        info["cgroup_cpu_max"] = None
    try:
        info["loadavg_1m"] = float(open("/proc/loadavg").read().split()[0])
    except Exception:  # noqa: BLE001
        pass
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                info["cpu_model"] = ln.split(":", 1)[1].strip()
                break
    except Exception:  # noqa: BLE001
        pass
    return info


def ncu_traffic(kernel, recs_per_launch):
    """dram bytes of one launch of `kernel` from the committed ncu --set full capture (profiles/r02_traffic.json),
    scaled to this launch's record count; None when no capture is recorded (ncu cannot run inside the bench)."""
    for name in ("r02_traffic.json",):
        p = os.path.join(ROOT, "profiles", name)
        try:
            t = json.load(open(p))[kernel]
            return (t["dram_bytes_read"] + t["dram_bytes_write"]) * recs_per_launch / t["records_per_launch"], t["source"]
        except Exception:  # noqa: BLE001
            continue
    return None, None


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def block_bytes(batch, table):
    """Byte census of one packed column block: total, fixed-width, Positions / offsets, variable bytes."""
    cols, nrows = batch.columns(table)
    fixed = aux = var = 0
    for c in cols:
        t = c.ch_type
        if t in (abi.CH_STR, abi.CH_LC_STR_ROWS):
            aux += 0 if c.native else nrows * 16
            var += c.buf.nbytes
        elif t in (abi.CH_ARR_STR, abi.CH_ARR_LC_STR):
            aux += nrows * 8 + (0 if c.native else c.pos.nbytes)
            var += c.buf.nbytes
        elif t in (abi.CH_ARR_UINT16, abi.CH_ARR_FLOAT64):
            aux += nrows * 8
            var += c.values.nbytes
        else:
            fixed += nrows * c.elem_size + (nrows if c.nulls is not None else 0)
    return {"total": fixed + aux + var, "fixed": fixed, "aux": aux, "var": var, "rows": nrows}


# ---------------------------------------------------------------------------------------------------------
# CPU side
# ---------------------------------------------------------------------------------------------------------
class CpuPool:
    """`threads` independent CPU decoder instances (the oracle: a C++ restatement of decoder.Run ...
    AppendToColumnBlock), each with its own dictionary view and caches — the reference's goroutine-per-decoder
    ownership.  Built once (in parallel), then fed workloads."""

    def __init__(self, synth, threads):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_binding import Oracle
        self.synth = synth
        ocfg = abi.Cfg()
        ocfg.flow_tag_cache_max_size = 1 << 20
        self.oras = [None] * threads

        def mk(i):
            o = Oracle(ocfg)
            o.load_synth(synth)
            self.oras[i] = o

        self._par(mk)

    def _par(self, fn):
        ts = [threading.Thread(target=fn, args=(i,)) for i in range(len(self.oras))]
        [t.start() for t in ts]
        [t.join() for t in ts]

    def run(self, msg, records_each, seed, flags=0):
        """Generates one workload per instance (untimed), then times all instances decoding concurrently."""
        wls = [None] * len(self.oras)

        def gen(i):
            wls[i] = self.synth.gen(msg, records_each, seed=seed + i, flags=flags)

        self._par(gen)
        t0 = time.perf_counter()
        self._par(lambda i: self.oras[i].submit(wls[i]))
        dt = time.perf_counter() - t0
        n = sum(w.nrecords for w in wls)
        return n / dt, n, dt

    def close(self):
        for o in self.oras:
            o.close()


def cpu_port_run(synth, msg, sample_records, threads, seed, warm=True, flags=0):
    pool = CpuPool(synth, threads)
    if warm:  # steady state, like a long-running ingester: column blocks and caches already sized
        pool.run(msg, sample_records, seed, flags)
    r = pool.run(msg, sample_records, seed, flags)
    pool.close()
    return r


def reference_arm(a):
    """--impl reference: the CPU restatement of the path on every host thread this process may use."""
    from deepflow_b200 import build
    from deepflow_b200.synth import Synth
    build.build_synth()
    build.build_oracle()
    synth = Synth(n_pods=a.pods)
    threads = len(os.sched_getaffinity(0))
    sample = 50000
    t0 = time.time()
    pool = CpuPool(synth, threads)
    setup_s = time.time() - t0
    for w in range(a.warmup):
        pool.run(L7, 4000, 900 + 1000 * w)
    t_all, n_all = 0.0, 0
    for s in range(a.steps):
        _v, n, dt = pool.run(L7, sample, 100_000 + s * 1000)
        t_all += dt
        n_all += n
    pool.close()
    value = n_all / t_all
    per_step = threads * sample
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "records/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * t_all / max(1, a.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32/u64 integer",
            "data": "synthetic",
            "config": {"workload": "C2: L7 HTTP AppProtoLogsData, SmartEncoding on, flow_tag on (BASELINE.json configs[1])",
                       # the GPU arm's workload definition, same keys and values ...
                       "records_per_gpu_per_step": a.records, "sub_batch_records": a.sub_batch, "dictionary_pods": a.pods,
                       "sharding": "vtap_id % n_gpus" if a.gpus > 1 else "none",
                       # ... of which every step of this arm runs a bounded sample
                       "sample_records_per_step": per_step,
                       "note": "a bounded sample of the GPU arm's workload per step: same generator, same dictionary, every "
                               "host thread; the CPU arm does not grow with --gpus (one host)",
                       "setup_seconds": setup_s, "host": host_info()},
            "cpu_baseline": {"value": value, "unit": "records/s", "cores": threads, "kind": "port",
                             "sample": f"{threads} decoder instances x {sample} L7 records per step = {per_step} records "
                                       "(C++ restatement of the Go path; the Go ingester cannot be built here)"},
            "e2e": {"value": value, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------------
class Feed:
    """The sub-batches one decoder pool works through in a step: host (pinned) and device-resident copies."""

    def __init__(self, torch, local, msg, subs):
        self.msg, self.subs = msg, subs
        self.pinned, self.dev = [], []
        for w in subs:
            t = torch.empty(max(1, w.nbytes), dtype=torch.uint8, pin_memory=True)
            t[:w.nbytes].copy_(torch.from_numpy(w.data))
            self.pinned.append(t)
            self.dev.append(t.to(f"cuda:{local}", non_blocking=True))
        self.records = sum(w.nrecords for w in subs)
        self.in_bytes = sum(w.nbytes for w in subs)
        self.frame_bytes = sum(w.frames.nbytes for w in subs)


def run_pools(torch, dist, local, world, pools, steps, warmup, after_step=None, before_step=None):
    """pools: list of (Feed, [Engine...]).  Returns timings of the device-resident and the end-to-end leg."""
    from deepflow_b200.native import lib
    engines = [e for _f, es in pools for e in es]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def worker_lists():
        out = []
        for feed, es in pools:
            for k, e in enumerate(es):
                out.append((e, feed, list(range(k, len(feed.subs), len(es)))))
        return out

    jobs = worker_lists()

    def resident_step():
        err = []

        def work(e, feed, idx):
            try:
                for i in idx:
                    w = feed.subs[i]
                    e.stage_ptr(feed.msg, feed.dev[i].data_ptr(), w.nbytes, w.frames.ctypes.data, len(w.frames))
                    e.run()
            except Exception as ex:  # noqa: BLE001
                err.append(ex)

        ts = [threading.Thread(target=work, args=j) for j in jobs]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if err:
            raise err[0]

    collected = []

    def e2e_step():
        err = []

        def work(e, feed, idx):
            try:
                for i in idx:
                    w = feed.subs[i]
                    out = C.c_void_p()
                    rc = lib().dfgpu_submit(e._h, feed.msg, feed.pinned[i].data_ptr(), w.nbytes, w.frames.ctypes.data,
                                            len(w.frames), None, None, C.byref(out))
                    if rc != 0:
                        raise RuntimeError(f"dfgpu_submit rc={rc}: {lib().dfgpu_last_error(e._h).decode()}")
                    if after_step is not None and feed.msg == L7:
                        collected.append(out)  # the merge reads the emitted keys, then releases
                    else:
                        lib().dfgpu_batch_release(out)
            except Exception as ex:  # noqa: BLE001
                err.append(ex)

        if before_step is not None:
            before_step()
        ts = [threading.Thread(target=work, args=j) for j in jobs]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if err:
            raise err[0]
        if after_step is not None:
            after_step(collected)
            collected.clear()

    streams = [torch.cuda.ExternalStream(e.stream, device=local) for e in engines]
    tstream = torch.cuda.Stream(device=local)
    for _ in range(warmup):
        if before_step is not None:
            before_step()
        resident_step()
    for e in engines:
        e.sync()
    launches0 = sum(e.launch_count for e in engines)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(tstream)
    for st in streams:
        st.wait_event(ev0)
    for _ in range(steps):
        if before_step is not None:
            before_step()
        resident_step()
    for st in streams:
        done = torch.cuda.Event()
        done.record(st)
        tstream.wait_event(done)
    ev1.record(tstream)
    for e in engines:
        e.sync()
    tstream.synchronize()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = sum(e.launch_count for e in engines) - launches0
    # every resident run must have fit and decoded: fetch the last batch of every engine
    for e, feed, idx in jobs:
        if idx:
            b = e.fetch()
            cnt = b.counters()
            assert cnt.error_count == 0, (cnt.error_count, cnt.count)
            b.release()
    for _ in range(warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    for e in engines:
        e.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    return dev_ms, e2e_s, launches


def solo_stages(engine, feed):
    """One decoder alone over its feed: the library's CUDA events around every stage (kernels of concurrent
    decoders would overlap the events)."""
    engine.sync()
    engine.profile_enable(True)
    for i, w in enumerate(feed.subs):
        engine.stage_ptr(feed.msg, feed.dev[i].data_ptr(), w.nbytes, w.frames.ctypes.data, len(w.frames))
        engine.run()
    engine.sync()
    st, runs = engine.profile_stages()
    pack_ms, pack_n, _ = engine.profile_read()
    engine.profile_enable(False)
    engine.fetch().release()
    return st, runs, pack_ms, pack_n


def reduce_max(torch, dist, world, local, vals):
    t = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def reduce_sum(torch, dist, world, local, vals):
    t = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def make_merge_hooks(torch, local, world, engines):
    """The one collective of the path: after an e2e step all ranks all-gather the flow_tag keys they emitted
    (NCCL), the lowest rank keeps the row, every decoder imports the union."""
    from deepflow_b200.native import Batch
    from deepflow_b200.sharded import FLOW_TAG_TABLES, TagMerge
    merge = TagMerge(device=torch.device("cuda", local))
    emitted = [0]

    def after(handles):
        fps, tms = [], []
        for h in handles:
            b = Batch(h)
            for t in FLOW_TAG_TABLES:
                f = b.tag_fingerprints(t)
                if len(f):
                    col, _ = b.column(t, 0)  # the `time` column
                    fps.append(f)
                    tms.append(col.values.astype(np.uint32))
            b.release()
        f = np.concatenate(fps) if fps else np.zeros((0, 2), np.uint64)
        t = np.concatenate(tms) if tms else np.zeros(0, np.uint32)
        emitted[0] += len(t)
        if world > 1:
            # keys stay on the GPU: NCCL all-gather, ownership on the device, every decoder of this rank imports
            # the union (its own keys included: the rank's other decoders have not seen them) from device memory
            t0 = time.perf_counter()
            _keep, others, mine = merge.exchange_device(f, t)
            allk = torch.cat([others, mine]) if len(mine) else others
            if len(allk):
                kf = allk[:, 0:2].contiguous()
                kt = allk[:, 2].to(torch.int32).contiguous()
                torch.cuda.current_stream().synchronize()
                for e in engines:
                    e.tags_import_ptr(kf.data_ptr(), kt.data_ptr(), len(kt))
            merge.seconds += time.perf_counter() - t0

    return merge, after, emitted


def roofline_for(stage_ms, runs, alg_bytes_by_stage, recs_per_launch, peak, peak_src, kernels):
    """Dominant stage = the one with the most device time in the solo leg."""
    stage_ms = {k: v for k, v in stage_ms.items()}
    dom = max(alg_bytes_by_stage, key=lambda k: stage_ms.get(k, 0.0))
    per_launch_s = stage_ms[dom] / 1000.0 / max(1, runs)
    achieved = alg_bytes_by_stage[dom] * recs_per_launch / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
    traffic, tsrc = ncu_traffic(kernels.get(dom, dom), recs_per_launch)
    tot = sum(stage_ms.values())
    return {"bound": "hbm", "kernel": kernels.get(dom, dom), "stage": dom, "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
            "traffic_source": tsrc, "peak_source": peak_src, "algorithmic_bytes_per_record": alg_bytes_by_stage[dom],
            "launches_timed": runs, "avg_launch_ms": stage_ms[dom] / max(1, runs),
            "timed": "one decoder alone, CUDA events of the library on its stream around every stage",
            "share_of_device_time": stage_ms[dom] / tot if tot else None,
            "stage_ms_per_step_one_decoder": stage_ms,
            "stage_frac": {k: (alg_bytes_by_stage[k] * recs_per_launch / (stage_ms[k] / 1000.0 / max(1, runs)) / 1e9 / peak
                               if stage_ms.get(k) else None) for k in alg_bytes_by_stage}}


def column_hashes(batch, tables):
    """One digest per column of every table: the bytes of every array of the ch-go container."""
    out = {}
    for t in tables:
        cols, n = batch.columns(t)
        for c in cols:
            h = hashlib.blake2b(digest_size=12)
            h.update(str(n).encode())
            for attr in ("values", "buf", "pos", "arr_offsets", "nulls"):
                x = getattr(c, attr)
                if x is not None:
                    h.update(attr.encode())
                    h.update(np.ascontiguousarray(x).tobytes())
            if c.dict is not None:
                for s in c.dict:
                    h.update(b"|" + s)
            out[f"{t}:{c.name}"] = h.hexdigest()
    return out


def leg_parity(a, synth, local, cfg, subs):
    """Parity AT the benchmarked configuration: full-size sub-batches, the bench's dictionary, `decoders` contexts
    concurrently on one GPU, every column of the four L7 tables hashed and compared with the oracle run on the
    same sub-batch from the same (empty) state."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from deepflow_b200.native import Engine
    from oracle_binding import Oracle
    tables = [abi.TABLE_L7_FLOW_LOG, abi.TABLE_L7_CUSTOM_FIELD, abi.TABLE_L7_CUSTOM_FIELD_VALUE, abi.TABLE_L7_APP_SERVICE]
    n = max(1, min(a.decoders, len(subs)))
    engs = [Engine(local, cfg) for _ in range(n)]
    oras = [Oracle(cfg) for _ in range(n)]
    for e, o in zip(engs, oras):
        e.load_synth(synth)
        o.load_synth(synth)
    g, o_h = [None] * n, [None] * n
    err = []

    def gpu(k):
        try:
            b = engs[k].submit(subs[k])
            g[k] = (column_hashes(b, tables), b.counters().as_dict())
            b.release()
        except Exception as ex:  # noqa: BLE001
            err.append(ex)

    def cpu(k):
        try:
            oras[k].submit(subs[k])

            class _B:
                def columns(self, t):
                    return oras[k].columns(t)
            o_h[k] = (column_hashes(_B(), tables), oras[k].counters().as_dict())
        except Exception as ex:  # noqa: BLE001
            err.append(ex)

    t0 = time.perf_counter()
    ts = [threading.Thread(target=gpu, args=(k,)) for k in range(n)] + [threading.Thread(target=cpu, args=(k,)) for k in range(n)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    if err:
        raise err[0]
    mism, ncols = [], 0
    for k in range(n):
        ncols += len(o_h[k][0])
        for name, hv in o_h[k][0].items():
            if g[k][0].get(name) != hv:
                mism.append(f"ctx{k}:{name}")
        for key in ("count", "error_count", "l7_http_count", "new_field_count", "new_field_value_count"):
            if g[k][1].get(key) != o_h[k][1].get(key):
                mism.append(f"ctx{k}:counter:{key}")
    digest = hashlib.blake2b("".join(sorted(f"{k}{n_}{h}" for k in range(n) for n_, h in g[k][0].items())).encode(),
                             digest_size=12).hexdigest()
    for e in engs:
        e.close()
    for o in oras:
        o.close()
    return {"contexts": n, "records_per_context": [s.nrecords for s in subs[:n]], "dictionary_pods": a.pods,
            "tables": len(tables), "columns_compared": ncols, "mismatches": mism[:20], "ok": not mism,
            "digest": digest, "seconds": time.perf_counter() - t0}


def leg_c3(a, torch, local, peak, peak_src):
    """BASELINE configs[2]: flow_metrics 1s -> 1m rollup, N decoded network_map.1s rows -> K groups, on one GPU,
    input resident in HBM (generated on the device).  Checks: every summed meter field totals the same before
    and after the fold, every max field has the same maximum, the number of groups equals the number of distinct
    keys (size-independent properties; bit-exact parity against the oracle is tests/test_metrics_gpu.py)."""
    from deepflow_b200.native import Engine
    table = abi.TABLE_METRICS_BASE + 6  # network_map.1s
    cfg = abi.Cfg()
    cfg.max_batch_bytes = 1 << 20
    cfg.max_batch_records = 4096
    cfg.rollup_max_groups = max(1 << 20, 2 * a.c3_groups)
    eng = Engine(local, cfg)
    elem, kind, nf = eng.rollup_describe(table)
    n, K = a.c3_rows, a.c3_groups
    dev = f"cuda:{local}"
    dts = {1: torch.int8, 2: torch.int16, 4: torch.int32, 8: torch.int64}
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    t0 = time.time()
    raw = torch.empty((n, nf), dtype=torch.int64, device=dev)
    group = torch.empty(n, dtype=torch.int64, device=dev)
    CH = 4_000_000
    for s in range(0, n, CH):
        m = min(CH, n - s)
        u = torch.rand(m, device=dev, generator=g, dtype=torch.float64)
        group[s:s + m] = torch.clamp((float(K) ** u).to(torch.int64) - 1, 0, K - 1)  # Zipf(1.0): rank = K^u
        v = torch.randint(0, 1 << 31, (m, nf), device=dev, generator=g, dtype=torch.int64)  # every field width holds it
        small = torch.rand((m, nf), device=dev, generator=g) < 0.9
        v = torch.where(small, v & 0xffff, v)
        v = torch.where(torch.rand((m, nf), device=dev, generator=g) < 0.6, torch.zeros_like(v), v)  # 60 % zeros
        raw[s:s + m] = v
        del u, v, small
    cols = []
    base_time = 1_700_000_040  # a minute boundary: 60 consecutive seconds -> one 1 m bucket
    for c, (es, kd) in enumerate(zip(elem, kind)):
        if c == 0:
            col = (base_time + (torch.arange(n, device=dev, dtype=torch.int64) % 60)).to(torch.int32)
        elif kd == 1:  # IPv6: 16 bytes per row, a function of the group
            col = torch.stack([group * 0x9E3779B1 + c, group * 0x85EBCA77 + 7 * c], dim=1).contiguous()
        elif kd == 2:
            col = (group % 17).to(torch.int8)  # canonical tap-side code 0..16
        else:
            col = ((group + 1) * (2654435761 + 40503 * c) >> 7).to(dts[es])  # any function of the group
        cols.append(col)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    key_bytes = sum(16 if kd == 1 else es for es, kd in zip(elem, kind))
    distinct = int(torch.unique(group).numel())
    stream = torch.cuda.ExternalStream(eng.stream, device=local)
    ptrs = [c.data_ptr() for c in cols]
    ms, out_rows, checks = [], None, {}
    for it in range(1 + a.steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.rollup_add_rows(table, 60, ptrs, raw.data_ptr(), n)
        e1.record(stream)
        b = eng.rollup_flush(table)  # D2H of the K groups: outside the timed fold
        torch.cuda.synchronize()
        if it:
            ms.append(e0.elapsed_time(e1))
        rcols, out_rows = b.columns(abi.TABLE_METRICS_REDUCED)
        if it == a.steps:
            by = {c.name: c for c in rcols}
            names = [c.name for c in rcols]
            # size-independent properties on fields the columns hold losslessly (integers)
            checks["groups_equal_distinct_keys"] = bool(out_rows == distinct)
            tot_in = raw.sum(dim=0)  # int64 wrap = uint64 wrap
            for fname, fidx in (("packet_tx", 0), ("byte_tx", 2), ("new_flow", 8)):
                if fname in by:
                    got = int(by[fname].values.astype(np.uint64).sum(dtype=np.uint64))
                    exp = int(tot_in[fidx].item()) & 0xffffffffffffffff
                    checks[f"sum_{fname}"] = bool(got == exp)
            if "rtt_max" in by:
                checks["max_rtt_max"] = bool(int(by["rtt_max"].values.max()) == int(raw[:, 15].max().item() & 0xffffffff))
            checks["columns"] = len(names)
        b.release()
    t = float(np.mean(ms)) / 1000.0
    alg = n * (key_bytes + nf * 8) + out_rows * (key_bytes + nf * 8)
    achieved = alg / t / 1e9
    res = {"workload": "C3: flow_metrics 1s -> 1m rollup (BASELINE.json configs[2])", "rows": n, "groups": out_rows,
           "distinct_keys_in_input": distinct, "key_bytes": key_bytes, "meter_fields": nf, "value": n / t, "unit": "rows/s",
           "ms_per_step": t * 1000, "input": "decoded rows resident in HBM (generated on the device)", "gen_seconds": gen_s,
           "roofline": {"bound": "hbm", "kernel": "k_ru_key + k_ru_hist/scatter + k_ru_fold + k_ru_edge (rollup.cu)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                        "algorithmic_bytes": alg, "multiple_of_bound": (t * peak * 1e9) / alg, "peak_source": peak_src},
           "checks": checks,
           "e2e": {"value": n / t, "unit": "rows/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": int(out_rows * (key_bytes + nf * 8)),
                   "note": "config 3 keeps its input in HBM by definition; the D2H of the K groups is the flush"}}
    if not a.no_cpu_baseline:  # one thread, one hash map, on the first 2 M rows of the same input
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_binding
        L = oracle_binding.lib()
        L.oracle_rollup_rows.restype = C.c_uint64
        L.oracle_rollup_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        m = min(n, 2_000_000)
        keys = np.concatenate([c[:m].cpu().numpy().reshape(m, -1).view(np.uint8) for c in cols], axis=1)
        keys[:, 0:4] = np.frombuffer(np.uint32(base_time).tobytes(), np.uint8)  # time -> bucket start
        keys = np.ascontiguousarray(keys)
        rawh = np.ascontiguousarray(raw[:m].cpu().numpy())
        ops = np.zeros(nf, np.uint8)  # sums: the cheapest rule (a lower bound of the CPU cost)
        sums = np.zeros(nf, np.uint64)
        t0 = time.perf_counter()
        kk = L.oracle_rollup_rows(keys.ctypes.data, keys.shape[1], rawh.ctypes.data, nf, ops.ctypes.data, m, sums.ctypes.data)
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": m / dt, "unit": "rows/s", "cores": 1, "kind": "port",
                               "sample": f"{m} decoded rows of the same input -> {kk} groups, one thread, one hash map, {dt:.1f} s"}
    del raw, group, cols
    eng.close()
    torch.cuda.empty_cache()
    return res


def main():
    a = parse_args()
    rank, world, local = dist_env()
    if world > 1:
        a.gpus = world
    if a.impl == "reference":
        return reference_arm(a) if rank == 0 else 0
    legs = set(x for x in a.legs.split(",") if x)

    import torch
    import torch.distributed as dist
    from deepflow_b200 import build
    from deepflow_b200.synth import F_HIGH_CARD, Synth
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: libdfgpu has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local)
    numa = bind_near_gpu(local)  # before any pinned allocation: first touch places the staging buffers
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: the JSON line stands alone
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        build.build_synth()
    from deepflow_b200.native import Engine
    peak, peak_src = measured_peak_hbm()
    sampler = ClockSampler(local)
    sampler.start()
    synth = Synth(n_pods=a.pods)

    def mk_cfg(max_bytes, cache=1 << 20, native=0, lz4=0):
        c = abi.Cfg()
        c.max_batch_bytes = max_bytes + (1 << 20)
        c.max_batch_records = a.sub_batch + a.sub_batch // 8 + 4096
        c.flow_tag_cache_max_size = cache  # the tag set of the C2 workload stays below it: no eviction either side
        c.native_strings = native
        c.block_lz4 = lz4
        return c

    # =================================================================================== C2 (headline)
    t0 = time.time()
    subs = gen_sub_batches(synth, L7, a.records, a.sub_batch, 1000 + 7919 * rank, world, rank)
    gen_s = time.time() - t0
    feed = Feed(torch, local, L7, subs)
    torch.cuda.synchronize()
    MODES = {"containers": (0, 0), "native": (1, 0), "lz4": (1, 1)}
    MODE_WHAT = {
        "containers": "ch-go containers: String columns as Buf + one 16-byte Position per value",
        "native": "dfgpu_cfg.native_strings=1: String / Array(String) columns leave the GPU as ClickHouse native column data "
                  "(uvarint length + bytes), no Position arrays over PCIe",
        "lz4": "dfgpu_cfg.native_strings=1 + block_lz4=1: every buffer of l7_flow_log crosses PCIe as ClickHouse compressed "
               "frames (LZ4 per 4 KiB, compressed on the GPU inside the timed region; checksum slots left for the host)"}
    cfg_plain = mk_cfg(max(w.nbytes for w in subs))  # container mode: byte census, parity leg
    cfg = mk_cfg(max(w.nbytes for w in subs), native=MODES[a.e2e_mode][0], lz4=MODES[a.e2e_mode][1])
    engines = [Engine(local, cfg) for _ in range(max(1, a.decoders))]
    for e in engines:
        e.load_synth(synth)
    merge, after, emitted = (None, None, [0])
    if world > 1:
        merge, after, emitted = make_merge_hooks(torch, local, world, engines)
    dev_ms, e2e_s, launches = run_pools(torch, dist, local, world, [(feed, engines)], a.steps, a.warmup, after_step=after)
    st_ms, st_runs, pack_ms, pack_n = solo_stages(engines[0], feed)
    def mode_bytes(engine, mode):
        """D2H bytes per row of l7_flow_log in an output mode (one sub-batch through `engine`)."""
        bb = engine.submit(subs[-1])
        if mode == "lz4":
            fb, fr = bb.frame_bytes(abi.TABLE_L7_FLOW_LOG)
            out = fb / max(1, fr)
        else:
            c = block_bytes(bb, abi.TABLE_L7_FLOW_LOG)
            out = c["total"] / max(1, c["rows"])
        bb.release()
        return out

    ceng = Engine(local, cfg_plain)
    ceng.load_synth(synth)
    b = ceng.submit(subs[-1])
    cen = block_bytes(b, abi.TABLE_L7_FLOW_LOG)
    b.release()
    ceng.close()
    row_b = cen["total"] / max(1, cen["rows"])
    e2e_row_b = row_b if a.e2e_mode == "containers" else mode_bytes(engines[0], a.e2e_mode)
    fixed_b, aux_b, var_b = (cen[k] / max(1, cen["rows"]) for k in ("fixed", "aux", "var"))
    wire_b = feed.in_bytes / feed.records
    kg_b = 116.0  # the 39 SmartEncoding columns + team_id / gprocess ids (schema_l7.h)
    alg = {"decode": wire_b + 4 + (fixed_b - kg_b), "pack": 2 * var_b + aux_b, "smartencoding": kg_b + 32 * 12, "flow_tag": 32 * 8.0}
    kernels = {"decode": "k_l7_decode", "pack": "k_l7_strings<0> + k_l7_strings<1>", "smartencoding": "k_l7_kg",
               "flow_tag": "k_l7_flow_tag<0..3>"}
    recs_per_launch = feed.records / max(1, st_runs)
    roof = roofline_for(st_ms, st_runs, alg, recs_per_launch, peak, peak_src, kernels)
    path_bytes = wire_b + 4 + row_b + 32 * 12 + 32 * 8
    path_ms = sum(st_ms.values())
    path_ach = path_bytes * feed.records / (path_ms / 1000.0) / 1e9 if path_ms else 0.0
    roof["path"] = {"kernels": "framing + decode + scans + pack + SmartEncoding + flow_tag (every kernel of dfgpu_run)",
                    "algorithmic_bytes_per_record": path_bytes, "ms_per_step": path_ms, "achieved": path_ach,
                    "frac": path_ach / peak}
    dev_ms_max, e2e_ms_max = reduce_max(torch, dist, world, local, [dev_ms, e2e_s * 1000.0])
    (total_recs,) = reduce_sum(torch, dist, world, local, [float(feed.records)])
    value = total_recs * a.steps / (dev_ms_max / 1000.0)
    e2e_value = total_recs * a.steps / (e2e_ms_max / 1000.0)
    line = {"metric": METRIC, "value": value, "unit": "records/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u32/u64 integer (f64 casts only at pack)", "data": "synthetic",
            "config": {"workload": "C2: L7 HTTP AppProtoLogsData, SmartEncoding on, flow_tag on (BASELINE.json configs[1])",
                       "records_per_gpu_per_step": a.records, "sub_batch_records": a.sub_batch, "dictionary_pods": a.pods,
                       "sharding": "vtap_id % n_gpus" if a.gpus > 1 else "none",
                       "l2": "inputs 4+ GB per step >> 126 MB L2 (no flush needed)", "mean_wire_bytes": wire_b,
                       "packed_bytes_per_row": row_b, "gen_seconds": gen_s, "decoders_per_gpu": len(engines),
                       "host_affinity": numa,
                       "tag_merge": (None if merge is None else {
                           "collective": "all_gather of emitted flow_tag key fingerprints, once per e2e step (NCCL)",
                           "bytes_sent_rank0": merge.bytes_sent, "seconds_rank0": merge.seconds,
                           "keys_emitted_rank0": emitted[0]})},
            "e2e": {"value": e2e_value, "unit": "records/s", "h2d_bytes_per_step": int(feed.in_bytes + feed.frame_bytes),
                    "d2h_bytes_per_step": int(e2e_row_b * feed.records), "ms_per_step": e2e_ms_max / a.steps,
                    "d2h_gbs_per_rank": e2e_row_b * feed.records * a.steps / e2e_s / 1e9,
                    "h2d_gbs_per_rank": (feed.in_bytes + feed.frame_bytes) * a.steps / e2e_s / 1e9,
                    "mode": a.e2e_mode, "what": MODE_WHAT[a.e2e_mode], "d2h_bytes_per_row": e2e_row_b,
                    "container_bytes_per_row": row_b},
            "gpu_launches": int(launches), "roofline": roof, "legs": {}}

    # one decoder alone against all of them: how much of the GPU a single context keeps busy
    if "c2" in legs and world == 1:
        d1, _e1, _l1 = run_pools(torch, dist, local, world, [(feed, engines[:1])], max(1, a.steps // 2), 1)
        line["config"]["one_decoder_value"] = feed.records * max(1, a.steps // 2) / (d1 / 1000.0)

    # ---- the same e2e call in the other output modes (N=1 only: the scaling runs keep the measured legs short) ----
    if "modes" in legs and world == 1:
        for mode in MODES:
            if mode == a.e2e_mode:
                continue
            for e in engines:
                e.sync()
            mcfg = mk_cfg(max(w.nbytes for w in subs), native=MODES[mode][0], lz4=MODES[mode][1])
            meng = [Engine(local, mcfg) for _ in range(len(engines))]
            for e in meng:
                e.load_synth(synth)
            _d, m_s, _l = run_pools(torch, dist, local, world, [(feed, meng)], a.steps, max(1, a.warmup - 1))
            mb = mode_bytes(meng[0], mode)
            for e in meng:
                e.close()
            key = {"containers": "e2e_containers", "native": "e2e_native_strings", "lz4": "e2e_lz4"}[mode]
            line[key] = {"value": feed.records * a.steps / m_s, "unit": "records/s",
                         "h2d_bytes_per_step": int(feed.in_bytes + feed.frame_bytes), "d2h_bytes_per_step": int(mb * feed.records),
                         "d2h_bytes_per_row": mb, "ms_per_step": m_s * 1000 / a.steps, "what": MODE_WHAT[mode]}

    # =================================================================================== parity at size
    if "parity" in legs and world == 1:
        # container mode: the same kernels up to the string emission; the native / LZ4 output modes are compared
        # with the oracle by tests/test_l7_native_gpu.py and tests/test_lz4_gpu.py
        line["legs"]["parity"] = leg_parity(a, synth, local, cfg_plain, subs)

    if not a.no_cpu_baseline and rank == 0 and world == 1:
        build.build_oracle()
        v, n, dt = cpu_port_run(synth, L7, 1_500_000, 1, 4242)
        line["cpu_baseline"] = {"value": v, "unit": "records/s", "cores": 1, "kind": "port",
                                "sample": f"{n} L7 records of the same generator, one decoder instance, {dt:.1f} s",
                                "host": host_info()}
    for e in engines:
        e.close()
    del feed
    torch.cuda.empty_cache()

    # =================================================================================== C4 mixed firehose
    if "c4" in legs:
        n4, n7, nm = int(a.records * 0.50), int(a.records * 0.35), a.records - int(a.records * 0.50) - int(a.records * 0.35)
        t0 = time.time()
        f4 = Feed(torch, local, L4, gen_sub_batches(synth, L4, n4, a.sub_batch, 4000 + 7919 * rank, world, rank))
        f7 = Feed(torch, local, L7, gen_sub_batches(synth, L7, n7, a.sub_batch, 5000 + 7919 * rank, world, rank))
        fm = Feed(torch, local, MET, gen_sub_batches(synth, MET, nm, a.sub_batch, 6000 + 7919 * rank, world, rank))
        gen4 = time.time() - t0
        mx = max(w.nbytes for f in (f4, f7, fm) for w in f.subs)
        c4cfg = mk_cfg(mx)  # flow_metrics tables: containers (block_lz4 covers the flow_log tables only)
        c4log = mk_cfg(mx, native=MODES[a.e2e_mode][0], lz4=MODES[a.e2e_mode][1])  # l4 / l7: the headline output mode
        pools = [(f4, [Engine(local, c4log) for _ in range(2)]), (f7, [Engine(local, c4log) for _ in range(2)]),
                 (fm, [Engine(local, c4cfg)])]
        c4plain = Engine(local, c4cfg)  # byte census of the packed blocks
        c4plain.load_synth(synth)
        all_e = [e for _f, es in pools for e in es]
        for e in all_e:
            e.load_synth(synth)
        m4, after4, em4 = (None, None, [0])
        if world > 1:
            m4, after4, em4 = make_merge_hooks(torch, local, world, pools[1][1])
        d_ms, e_s, l4n = run_pools(torch, dist, local, world, pools, a.steps, a.warmup, after_step=after4)
        per_type = {}
        tot_alg = 0.0
        for name, (f, es) in zip(("l4", "l7", "metrics"), pools):
            st, runs, _pm, _pn = solo_stages(es[0], f)
            bb = c4plain.submit(f.subs[-1])
            if f.msg == MET:
                rb = sum(block_bytes(bb, abi.TABLE_METRICS_BASE + t)["total"] for t in range(9)) / max(1, f.subs[-1].nrecords)
            else:
                c = block_bytes(bb, MAIN_TABLE[f.msg])
                rb = c["total"] / max(1, c["rows"])
            bb.release()
            d2h_rb = rb
            if f.msg != MET and a.e2e_mode != "containers":
                bb = es[0].submit(f.subs[-1])
                if a.e2e_mode == "lz4":
                    fb, fr = bb.frame_bytes(MAIN_TABLE[f.msg])
                    d2h_rb = fb / max(1, fr)
                else:
                    c = block_bytes(bb, MAIN_TABLE[f.msg])
                    d2h_rb = c["total"] / max(1, c["rows"])
                bb.release()
            algb = f.in_bytes / f.records + 4 + rb + 32 * (12 if f.msg != MET else 8)
            ms_solo = sum(st.values())
            ach = algb * f.records / (ms_solo / 1000.0) / 1e9 if ms_solo else 0.0
            per_type[name] = {"records_per_step": f.records, "mean_wire_bytes": f.in_bytes / f.records, "packed_bytes_per_record": rb,
                              "d2h_bytes_per_record": d2h_rb,
                              "algorithmic_bytes_per_record": algb, "solo_ms_per_step": ms_solo, "solo_records_per_s": f.records / (ms_solo / 1000.0) if ms_solo else None,
                              "stage_ms": st, "achieved_gbs": ach, "frac": ach / peak}
            tot_alg += algb * f.records
        recs = f4.records + f7.records + fm.records
        d_mx, e_mx = reduce_max(torch, dist, world, local, [d_ms, e_s * 1000.0])
        (recs_all,) = reduce_sum(torch, dist, world, local, [float(recs)])
        v4 = recs_all * a.steps / (d_mx / 1000.0)
        ach4 = tot_alg * a.steps / (d_ms / 1000.0) / 1e9
        dom = max(per_type, key=lambda k: per_type[k]["solo_ms_per_step"])
        h2d = sum(f.in_bytes + f.frame_bytes for f in (f4, f7, fm))
        d2h = sum(per_type[k]["d2h_bytes_per_record"] * per_type[k]["records_per_step"] for k in per_type)
        leg = {"workload": "C4: mixed firehose 50 % L4 / 35 % L7 / 15 % metric by record count (BASELINE.json configs[3]), "
                           "one decoder pool per message type (2 + 2 + 1 contexts per GPU), sharded by vtap_id % n_gpus",
               "records_per_gpu_per_step": recs, "value": v4, "unit": "records/s", "ms_per_step": d_mx / a.steps,
               "e2e": {"value": recs_all * a.steps / (e_mx / 1000.0), "unit": "records/s", "h2d_bytes_per_step": int(h2d),
                       "d2h_bytes_per_step": int(d2h), "ms_per_step": e_mx / a.steps,
                       "mode": f"l4 / l7 blocks: {a.e2e_mode}; flow_metrics tables: containers"},
               "gpu_launches": int(l4n), "gen_seconds": gen4, "per_type": per_type,
               "roofline": {"bound": "hbm", "kernel": f"the {dom} pipeline (largest share of device time)",
                            "achieved": ach4, "peak": peak, "unit": "GB/s", "frac": ach4 / peak, "traffic": None,
                            "what": "algorithmic bytes of all three message types per step / device time of the step with every "
                                    "decoder pool running concurrently (rank 0)", "peak_source": peak_src},
               "tag_merge": None if m4 is None else {"bytes_sent_rank0": m4.bytes_sent, "seconds_rank0": m4.seconds,
                                                     "keys_emitted_rank0": em4[0]}}
        if not a.no_cpu_baseline and rank == 0 and world == 1:
            parts = []
            tsec = 0.0
            for msg, share in ((L4, 0.5), (L7, 0.35), (MET, 0.15)):
                nn = int(200_000 * share)
                _v, n_, dt = cpu_port_run(synth, msg, nn, 1, 777)
                parts.append(n_)
                tsec += dt
            leg["cpu_baseline"] = {"value": sum(parts) / tsec, "unit": "records/s", "cores": 1, "kind": "port",
                                   "sample": f"{parts} L4/L7/metric records of the same generator, one decoder instance each, {tsec:.1f} s"}
        line["legs"]["c4_mixed"] = leg
        c4plain.close()
        for e in all_e:
            e.close()
        del f4, f7, fm, pools
        torch.cuda.empty_cache()

    # =================================================================================== C5 high cardinality
    if "c5" in legs:
        t0 = time.time()
        big = Synth(n_pods=a.c5_pods)
        synth_s = time.time() - t0
        t0 = time.time()
        subs5 = gen_sub_batches(big, L7, a.c5_records, a.sub_batch, 9000 + 7919 * rank, world, rank, flags=F_HIGH_CARD)
        gen5 = time.time() - t0
        f5 = Feed(torch, local, L7, subs5)
        c5cfg = mk_cfg(max(w.nbytes for w in subs5), cache=1 << 22)
        e5 = [Engine(local, c5cfg) for _ in range(max(1, a.decoders))]
        t0 = time.time()
        for e in e5:
            e.load_synth(big)
        dict_s = (time.time() - t0) / len(e5)
        m5, after5, em5 = make_merge_hooks(torch, local, world, e5)

        def cold():
            for e in e5:
                e.tags_reset()  # every step starts from an empty flow_tag set: all keys are new, all are merged

        d_ms, e_s, l5 = run_pools(torch, dist, local, world, [(f5, e5)], a.steps, max(1, a.warmup - 1), after_step=after5,
                                  before_step=cold)
        cold()
        st5, runs5, _pm, _pn = solo_stages(e5[0], f5)
        bb = e5[0].submit(subs5[0])
        cnt = bb.counters()
        unrec, evi = int(cnt.flow_tag_unrecorded), int(cnt.flow_tag_evicted)
        bb.release()
        d_mx, e_mx = reduce_max(torch, dist, world, local, [d_ms, e_s * 1000.0])
        recs_all, keys_all, bytes_all = reduce_sum(torch, dist, world, local, [float(f5.records), float(em5[0]), float(m5.bytes_sent)])
        (merge_s,) = reduce_max(torch, dist, world, local, [m5.seconds])
        steps_e2e = a.steps + max(1, a.warmup - 1)
        line["legs"]["c5_highcard"] = {
            "workload": "C5: high-cardinality SmartEncoding (BASELINE.json configs[4]): P = %d pod dictionary, L7 records whose "
                        "attributes draw from ~1 M (name, value) pairs, flow_tag set emptied before every step" % a.c5_pods,
            "records_per_gpu_per_step": f5.records, "dictionary_pods": a.c5_pods, "dict_build_seconds": synth_s,
            "dict_load_seconds_per_context": dict_s, "gen_seconds": gen5,
            "value": recs_all * a.steps / (d_mx / 1000.0), "unit": "records/s", "ms_per_step": d_mx / a.steps,
            "e2e": {"value": recs_all * a.steps / (e_mx / 1000.0), "unit": "records/s",
                    "h2d_bytes_per_step": int(f5.in_bytes + f5.frame_bytes), "d2h_bytes_per_step": None, "ms_per_step": e_mx / a.steps},
            "flow_tag_keys_emitted_per_step_all_ranks": keys_all / steps_e2e,
            "merge": {"collective": "all_gather of (128-bit fingerprint, time) of every emitted flow_tag row, once per step (NCCL)",
                      "bytes_per_step_all_ranks": bytes_all / steps_e2e, "seconds_per_step_max_rank": merge_s / steps_e2e,
                      "share_of_e2e_step": (merge_s / steps_e2e) / (e_mx / 1000.0 / a.steps) if e_mx else None},
            "flow_tag_unrecorded": unrec, "flow_tag_evicted": evi, "gpu_launches": int(l5),
            "roofline": {"bound": "hbm", "kernel": "k_l7_flow_tag<0..3> (cold set: every key is inserted and emitted)",
                         "stage_ms_per_step_one_decoder": st5, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                         "traffic": None,
                         "what": "the flow_tag stage against a cold set next to the other stages of the same run; "
                                 "compare with the C2 line's warm-set stage times"}}
        for e in e5:
            e.close()
        del f5
        big.close()
        torch.cuda.empty_cache()

    # =================================================================================== C3 rollup
    if "c3" in legs and world == 1:
        try:
            line["legs"]["c3_rollup"] = leg_c3(a, torch, local, peak, peak_src)
        except torch.OutOfMemoryError as ex:  # noqa: BLE001
            line["legs"]["c3_rollup"] = {"error": f"out of device memory for {a.c3_rows} rows: {ex}"}

    line["clocks"] = sampler.stop()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
