#!/usr/bin/env python3
"""bench.py — flow records/sec of the ingest hot path (decode -> SmartEncoding ->
flow_tag -> ClickHouse column pack) on B200, per BASELINE.json.

Workload at N=1 (BASELINE.json configs[1]): 10 M synthetic L7 HTTP
AppProtoLogsData records, submitted the way the ingester's decoder sees them:
sub-batches of <= 1024 agent frames (decoder.go:56 BUFFER_SIZE x 256 KiB frames).
One "step" = one pass over all sub-batches.  Under torchrun every rank owns the
vtap_id % N == rank shard of the firehose (weak scaling: 10 M records per GPU);
per-record work has no collective — the only exchange is the all-gather of newly
emitted flow_tag keys after every e2e step (deepflow_b200/sharded.py).

  value     device-resident: inputs already in HBM, kernels only; every decoder context of
            the GPU (default 3, like the ingester's decoder goroutines) runs its share of
            the sub-batches on its own stream, timed by CUDA events on a timing stream that
            waits for all of them
  e2e       through the C ABI with HOST buffers: pinned H2D of every sub-batch, all
            kernels, D2H of every column block, inside the timed region
  roofline  the dominant kernel (EMIT/pack pass), timed live in a separate leg where ONE
            decoder runs alone (CUDA events on its stream around each launch):
            algorithmic bytes / average launch time vs the measured HBM peak; `traffic` is
            the DRAM bytes of one launch from the committed ncu capture (profiles/);
            `path` = the same for SIZE + EMIT + SmartEncoding over SURVEY §8(d)'s bytes
  cpu_baseline   the CPU oracle (a C++ restatement of the Go path) on a bounded sample

--impl reference times that CPU restatement with all host threads instead (the Go
reference cannot be built in this image: no Go toolchain).
"""
import argparse
import ctypes as C
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

MSG = abi.DFGPU_MSG_PROTOCOLLOG
METRIC = "flow records/sec ingest+SmartEncode (L7 decode + tag-inject + flow_tag + column pack)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dfgpu", choices=["dfgpu", "reference"])
    ap.add_argument("--records", type=int, default=10_000_000, help="records per GPU per step")
    ap.add_argument("--sub-batch", type=int, default=625_000, help="records per dfgpu_submit")
    ap.add_argument("--pods", type=int, default=100_000, help="dictionary size P (SURVEY §8d)")
    ap.add_argument("--decoders", type=int, default=3, help="contexts per GPU in the e2e leg (DecoderQueueCount)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-native-leg", action="store_true", help="skip the second e2e leg (native_strings=1)")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def gen_sub_batches(synth, n_total, sub, seed0, world, rank, threads=16):
    """Generate ceil(n_total/sub) sub-batches in parallel threads (ctypes drops the GIL)."""
    sizes = [min(sub, n_total - i) for i in range(0, n_total, sub)]
    out = [None] * len(sizes)

    def work(i):
        out[i] = synth.gen(MSG, sizes[i], seed=seed0 + i, vtap_mod=world if world > 1 else 0, vtap_rem=rank)

    it = iter(range(len(sizes)))
    lock = threading.Lock()

    def loop():
        while True:
            with lock:
                i = next(it, None)
            if i is None:
                return
            work(i)

    ts = [threading.Thread(target=loop) for _ in range(min(threads, len(sizes)))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.p, self.lines = index, None, []

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
        sm, mx, reasons = [], [], set()
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_near_gpu(index):
    """Run this rank's host threads (and, by first touch, place its pinned staging buffers) on the CPUs
    local to its GPU's PCIe root, like a NUMA-aware ingester deployment would.  Returns what it did."""
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


def ncu_traffic(kernel, recs_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed
    `ncu --set full` capture (profiles/r01_traffic.json names the capture), scaled to this launch's record
    count.  None when no capture is recorded: ncu cannot run inside the bench."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        t = json.load(open(p))[kernel]
        b = (t["dram_bytes_read"] + t["dram_bytes_write"]) * recs_per_launch / t["records_per_launch"]
        return b, t["source"]
    except Exception:  # noqa: BLE001
        return None, None


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def output_bytes(batch):
    """Exact size of the packed l7_flow_log column block (what the EMIT kernel writes)."""
    cols, nrows = batch.columns(abi.TABLE_L7_FLOW_LOG)
    total = strings = 0
    for c in cols:
        t = c.ch_type
        if t in (abi.CH_STR, abi.CH_LC_STR_ROWS):
            total += (0 if c.native else nrows * 16) + c.buf.nbytes
            strings += c.buf.nbytes
        elif t in (abi.CH_ARR_STR, abi.CH_ARR_LC_STR):
            total += nrows * 8 + (0 if c.native else c.pos.nbytes) + c.buf.nbytes
            strings += c.buf.nbytes
        elif t in (abi.CH_ARR_UINT16, abi.CH_ARR_FLOAT64):
            total += nrows * 8 + c.values.nbytes
            strings += c.values.nbytes
        else:
            total += nrows * c.elem_size + (nrows if c.nulls is not None else 0)
    return total, nrows, strings


def cpu_port_run(synth, sample_records, threads, seed, warm=True):
    """The CPU oracle (C++ restatement of decoder.Run ... AppendToColumnBlock) on `threads`
    independent decoder instances, each with its own dictionary view and caches."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle
    oras, wls = [], []
    ocfg = abi.Cfg()
    ocfg.flow_tag_cache_max_size = 1 << 20
    for t in range(threads):
        o = Oracle(ocfg)
        o.load_synth(synth)
        oras.append(o)
        wls.append(synth.gen(MSG, sample_records, seed=seed + t))
    if warm:
        # steady state, like a long-running ingester: column blocks and caches already sized
        ts = [threading.Thread(target=o.submit, args=(w,)) for o, w in zip(oras, wls)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    t0 = time.perf_counter()
    ts = [threading.Thread(target=o.submit, args=(w,)) for o, w in zip(oras, wls)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    dt = time.perf_counter() - t0
    n = sum(w.nrecords for w in wls)
    for o in oras:
        o.close()
    return n / dt, n, dt


def main():
    a = parse_args()
    rank, world, local = dist_env()
    if world != a.gpus and world > 1:
        a.gpus = world
    from deepflow_b200 import build
    from deepflow_b200.synth import Synth
    if rank == 0:
        build.build_synth()
    synth = None
    cfg_common = {"workload": "C2: L7 HTTP AppProtoLogsData, SmartEncoding on, flow_tag on (BASELINE.json configs[1])",
                  "records_per_gpu_per_step": a.records, "sub_batch_records": a.sub_batch,
                  "dictionary_pods": a.pods, "sharding": "vtap_id % n_gpus" if a.gpus > 1 else "none",
                  "l2": "inputs 4+ GB per step >> 126 MB L2 (no flush needed)"}

    # ------------------------------------------------------------------ reference arm --
    if a.impl == "reference":
        if rank != 0:
            return 0
        build.build_oracle()
        synth = Synth(n_pods=a.pods)
        cores = os.cpu_count() or 1
        threads = min(cores, 64)
        sample = 50000
        for _ in range(a.warmup):
            cpu_port_run(synth, 2000, threads, 900)
        t_all, n_all = 0.0, 0
        for s in range(a.steps):
            v, n, dt = cpu_port_run(synth, sample, threads, 1000 + s * 100)
            t_all += dt
            n_all += n
        value = n_all / t_all
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "records/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000 * t_all / max(1, a.steps),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32/u64 integer",
                "data": "synthetic", "config": cfg_common,
                "cpu_baseline": {"value": value, "unit": "records/s", "cores": threads, "kind": "port",
                                 "sample": f"{threads} decoder instances x {sample} L7 records per step "
                                           "(C++ restatement of the Go path; the Go ingester cannot be built here)"},
                "e2e": {"value": value, "unit": "records/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------------ GPU arm --
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: libdfgpu has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local)
    numa = bind_near_gpu(local)  # before any pinned allocation: first touch places the staging buffers
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from deepflow_b200.native import Engine
    synth = Synth(n_pods=a.pods)
    t0 = time.time()
    subs = gen_sub_batches(synth, a.records, a.sub_batch, 1000 + 7919 * rank, world, rank)
    nrec_step = sum(w.nrecords for w in subs)
    in_bytes = sum(w.nbytes for w in subs)
    gen_s = time.time() - t0

    # pinned host copies of every sub-batch (the Go shim's staging buffers)
    pinned = []
    for w in subs:
        t = torch.empty(max(1, w.nbytes), dtype=torch.uint8, pin_memory=True)
        t[:w.nbytes].copy_(torch.from_numpy(w.data))
        pinned.append(t)
    # device-resident copies for the `value` leg
    dev_data = [p.to(f"cuda:{local}", non_blocking=True) for p in pinned]
    torch.cuda.synchronize()

    cfg = abi.Cfg()
    cfg.max_batch_bytes = max(w.nbytes for w in subs) + (1 << 20)
    cfg.flow_tag_cache_max_size = 1 << 20  # tag set of this workload stays below it: no LRU eviction either side
    engines = [Engine(local, cfg) for _ in range(max(1, a.decoders))]
    for e in engines:
        e.load_synth(synth)
    main_eng = engines[0]
    stream = torch.cuda.ExternalStream(main_eng.stream, device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, kernels only --------------------------------------
    # The decoders of this GPU (DecoderQueueCount contexts, the reference default is 2) each take
    # every len(engines)-th sub-batch on their own stream, like the e2e leg and like the ingester's
    # decoder goroutines; the framing chase of one overlaps the passes of the other.
    def resident_step():
        err = []

        def worker(k):
            try:
                eng = engines[k]
                for i in range(k, len(subs), len(engines)):
                    w = subs[i]
                    eng.stage_ptr(MSG, dev_data[i].data_ptr(), w.nbytes, w.frames.ctypes.data, len(w.frames))
                    eng.run()
            except Exception as ex:  # noqa: BLE001
                err.append(ex)

        ts = [threading.Thread(target=worker, args=(k,)) for k in range(len(engines))]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if err:
            raise err[0]

    streams = [torch.cuda.ExternalStream(e.stream, device=local) for e in engines]
    tstream = torch.cuda.Stream(device=local)  # timing stream: start event before, end event after every decoder
    # clocks / throttle reasons: sampled from the first warm-up step to the end of the e2e leg (the
    # value leg alone lasts ~0.1 s, less than nvidia-smi needs to start on an 8-GPU box)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(a.warmup):
        resident_step()
    for e in engines:
        e.sync()
        e.profile_enable(True)
    launches0 = sum(e.launch_count for e in engines)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(tstream)
    for st in streams:
        st.wait_event(ev0)
    for _ in range(a.steps):
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
    emit_ms, emit_n, emit_in = 0.0, 0, 0
    stage_ms = {}
    for e in engines:
        st_ms, _n = e.profile_stages()
        for k, v in st_ms.items():
            stage_ms[k] = stage_ms.get(k, 0.0) + v
        ms, n, nb = e.profile_read()
        emit_ms, emit_n, emit_in = emit_ms + ms, emit_n + n, emit_in + nb
        e.profile_enable(False)
    launches_value = sum(e.launch_count for e in engines) - launches0
    # validate the last resident run + get exact output sizes for the roofline
    last_i = len(subs) - 1
    b = engines[last_i % len(engines)].fetch()
    cnt = b.counters()
    assert cnt.error_count == 0 and cnt.count == subs[last_i].nrecords, (cnt.error_count, cnt.count)
    out_b_last, rows_last, str_b_last = output_bytes(b)
    b.release()

    # ---- roofline leg: the dominant kernel timed alone -------------------------------------------------
    # In the value leg the decoders' kernels share the SMs, so a kernel's event time there contains its
    # neighbours.  One more pass of one decoder by itself: CUDA events of the library, on its stream,
    # right around every EMIT launch.
    main_eng.sync()
    main_eng.profile_enable(True)
    for i, w in enumerate(subs):
        main_eng.stage_ptr(MSG, dev_data[i].data_ptr(), w.nbytes, w.frames.ctypes.data, len(w.frames))
        main_eng.run()
    main_eng.sync()
    solo_stage_ms, solo_runs = main_eng.profile_stages()
    solo_emit_ms, solo_emit_n, _ = main_eng.profile_read()
    main_eng.profile_enable(False)
    main_eng.fetch().release()

    # ---- e2e: host buffers through the C ABI, H2D + kernels + D2H per sub-batch -------------
    d2h_total = [0]
    from deepflow_b200.native import Batch, lib
    from deepflow_b200.sharded import FLOW_TAG_TABLES, TagMerge
    merge = TagMerge(device=torch.device("cuda", local)) if world > 1 else None

    def e2e_step(engines=engines):
        err = []
        new_fps, new_times = [], []

        def worker(k):
            eng = engines[k]
            try:
                for i in range(k, len(subs), len(engines)):
                    w = subs[i]
                    out = C.c_void_p()
                    rc = lib().dfgpu_submit(eng._h, MSG, pinned[i].data_ptr(), w.nbytes, w.frames.ctypes.data,
                                            len(w.frames), None, None, C.byref(out))
                    if rc != 0:
                        raise RuntimeError(f"dfgpu_submit rc={rc}")
                    if merge is not None:  # keys this decoder discovered: published after the step
                        b = Batch(out)
                        for t in FLOW_TAG_TABLES:
                            f = b.tag_fingerprints(t)
                            if len(f):
                                cols, _ = b.columns(t)
                                new_fps.append(f)
                                new_times.append(cols[0].values.astype(np.uint32))
                        b.release()
                    else:
                        lib().dfgpu_batch_release(out)
            except Exception as ex:  # noqa: BLE001
                err.append(ex)

        ts = [threading.Thread(target=worker, args=(k,)) for k in range(len(engines))]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if err:
            raise err[0]
        if merge is not None:
            # the one collective of the path: all-gather of the newly emitted flow_tag keys (NCCL),
            # then every decoder of this rank imports the union
            fps = np.concatenate(new_fps) if new_fps else np.zeros((0, 2), np.uint64)
            tms = np.concatenate(new_times) if new_times else np.zeros(0, np.uint32)
            _keep, rf, rt = merge.exchange(fps, tms)
            allf, allt = np.concatenate([rf, fps]), np.concatenate([rt, tms])
            for e in engines:
                e.tags_import(allf, allt)

    for _ in range(a.warmup):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    for e in engines:
        e.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()

    # ---- e2e again with dfgpu_cfg.native_strings (first step): String / Array(String)
    # columns cross PCIe in ClickHouse native encoding, their Position arrays stay on the GPU ------------
    e2e_nat_s = nat_bytes_per_row = None
    if not a.no_native_leg and world == 1:  # N=1 only: the scaling runs keep to the two legs above
        for e in engines:
            e.sync()
        ncfg = abi.Cfg()
        ncfg.max_batch_bytes = cfg.max_batch_bytes
        ncfg.flow_tag_cache_max_size = cfg.flow_tag_cache_max_size
        ncfg.native_strings = 1
        nat_engines = [Engine(local, ncfg) for _ in range(len(engines))]
        for e in nat_engines:
            e.load_synth(synth)
        for _ in range(a.warmup):
            e2e_step(nat_engines)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step(nat_engines)
        for e in nat_engines:
            e.sync()
        barrier()
        e2e_nat_s = time.perf_counter() - t0
        w = subs[-1]
        nb = nat_engines[0].submit(w)
        nc = nb.counters()
        assert nc.error_count == 0 and nc.count == w.nrecords, (nc.error_count, nc.count)
        nat_out_b, nat_rows, _ = output_bytes(nb)
        nat_bytes_per_row = nat_out_b / max(1, nat_rows)
        nb.release()
        for e in nat_engines:
            e.close()

    # per-step byte accounting for e2e (H2D = staging buffers + frame tables; D2H = column blocks)
    bytes_per_row = out_b_last / max(1, rows_last)
    d2h_per_step = int(bytes_per_row * nrec_step)
    h2d_per_step = int(in_bytes + sum(w.frames.nbytes for w in subs))

    # ---- max over ranks ----------------------------------------------------------------------
    times = torch.tensor([dev_ms, e2e_s * 1000.0, (e2e_nat_s or 0.0) * 1000.0], dtype=torch.float64,
                         device=f"cuda:{local}")
    recs = torch.tensor([float(nrec_step)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        dist.all_reduce(recs, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max, e2e_nat_ms_max = times.tolist()
    total_recs_step = recs.item()

    if rank == 0:
        value = total_recs_step * a.steps / (dev_ms_max / 1000.0)
        e2e_value = total_recs_step * a.steps / (e2e_ms_max / 1000.0)
        peak, peak_src = measured_peak_hbm()
        # Dominant kernel = the EMIT/pack pass.  Its algorithmic bytes per record: every byte of the packed
        # row it writes (all of the block except the 41 SmartEncoding columns the KG kernel writes: 116 B)
        # + the string / array bytes it copies out of the wire record.  The slot buffer it also reads is an
        # artefact of this design, not compulsory traffic, and is NOT counted.
        str_per_row = str_b_last / max(1, rows_last)
        alg_per_rec = (bytes_per_row - 116) + str_per_row
        emit_launch_s = solo_emit_ms / 1000.0 / max(1, solo_emit_n)
        recs_per_launch = nrec_step / max(1, solo_emit_n)
        achieved = alg_per_rec * recs_per_launch / emit_launch_s / 1e9 if emit_launch_s > 0 else 0.0
        # the whole per-record path (SURVEY §8d B_L7: wire + 4 + packed row + 32 B x 12 probes) over the
        # three kernels that carry it (SIZE/decode, EMIT, SmartEncoding)
        path_bytes = in_bytes / nrec_step + bytes_per_row + 32 * 12
        path_ms = solo_stage_ms.get("size_pass", 0) + solo_stage_ms.get("emit_pass", 0)
        path_achieved = path_bytes * nrec_step / (path_ms / 1000.0) / 1e9 if path_ms else 0.0
        traffic, traffic_src = ncu_traffic("k_l7_pass<EMIT>", recs_per_launch)
        roof = {"bound": "hbm", "kernel": "k_l7_pass<EMIT>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_record": alg_per_rec, "launches_timed": solo_emit_n,
                "avg_launch_ms": solo_emit_ms / max(1, solo_emit_n),
                "timed": "one decoder alone (kernels of concurrent decoders would overlap the events)",
                "share_of_device_time": solo_emit_ms / sum(solo_stage_ms.values()) if solo_stage_ms else None,
                "stage_ms_per_step_one_decoder": solo_stage_ms,
                "path": {"kernels": "k_l7_pass<SIZE> + k_l7_pass<EMIT> + k_l7_kg",
                         "algorithmic_bytes_per_record": path_bytes, "ms_per_step": path_ms,
                         "achieved": path_achieved, "frac": path_achieved / peak},
                "stage_ms_per_step_all_decoders": {k: v / a.steps for k, v in stage_ms.items()}}
        line = {"metric": METRIC, "value": value, "unit": "records/s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": dev_ms_max / a.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32/u64 integer (f64 casts only at pack)",
                "data": "synthetic", "config": dict(cfg_common, mean_wire_bytes=in_bytes / nrec_step,
                                                    packed_bytes_per_row=bytes_per_row, gen_seconds=gen_s,
                                                    decoders_per_gpu=len(engines), host_affinity=numa,
                                                    tag_merge=(None if merge is None else {
                                                        "collective": "all_gather of emitted flow_tag key fingerprints, "
                                                                      "once per e2e step (NCCL)",
                                                        "bytes_sent_rank0": merge.bytes_sent,
                                                        "seconds_rank0": merge.seconds})),
                "e2e": {"value": e2e_value, "unit": "records/s", "h2d_bytes_per_step": h2d_per_step,
                        "d2h_bytes_per_step": d2h_per_step, "ms_per_step": e2e_ms_max / a.steps},
                "gpu_launches": int(launches_value), "clocks": clocks, "roofline": roof}
        if e2e_nat_s:
            line["e2e_native_strings"] = {
                "value": total_recs_step * a.steps / (e2e_nat_ms_max / 1000.0), "unit": "records/s",
                "h2d_bytes_per_step": h2d_per_step, "d2h_bytes_per_step": int(nat_bytes_per_row * nrec_step),
                "ms_per_step": e2e_nat_ms_max / a.steps,
                "what": "same call, dfgpu_cfg.native_strings=1: String / Array(String) columns leave the GPU as "
                        "ClickHouse native column data (uvarint length + bytes), no Position arrays over PCIe"}
        if not a.no_cpu_baseline:
            build.build_oracle()
            v, n, dt = cpu_port_run(synth, 400_000, 1, 4242)
            line["cpu_baseline"] = {"value": v, "unit": "records/s", "cores": 1, "kind": "port",
                                    "sample": f"{n} L7 records of the same generator, one decoder instance, {dt:.1f} s"}
        print(json.dumps(line))
    for e in engines:
        e.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
