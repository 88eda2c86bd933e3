"""Multi-GPU host logic: one process per GPU, records sharded by vtap_id, and the
only collective of the path — the merge of newly discovered flow_tag keys.

Per-record work needs no exchange: the dictionary is replicated, `vtap_id` is part of every
flow_metrics table code (server/libs/flow-metrics/tag.go:497-506) so reduced groups are
disjoint across shards, and the `_id` / trace-index counters are per decoder exactly like the
reference's per-goroutine state (flow_log.go:226-247).  What the reference does NOT share —
the flow_tag lru of each decoder (flow_tag_writer.go:65-77) — is merged here so that the
shards together emit what one global cache would:

  step k:  every rank decodes its shard and emits rows for keys new to ITS set
           all ranks all-gather (fingerprint, time) of the rows they emitted      [NCCL]
           ownership: the lowest rank that emitted a key in this step keeps the row,
                      higher ranks drop theirs
           every rank imports the other ranks' keys into its device set

Pure torch.distributed code: runs over NCCL with CUDA tensors (bench.py, one rank per GPU)
and over gloo with CPU tensors (tests/test_sharded_cpu.py, world_size 2).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_of(vtap_id, world_size):
    """gpu = hash(vtap_id) % G with hash = identity."""
    return int(vtap_id) % int(world_size)


def _void16(a):
    a = np.ascontiguousarray(a, np.uint64).reshape(-1, 2)
    return a.view(np.dtype((np.void, 16))).ravel()


def duplicates_of(mine, lower):
    """mine (n, >=2), lower (m, >=2) int64 tensors of 128-bit fingerprints (lo, hi) on one device -> bool (n,):
    mine[i] also occurs in lower.  `lower` is sorted lexicographically by (lo, hi); a binary search on lo gives the
    run of entries sharing the low half (the same key emitted by several decoders: long runs of equal pairs), a
    second, vectorised binary search on hi inside that run decides.  No per-element Python anywhere: a step of the
    high-cardinality leg merges tens of millions of keys."""
    n = len(mine)
    dup = torch.zeros(n, dtype=torch.bool, device=mine.device)
    if n == 0 or len(lower) == 0:
        return dup
    lo, hi = lower[:, 0].contiguous(), lower[:, 1].contiguous()
    order = torch.argsort(hi, stable=True)
    order = order[torch.argsort(lo[order], stable=True)]  # lexicographic (lo, hi)
    lo_s, hi_s = lo[order], hi[order]
    mlo, mhi = mine[:, 0].contiguous(), mine[:, 1].contiguous()
    a = torch.searchsorted(lo_s, mlo, right=False)
    b = torch.searchsorted(lo_s, mlo, right=True)
    last = len(lo_s) - 1
    steps = int(torch.log2((b - a).max().clamp(min=1).double()).ceil().item()) + 1
    for _ in range(steps):  # lower bound of mhi inside [a, b)
        active = a < b
        mid = (a + b) // 2
        right = active & (hi_s[mid.clamp(max=last)] < mhi)
        a = torch.where(right, mid + 1, a)
        b = torch.where(active & ~right, mid, b)
    end = torch.searchsorted(lo_s, mlo, right=True)
    return (a < end) & (lo_s[a.clamp(max=last)] == mlo) & (hi_s[a.clamp(max=last)] == mhi)


class TagMerge:
    """All-gather of (fingerprint, time) lists + the ownership rule."""

    def __init__(self, group=None, device=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device if device is not None else torch.device("cpu")
        self.bytes_sent = 0       # payload this rank contributed to all-gathers
        self.seconds = 0.0

    def exchange(self, fps, times):
        """fps (n, 2) uint64, times (n,) uint32 (numpy) of the rows this rank emitted in this step.
        Returns keep (n,) bool, remote_fps (m, 2) uint64, remote_times (m,) uint32."""
        import time as _t
        t0 = _t.perf_counter()
        fps = np.ascontiguousarray(fps, np.uint64).reshape(-1, 2)
        times = np.ascontiguousarray(times, np.uint32).reshape(-1)
        n = len(times)
        assert len(fps) == n
        if self.world == 1:
            return np.ones(n, bool), np.zeros((0, 2), np.uint64), np.zeros(0, np.uint32)
        if self.device.type == "cuda":
            keep, others, _mine = self.exchange_device(fps, times)
            o = others.cpu().numpy()
            self.seconds += _t.perf_counter() - t0
            return keep, np.ascontiguousarray(o[:, 0:2]).view(np.uint64), o[:, 2].astype(np.uint32)
        # 1) lengths
        cnt = torch.tensor([n], dtype=torch.int64, device=self.device)
        cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(cnts, cnt, group=self.group)
        counts = [int(c.item()) for c in cnts]
        m = max(counts)
        remote_fps = np.zeros((0, 2), np.uint64)
        remote_times = np.zeros(0, np.uint32)
        keep = np.ones(n, bool)
        if m:
            # 2) padded payload: (m, 3) int64 = lo, hi, time
            pay = np.zeros((m, 3), np.int64)
            pay[:n, 0:2] = fps.view(np.int64)
            pay[:n, 2] = times
            mine = torch.from_numpy(pay).to(self.device)
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            self.bytes_sent += n * 24
            got = [p.cpu().numpy() for p in parts]
            lower = [got[r][:counts[r], 0:2] for r in range(self.rank) if counts[r]]
            if lower and n:
                keep = ~np.isin(_void16(fps), _void16(np.concatenate(lower).view(np.uint64)))
            others = [got[r][:counts[r]] for r in range(self.world) if r != self.rank and counts[r]]
            if others:
                o = np.concatenate(others)
                remote_fps = np.ascontiguousarray(o[:, 0:2]).view(np.uint64)
                remote_times = o[:, 2].astype(np.uint32)
        self.seconds += _t.perf_counter() - t0
        return keep, remote_fps, remote_times


    def exchange_device(self, fps, times):
        """The same exchange with the keys kept on the GPU (NCCL all-gather, ownership by sort + binary search).
        Returns keep (n,) bool numpy, others (m, 3) int64 CUDA tensor (lo, hi, time) of the other ranks' keys,
        mine (n, 3) int64 CUDA tensor.  Ownership compares both halves of the fingerprint (duplicates_of)."""
        fps = np.ascontiguousarray(fps, np.uint64).reshape(-1, 2)
        times = np.ascontiguousarray(times, np.uint32).reshape(-1)
        n = len(times)
        cnt = torch.tensor([n], dtype=torch.int64, device=self.device)
        cnts = [torch.zeros_like(cnt) for _ in range(self.world)]
        dist.all_gather(cnts, cnt, group=self.group)
        counts = [int(c.item()) for c in cnts]
        m = max(counts)
        empty = torch.zeros((0, 3), dtype=torch.int64, device=self.device)
        if m == 0:
            return np.ones(n, bool), empty, empty
        pay = np.zeros((m, 3), np.int64)
        pay[:n, 0:2] = fps.view(np.int64)
        pay[:n, 2] = times
        mine_pad = torch.from_numpy(pay).to(self.device)
        parts = [torch.empty_like(mine_pad) for _ in range(self.world)]
        dist.all_gather(parts, mine_pad, group=self.group)
        self.bytes_sent += n * 24
        mine = mine_pad[:n]
        keep = np.ones(n, bool)
        lower = [parts[r][:counts[r]] for r in range(self.rank) if counts[r]]
        if lower and n:
            keep = (~duplicates_of(mine, torch.cat(lower))).cpu().numpy()
        others = [parts[r][:counts[r]] for r in range(self.world) if r != self.rank and counts[r]]
        return keep, (torch.cat(others) if others else empty), mine


FLOW_TAG_TABLES = (2, 3, 4)  # DFGPU_TABLE_L7_CUSTOM_FIELD, _VALUE, L7_APP_SERVICE


class ShardedDecoder:
    """One rank's decoder: an Engine on this rank's GPU + the tag merge after every L7 batch."""

    def __init__(self, engine, merge):
        self.engine = engine
        self.merge = merge

    def submit(self, wl):
        """Returns (batch, {table: keep mask}).  Rows with keep == False were emitted by a lower
        rank in the same step; the host shim drops them before the ckwriter hand-off."""
        b = self.engine.submit(wl)
        keeps = {}
        if wl.msg_type != 5:  # only the L7 path discovers flow tags
            return b, keeps
        fps, times, sizes = [], [], []
        for t in FLOW_TAG_TABLES:
            f = b.tag_fingerprints(t)
            cols, n = b.columns(t)
            tm = cols[0].values.astype(np.uint32) if n else np.zeros(0, np.uint32)  # the `time` column
            fps.append(f)
            times.append(tm)
            sizes.append(n)
        allf = np.concatenate(fps) if sum(sizes) else np.zeros((0, 2), np.uint64)
        allt = np.concatenate(times) if sum(sizes) else np.zeros(0, np.uint32)
        keep, rf, rt = self.merge.exchange(allf, allt)
        off = 0
        for t, n in zip(FLOW_TAG_TABLES, sizes):
            keeps[t] = keep[off:off + n]
            off += n
        self.engine.tags_import(rf, rt)
        return b, keeps
