"""CPU, world_size 2 over gloo: the host logic of the N>1 path — vtap sharding, the
all-gather of emitted flow_tag keys, the ownership rule and the import list — checked against
ONE global cache: the CPU oracle fed both shards.  (The device side of the merge,
dfgpu_tags_import / dfgpu_batch_tag_fingerprints, is covered by tests/test_sharded_gpu.py.)"""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
WORLD = 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fp(row):
    d = hashlib.blake2b(repr(row).encode(), digest_size=16).digest()
    return np.frombuffer(d, np.uint64)


def _rows(cols, n):
    names = [c.name for c in cols]
    data = [c.rows() for c in cols]
    return [tuple((nm, data[j][i]) for j, nm in enumerate(names) if nm != "time") for i in range(n)], \
        [int(data[names.index("time")][i]) for i in range(n)]


def _worker(rank, port, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        from deepflow_b200 import abi
        from deepflow_b200.sharded import TagMerge, shard_of
        from deepflow_b200.synth import Synth
        from oracle_binding import Oracle

        merge = TagMerge()
        # ---- 1. pure list logic ------------------------------------------------------------------
        A = np.arange(10, dtype=np.uint64).reshape(5, 2) + 100          # only rank 0
        B = np.arange(8, dtype=np.uint64).reshape(4, 2) + 900           # only rank 1
        Cc = np.arange(6, dtype=np.uint64).reshape(3, 2) + 5000         # both
        mine = np.concatenate([A, Cc]) if rank == 0 else np.concatenate([Cc[::-1], B])
        times = np.arange(len(mine), dtype=np.uint32) + 1000 * (rank + 1)
        keep, rf, rt = merge.exchange(mine, times)
        if rank == 0:
            assert keep.all()
            assert sorted(map(tuple, rf.tolist())) == sorted(map(tuple, np.concatenate([Cc, B]).tolist()))
            assert set(rt.tolist()) == set(range(2000, 2007))
        else:
            assert keep.tolist() == [False, False, False, True, True, True, True]
            assert sorted(map(tuple, rf.tolist())) == sorted(map(tuple, np.concatenate([A, Cc]).tolist()))
        # ragged / empty sides
        keep, rf, rt = merge.exchange(np.zeros((0, 2), np.uint64) if rank == 0 else B, np.zeros(0, np.uint32) if rank == 0 else times[:4])
        assert len(keep) == (0 if rank == 0 else 4) and keep.all()
        assert len(rf) == (4 if rank == 0 else 0)
        keep, rf, rt = merge.exchange(np.zeros((0, 2), np.uint64), np.zeros(0, np.uint32))
        assert len(keep) == 0 and len(rf) == 0

        # ---- 2. against one global cache (the CPU oracle) --------------------------------------------
        synth = Synth(n_pods=2000)
        cfg = abi.Cfg()
        cfg.flow_tag_cache_max_size = 1 << 21
        local = Oracle(cfg)
        local.load_synth(synth)
        glob = Oracle(cfg) if rank == 0 else None
        if glob:
            glob.load_synth(synth)
        seen_remote = set()
        T = abi.TABLE_L7_CUSTOM_FIELD_VALUE
        for step in range(2):
            shards = [synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 2500, seed=50 + step, vtap_mod=WORLD, vtap_rem=r)
                      for r in range(WORLD)]
            for r, wl in enumerate(shards):
                assert all(shard_of(v, WORLD) == r for v in wl.frames["vtap_id"])
            wl = shards[rank]
            local.submit(wl)
            cols, n = local.columns(T)
            rows, tms = _rows(cols, n)
            fps = np.array([_fp(r) for r in rows], np.uint64).reshape(-1, 2)
            # what dfgpu_tags_import does on the device: keys another rank emitted earlier are "seen"
            not_seen = np.array([tuple(f) not in seen_remote for f in fps.tolist()], bool).reshape(-1)
            keep, rf, rt = merge.exchange(fps[not_seen], np.array(tms, np.uint32)[not_seen])
            for f in rf.tolist():
                seen_remote.add(tuple(f))
            kept = [r for r, ns in zip(rows, not_seen) if ns]
            kept = [r for r, k in zip(kept, keep) if k]
            gathered = [None] * WORLD
            dist.all_gather_object(gathered, kept)
            if rank == 0:
                # one decoder, one cache, both shards of the step in rank order
                want = []
                for swl in shards:
                    glob.submit(swl)
                    gcols, gn = glob.columns(T)
                    grows, _ = _rows(gcols, gn)
                    out_q.put(("global", step, len(grows)))
                    want.extend(grows)
                union = [r for part in gathered for r in part]
                assert len(union) == len(set(union)), "a key was emitted by two ranks in one step"
                assert set(union) == set(want), (len(union), len(want))
                assert len(union) == len(want)
        out_q.put(("ok", rank, merge.bytes_sent))
    finally:
        dist.destroy_process_group()


def test_tag_merge_world2_gloo(built):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    codes = [p.exitcode for p in procs]
    msgs = []
    while not q.empty():
        msgs.append(q.get())
    assert codes == [0, 0], (codes, msgs)
    oks = [m for m in msgs if m[0] == "ok"]
    assert len(oks) == WORLD and all(m[2] > 0 for m in oks)


def test_duplicates_of_matches_set_membership():
    """The ownership rule of the device-side merge (sharded.duplicates_of) against plain set membership, including
    fingerprints that share their low half."""
    import numpy as np
    import torch
    from deepflow_b200.sharded import duplicates_of
    rng = np.random.default_rng(3)
    lower = rng.integers(-2**62, 2**62, size=(5000, 3), dtype=np.int64)
    mine = rng.integers(-2**62, 2**62, size=(3000, 3), dtype=np.int64)
    mine[:1000, :2] = lower[rng.integers(0, 5000, 1000), :2]        # real duplicates
    mine[1000:1100, 0] = lower[:100, 0]                              # same low half, different high half
    lower[200:300, 0] = lower[100:200, 0]                            # runs of equal low halves inside `lower`
    mine[1100:1150, :2] = lower[250:300, :2]
    got = duplicates_of(torch.from_numpy(mine), torch.from_numpy(lower)).numpy()
    keys = {(int(a), int(b)) for a, b in lower[:, :2]}
    want = np.array([(int(a), int(b)) in keys for a, b in mine[:, :2]])
    assert (got == want).all()
    assert not duplicates_of(torch.from_numpy(mine), torch.zeros((0, 3), dtype=torch.int64)).any()
