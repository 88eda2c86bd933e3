"""GPU: the device side of the multi-GPU tag merge — per-row lru-key fingerprints
(dfgpu_batch_tag_fingerprints) and dfgpu_tags_import — on ONE GPU with two contexts standing
in for two ranks.  After ownership + import the two shards together must emit exactly the
flow_tag rows ONE global cache emits (the CPU oracle fed both shards)."""
import numpy as np
import pytest

from deepflow_b200 import abi

pytestmark = pytest.mark.gpu

TABLES = (abi.TABLE_L7_CUSTOM_FIELD, abi.TABLE_L7_CUSTOM_FIELD_VALUE, abi.TABLE_L7_APP_SERVICE)


def rows_of(batch_or_oracle, table):
    cols, n = batch_or_oracle.columns(table)
    names = [c.name for c in cols]
    data = [c.rows() for c in cols]
    rows = [tuple((nm, data[j][i]) for j, nm in enumerate(names) if nm != "time") for i in range(n)]
    times = np.array([data[names.index("time")][i] for i in range(n)], np.uint32)
    return rows, times


def void16(a):
    return np.ascontiguousarray(a, np.uint64).reshape(-1, 2).view(np.dtype((np.void, 16))).ravel()


def test_two_contexts_equal_one_global_cache(synth, built):
    from deepflow_b200.native import Engine
    from oracle_binding import Oracle
    cfg = abi.Cfg()
    cfg.flow_tag_cache_max_size = 1 << 21
    ranks = [Engine(0, cfg), Engine(0, cfg)]
    glob = Oracle(cfg)
    glob.load_synth(synth)
    for e in ranks:
        e.load_synth(synth)
    try:
        for step in range(4):
            shards = [synth.gen(abi.DFGPU_MSG_PROTOCOLLOG, 6000, seed=70 + step, vtap_mod=2, vtap_rem=r) for r in range(2)]
            batches = [ranks[r].submit(shards[r]) for r in range(2)]
            emitted = []  # per rank: {table: (rows, fps, times)}
            for b in batches:
                per = {}
                for t in TABLES:
                    rows, times = rows_of(b, t)
                    fps = b.tag_fingerprints(t)
                    assert fps.shape == (len(rows), 2)
                    # a fingerprint identifies the lru key: equal rows <=> equal fingerprints inside a batch
                    assert len(set(map(tuple, fps.tolist()))) == len(set(rows)) == len(rows)
                    per[t] = (rows, fps, times)
                emitted.append(per)
            union = {t: [] for t in TABLES}
            for t in TABLES:
                rows0, fps0, _ = emitted[0][t]
                rows1, fps1, _ = emitted[1][t]
                keep1 = ~np.isin(void16(fps1), void16(fps0)) if len(fps1) and len(fps0) else np.ones(len(rows1), bool)
                union[t] = rows0 + [r for r, k in zip(rows1, keep1) if k]
            # import: each rank learns the other's keys
            for r in range(2):
                o = 1 - r
                fps = np.concatenate([emitted[o][t][1] for t in TABLES])
                tms = np.concatenate([emitted[o][t][2] for t in TABLES])
                ranks[r].tags_import(fps, tms)
            want = {t: [] for t in TABLES}
            for wl in shards:
                glob.submit(wl)
                for t in TABLES:
                    want[t] += rows_of(glob, t)[0]
            for t in TABLES:
                assert len(union[t]) == len(set(union[t])), (step, t)
                assert set(union[t]) == set(want[t]), (step, t, len(union[t]), len(want[t]))
            for b in batches:
                b.release()
    finally:
        for e in ranks:
            e.close()
        glob.close()
