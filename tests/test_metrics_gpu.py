"""GPU parity: flow_metrics path (SURVEY A10, A5 twin, A8 metrics blocks, A9) through the C
ABI against the CPU oracle — every column of the 9 flow_metrics tables, the unmarshaller's
app_service tag rows and the decoder counters, bit exact, across consecutive batches; then
the reduce-by-key operator on every table."""
import pytest

from deepflow_b200 import abi
from deepflow_b200 import synth as S

pytestmark = pytest.mark.gpu

MSG = abi.DFGPU_MSG_METRICS
TABLES = [abi.TABLE_METRICS_BASE + i for i in range(9)] + [abi.TABLE_METRICS_APP_SERVICE]
COUNTERS = ("raw_count", "count", "error_count", "drop_count")


def fresh(synth, cfg=None):
    from deepflow_b200.native import Engine
    from oracle_binding import Oracle
    if cfg is None:
        cfg = abi.Cfg()
    cfg.rollup_tables = 0x1ff  # keep raw meters of every table beside its columns: the reduce folds them
    eng = Engine(0, cfg)
    eng.load_synth(synth)
    ora = Oracle(cfg)
    ora.load_synth(synth)
    return eng, ora


def run_both(eng, ora, wl, reduce_tables=()):
    from parity import compare_counters, compare_tables
    ora.submit(wl)
    b = eng.submit(wl)
    problems = compare_tables(b, ora, TABLES)
    problems += compare_counters(b.counters(), ora.counters(), keys=COUNTERS)
    gc, oc = b.counters(), ora.counters()
    for t in TABLES:
        if gc.rows[t] != oc.rows[t]:
            problems.append(f"rows[{t}] {gc.rows[t]} != {oc.rows[t]}")
    b.release()
    for t in reduce_tables:
        for bucket in (60, 1):
            assert ora.metrics_reduce(t, bucket) == 0
            rb = eng.metrics_reduce(t, bucket)
            problems += [f"reduce(table {t}, {bucket}s): {p}"
                         for p in compare_tables(rb, ora, [abi.TABLE_METRICS_REDUCED])]
            rb.release()
    return problems


@pytest.mark.parametrize("sizes,seed,flags", [
    ([1], 1, 0), ([31, 33], 2, 0), ([5000, 5000, 5000], 3, 0), ([80000], 4, 0),
    ([20000, 20000], 5, S.F_ALL_WIRE), ([20000, 3000], 6, S.F_ALL_WIRE | S.F_CORRUPT),
])
def test_metrics_parity(synth, built, sizes, seed, flags):
    eng, ora = fresh(synth)
    try:
        for k, n in enumerate(sizes):
            wl = synth.gen(MSG, n, seed=seed * 100 + k, flags=flags)
            problems = run_both(eng, ora, wl)
            assert not problems, f"batch {k}:\n" + "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()


def test_metrics_reduce_every_table(synth, built):
    eng, ora = fresh(synth)
    try:
        wl = synth.gen(MSG, 60000, seed=77)
        problems = run_both(eng, ora, wl, reduce_tables=TABLES[:9])
        assert not problems, "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()


def test_metrics_reduce_heavy_duplication(synth, built):
    """Many rows per group: a small endpoint pool makes the 5-tuple keys repeat (the 1s -> 1m
    rollup of SURVEY C3 in miniature); sums, maxes, first/last rules all matter."""
    import numpy as np
    eng, ora = fresh(synth)
    try:
        wl = synth.gen(MSG, 50000, seed=78)
        # replay the same documents three times: every group gets >= 3 rows
        data = np.concatenate([wl.data] * 3)
        frames = np.concatenate([wl.frames] * 3).copy()
        for r in range(3):
            frames["offset"][r * len(wl.frames):(r + 1) * len(wl.frames)] += r * wl.data.nbytes
        big = S.Workload(data, frames, wl.nrecords * 3, MSG)
        problems = run_both(eng, ora, big, reduce_tables=[abi.TABLE_METRICS_BASE + 6, abi.TABLE_METRICS_BASE + 3,
                                                           abi.TABLE_METRICS_BASE + 4])
        assert not problems, "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()


def test_metrics_empty_and_tiny_frames(synth, built):
    import numpy as np
    eng, ora = fresh(synth)
    try:
        fr = np.zeros(2, abi.FRAME_DTYPE)
        fr["org_id"], fr["team_id"], fr["vtap_id"] = 1, 1, 1
        empty = S.Workload(np.zeros(0, np.uint8), fr, 0, MSG)
        problems = run_both(eng, ora, empty)
        assert not problems, "\n".join(problems)
        wl = synth.gen(MSG, 300, seed=9, max_frame_bytes=600)  # one or two documents per frame
        problems = run_both(eng, ora, wl, reduce_tables=[abi.TABLE_METRICS_BASE + 1])
        assert not problems, "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()


def test_metrics_app_service_expiry_across_batches(synth, built):
    """The unmarshaller's app_service tag cache (lru, 1800 s): same documents again inside the
    window emit nothing; the generator's +3600 s documents re-emit inside one batch."""
    eng, ora = fresh(synth)
    try:
        wl = synth.gen(MSG, 30000, seed=31)
        assert not run_both(eng, ora, wl)
        ora.submit(wl)
        b = eng.submit(wl)
        _, n = b.columns(abi.TABLE_METRICS_APP_SERVICE)
        _, on = ora.columns(abi.TABLE_METRICS_APP_SERVICE)
        b.release()
        assert n == on
    finally:
        eng.close()
        ora.close()


def test_rollup_across_submits(synth, built):
    """The 1s -> 1m rollup proper: the accumulator persists across submits.  Three consecutive batches whose
    groups overlap (the same documents replayed with others in between) are folded one by one; the flush
    must equal the oracle's accumulation — groups in first-appearance order across ALL batches, sums / maxes
    over every batch, OooTx/OooRx from the LAST batch's last row, the *_fail fields from the FIRST row."""
    import numpy as np
    from parity import compare_tables
    eng, ora = fresh(synth)
    try:
        wls = [synth.gen(MSG, 30000, seed=300), synth.gen(MSG, 20000, seed=301), synth.gen(MSG, 30000, seed=300),
               synth.gen(MSG, 25000, seed=302)]
        tables = [abi.TABLE_METRICS_BASE + 6, abi.TABLE_METRICS_BASE + 1, abi.TABLE_METRICS_BASE + 3,
                  abi.TABLE_METRICS_BASE + 8, abi.TABLE_METRICS_BASE + 4]
        for wl in wls:
            ora.submit(wl)
            eng.submit(wl).release()
            for t in tables:
                assert ora.rollup_add(t, 60) == 0
                eng.rollup_add(t, 60)
        problems = []
        for t in tables:
            assert ora.rollup_flush(t) == 0
            rb = eng.rollup_flush(t)
            cols, n = rb.columns(abi.TABLE_METRICS_REDUCED)
            assert n > 0
            problems += [f"rollup(table {t}): {p}" for p in compare_tables(rb, ora, [abi.TABLE_METRICS_REDUCED])]
            rb.release()
        assert not problems, "\n".join(problems[:20])
        # the accumulator started over: one more batch reduces like a fresh one
        ora.submit(wls[1])
        eng.submit(wls[1]).release()
        t = tables[0]
        assert ora.metrics_reduce(t, 60) == 0
        rb = eng.metrics_reduce(t, 60)
        problems = compare_tables(rb, ora, [abi.TABLE_METRICS_REDUCED])
        rb.release()
        assert not problems, "\n".join(problems[:20])
    finally:
        eng.close()
        ora.close()
