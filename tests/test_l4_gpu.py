"""GPU parity: L4 TaggedFlow path through the C ABI against the
CPU oracle — every one of the 154 l4_flow_log columns, bit exact."""
import pytest

from deepflow_b200 import abi
from deepflow_b200 import synth as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines(synth, built):
    from deepflow_b200.native import Engine
    from oracle_binding import Oracle
    eng = Engine(0)
    eng.load_synth(synth)
    ora = Oracle()
    ora.load_synth(synth)
    yield eng, ora
    eng.close()
    ora.close()


def run_both(eng, ora, wl, tables):
    from parity import compare_counters, compare_tables
    ora.id_counter.value = eng.id_counter.value
    ora.submit(wl)
    b = eng.submit(wl)
    problems = compare_tables(b, ora, tables) + compare_counters(b.counters(), ora.counters())
    if eng.id_counter.value != ora.id_counter.value:
        problems.append(f"id counter {eng.id_counter.value} != {ora.id_counter.value}")
    b.release()
    return problems


@pytest.mark.parametrize("n,seed,flags", [
    (1, 1, 0), (31, 2, 0), (33, 3, 0), (5000, 4, 0), (60000, 5, 0),
    (20000, 6, S.F_ALL_WIRE), (20000, 7, S.F_ALL_WIRE | S.F_CORRUPT), (3000, 8, S.F_CORRUPT),
])
def test_l4_parity(engines, synth, n, seed, flags):
    eng, ora = engines
    wl = synth.gen(abi.DFGPU_MSG_TAGGEDFLOW, n, seed=seed, flags=flags)
    problems = run_both(eng, ora, wl, [abi.TABLE_L4_FLOW_LOG])
    assert not problems, "\n".join(problems[:20])


def test_l4_empty_and_ragged(engines, synth):
    import numpy as np
    eng, ora = engines
    wl = synth.gen(abi.DFGPU_MSG_TAGGEDFLOW, 100, seed=9, max_frame_bytes=2000)
    # add empty frames and a frame holding a zero-length record
    fr = wl.frames.copy()
    extra = np.zeros(2, abi.FRAME_DTYPE)
    extra["org_id"] = 1
    extra["team_id"] = 1
    extra["vtap_id"] = 1
    data = np.concatenate([wl.data, np.zeros(4, np.uint8)])
    extra[1]["offset"] = len(wl.data)
    extra[1]["len"] = 4  # one record of length 0: decodes to an all-default flow
    wl2 = S.Workload(data, np.concatenate([fr[:1], extra, fr[1:]]), None, wl.msg_type)
    problems = run_both(eng, ora, wl2, [abi.TABLE_L4_FLOW_LOG])
    assert not problems, "\n".join(problems[:20])
    # no frames at all
    wl3 = S.Workload(np.zeros(0, np.uint8), np.zeros(0, abi.FRAME_DTYPE), 0, wl.msg_type)
    problems = run_both(eng, ora, wl3, [abi.TABLE_L4_FLOW_LOG])
    assert not problems, "\n".join(problems[:20])
