"""Shared parity harness: run the same workload through the CUDA library (C ABI)
and the CPU oracle, compare every column of every table bit for bit."""
from deepflow_b200 import abi


def compare_tables(gpu_batch, oracle, tables):
    problems = []
    for t in tables:
        gcols, gn = gpu_batch.columns(t)
        ocols, on = oracle.columns(t)
        if gn != on:
            problems.append(f"table {t}: rows {gn} (gpu) != {on} (oracle)")
            continue
        if len(gcols) != len(ocols):
            problems.append(f"table {t}: ncols {len(gcols)} != {len(ocols)}")
            continue
        for g, o in zip(gcols, ocols):
            msg = abi.columns_equal(g, o)
            if msg:
                problems.append(f"table {t}: {msg}")
    return problems


def compare_counters(gc, oc, keys=("raw_count", "count", "error_count", "drop_count", "l7_http_count",
                                   "l7_dns_count", "l7_sql_count", "l7_nosql_count", "l7_rpc_count",
                                   "l7_mq_count")):
    problems = []
    for k in keys:
        if getattr(gc, k) != getattr(oc, k):
            problems.append(f"counter {k}: {getattr(gc, k)} (gpu) != {getattr(oc, k)} (oracle)")
    return problems
