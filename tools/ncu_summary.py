#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the few numbers DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py <report.ncu-rep> > profiles/<name>.md
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    kn = hdr.index("Kernel Name")
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n")
    for row in rows[2:]:
        print(f"## {row[kn]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for w, i in idx:
            print(f"| {w} | {row[i]} | {units[i]} |")
        try:
            rd = float(row[hdr.index('dram__bytes_read.sum')])
            wr = float(row[hdr.index('dram__bytes_write.sum')])
            u = units[hdr.index('dram__bytes_read.sum')]
            print(f"| dram traffic (read+write) | {rd + wr:.6f} | {u} |")
        except Exception:
            pass
        print()


if __name__ == "__main__":
    main()
