#!/usr/bin/env python
"""Where a kernel's instructions go: reads `ncu -i X.ncu-rep --page source --csv --print-source sass`
output and prints contiguous SASS regions with their share of executed warp instructions, average
active lanes and stall samples.  Usage: ncu_sass_profile.py sass.csv [nregions]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    H = rows[hdr]
    ix = {k: H.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
    ins = []
    for r in rows[hdr + 1:]:
        if len(r) < len(H):
            continue
        ins.append((r[ix["Source"]].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0),
                    int(r[ix["Thread Instructions Executed"]] or 0)))
    tot_i = sum(i[2] for i in ins) or 1
    tot_s = sum(i[1] for i in ins) or 1
    print(f"instructions {len(ins)}  executed {tot_i}  samples {tot_s}  lanes {sum(i[3] for i in ins) / tot_i:.1f}")
    # regions: split where the execution count changes by > 25 %
    regs = []
    start = 0
    for k in range(1, len(ins) + 1):
        if k == len(ins) or abs(ins[k][2] - ins[k - 1][2]) > 0.25 * max(ins[k][2], ins[k - 1][2], 1):
            regs.append((start, k))
            start = k
    out = []
    for a, b in regs:
        e = sum(i[2] for i in ins[a:b])
        s = sum(i[1] for i in ins[a:b])
        t = sum(i[3] for i in ins[a:b])
        out.append((e, s, a, b, t))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    for e, s, a, b, t in sorted(out, reverse=True)[:n]:
        print(f"[{a:5d},{b:5d}) n={b - a:4d} exec/inst={e // (b - a):9d} share={100 * e / tot_i:5.1f}% samples={100 * s / tot_s:5.1f}% lanes={t / max(e, 1):4.1f}  {ins[a][0][:60]}")


if __name__ == "__main__":
    main()
