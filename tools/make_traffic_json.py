#!/usr/bin/env python
"""profiles/r02_traffic.json from `ncu --set full` reports: per kernel the DRAM bytes of one launch
(dram__bytes_read.sum / dram__bytes_write.sum) and the records that launch processed.  bench.py reports them as
`roofline.traffic`.  usage: make_traffic_json.py records_per_launch report.ncu-rep [report2 ...] > profiles/r02_traffic.json"""
import csv
import json
import re
import subprocess
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    recs = int(sys.argv[1])
    out = {}
    for rep in sys.argv[2:]:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        kn, r, w = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for row in rows[2:]:
            name = re.sub(r"\(.*$", "", row[kn]).replace("void ", "").replace("unnamed>::", "").strip()
            if name in out:
                continue  # first launch of a kernel in the first report that has it
            out[name] = {"records_per_launch": recs, "dram_bytes_read": int(float(row[r]) * UNIT[units[r]]),
                         "dram_bytes_write": int(float(row[w]) * UNIT[units[w]]),
                         "source": f"profiles/{rep.split('/')[-1].replace('.ncu-rep', '.md')} (ncu --set full --clock-control none)"}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
