#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 300 python tools/prof_l7.py --records 625000 > $O/stage_breakdown_msg5.txt 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none --launch-skip 57 -c 19 -f -o $O/ncu_l7 python tools/prof_l7.py --runs 1 > $O/ncu_l7.log 2>&1; echo "ncu l7 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_l7_subbatch.csv python tools/prof_l7.py --runs 2 > $O/launches_l7.log 2>&1; echo "launch list rc=$?"
bash tools/gpu_bench.sh
