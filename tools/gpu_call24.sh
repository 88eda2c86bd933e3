#!/bin/bash
O=gpurun_out/r02
mkdir -p $O
timeout 1200 ncu --set full --import-source on --clock-control none --launch-skip 57 -c 19 -f -o $O/ncu_l7 python tools/prof_l7.py --runs 1 > $O/ncu_l7.log 2>&1; echo "ncu l7 rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"; tail -c 600 $O/bench_reference_arm.json
