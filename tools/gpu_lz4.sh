#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_l7.py --lz4 --runs 4 2>&1 | tee gpurun_out/prof_lz4.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_lz4 --launch-skip 3 -c 3 -f -o gpurun_out/lz4_r02a python tools/prof_l7.py --lz4 --runs 1 > gpurun_out/ncu_lz4.log 2>&1; echo "ncu rc=$?"
