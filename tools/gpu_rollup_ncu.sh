#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_ru_key|k_ru_fold" -c 2 -f -o gpurun_out/r02/ncu_rollup python bench.py --steps 1 --warmup 0 --legs c3 --no-cpu-baseline > gpurun_out/r02/ncu_rollup.log 2>&1; echo "rc=$?"
