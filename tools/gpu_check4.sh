#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python tools/prof_l7.py --msg 5 --records 625000 > gpurun_out/prof_base.log 2>&1; echo "== base"; grep -E "decode|pack |smart|flow_tag|total" gpurun_out/prof_base.log
for v in kg2 kg3 dec4 dec6; do
  DFGPU_LIB=abl/libdfgpu_$v.so timeout 300 python tools/prof_l7.py --msg 5 --records 625000 > gpurun_out/prof_$v.log 2>&1; echo "== $v"; grep -E "decode|pack |smart|flow_tag|total" gpurun_out/prof_$v.log
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_l7_decode|k_l7_strings" --launch-skip 6 -c 3 -o gpurun_out/l7_r02c python tools/prof_l7.py --runs 1 > gpurun_out/ncu_l7.log 2>&1; echo "ncu rc=$?"
