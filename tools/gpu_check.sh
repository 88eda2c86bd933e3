#!/bin/bash
# one gpurun call: parity tests, then stage breakdowns of the three message types
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
for m in 5 4 3; do
  timeout 300 python tools/prof_l7.py --msg $m --records 625000 > gpurun_out/prof_msg$m.log 2>&1; echo "rc=$?" >> gpurun_out/prof_msg$m.log
  cat gpurun_out/prof_msg$m.log
done
