#!/bin/bash
# L4 parity tests + stage breakdown + one ncu capture of the L4 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "l4 or L4 or abi or smoke or tagged" > gpurun_out/pytest_l4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_l4.log
tail -15 gpurun_out/pytest_l4.log
timeout 300 python tools/prof_l7.py --msg 4 --records 625000 2>&1 | tee gpurun_out/prof_msg4.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_l4 --launch-skip 12 -c 4 -f -o gpurun_out/l4_r02a python tools/prof_l7.py --msg 4 --runs 1 > gpurun_out/ncu_l4.log 2>&1; echo "ncu rc=$?"
