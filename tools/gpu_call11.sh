#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "metrics or smoke or host or sharded" > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_m.log
tail -15 gpurun_out/pytest_m.log
timeout 300 python tools/prof_l7.py --msg 3 --records 625000 2>&1 | tee gpurun_out/prof_msg3.log
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_l7_kg|k_l7_flow_tag|k_l7_ft|k_l7_decode" --launch-skip 18 -c 8 -f -o gpurun_out/l7_r02d python tools/prof_l7.py --runs 1 > gpurun_out/ncu_l7d.log 2>&1; echo "ncu l7 rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_m_decode|k_m_emit" --launch-skip 6 -c 2 -f -o gpurun_out/m_r02a python tools/prof_l7.py --msg 3 --runs 1 > gpurun_out/ncu_m.log 2>&1; echo "ncu m rc=$?"
