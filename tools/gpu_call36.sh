#!/bin/bash
run() { echo "== overlap=$1 kg=$2 str=$3"; DFGPU_KG_OVERLAP=$1 DFGPU_KG_CTAS=$2 DFGPU_STR_CTAS=$3 timeout 300 python tools/prof_l7.py --records 625000 2>&1 | grep -E "pack  |smartencoding|total"; }
run 0 0 0
run 1 0 0
run 1 2 0
run 1 2 4
run 1 3 4
run 1 2 3
run 1 1 5
DFGPU_KG_OVERLAP=1 DFGPU_KG_CTAS=2 DFGPU_STR_CTAS=4 timeout 600 python -m pytest tests/test_l7_gpu.py tests/test_l7_native_gpu.py -x -q 2>&1 | tail -2
