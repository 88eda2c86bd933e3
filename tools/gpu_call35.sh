#!/bin/bash
timeout 900 python -m pytest tests/test_l7_gpu.py tests/test_l7_native_gpu.py tests/test_sharded_gpu.py tests/test_soak_gpu.py tests/test_host_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/prof_l7.py --records 625000 2>&1 | grep -E "flow_tag|total"
