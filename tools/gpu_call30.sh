#!/bin/bash
timeout 900 python -m pytest tests/test_metrics_gpu.py -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 --legs c3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
c=d['legs']['c3_rollup']; print('c3', c['value']/1e9, c['ms_per_step'], c['checks'], c['roofline']['frac'])"
bash tools/gpu_call29.sh 2>&1 | tail -12 | cut -c1-220
