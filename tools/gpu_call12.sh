#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_l7.py --msg 3 --records 625000 2>&1 | tee gpurun_out/prof_msg3.log
timeout 900 python bench.py --steps 3 --warmup 3 --legs c2,native,lz4 --no-cpu-baseline > gpurun_out/bench_lz4.json 2> gpurun_out/bench_lz4.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_lz4.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_lz4.json').read().strip().splitlines()[-1])
    print('value',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1))
    print('native', d.get('e2e_native_strings'))
    print('lz4', d.get('e2e_lz4'))
except Exception as e:
    print('parse failed', e)
PY
