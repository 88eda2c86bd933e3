#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_lz4_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/prof_l7.py --lz4 --runs 3 2>&1 | grep -E "frame bytes|fetch" | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --legs c2,native,lz4 --no-cpu-baseline > gpurun_out/bench_lz4.json 2> gpurun_out/bench_lz4.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_lz4.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_lz4.json').read().strip().splitlines()[-1])
    print('value',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1))
    print('native', round(d['e2e_native_strings']['value']/1e6,1))
    z=d['e2e_lz4']; print('lz4', round(z['value']/1e6,1), z['frame_bytes_per_row'], z['ms_per_step'])
except Exception as e:
    print('parse failed', e)
PY
