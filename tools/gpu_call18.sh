#!/bin/bash
mkdir -p gpurun_out
for d in 4 6; do
  timeout 600 python bench.py --steps 3 --warmup 3 --legs c2 --no-cpu-baseline --decoders $d > gpurun_out/bench_d$d.json 2> gpurun_out/bench_d$d.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_d$d.json').read().strip().splitlines()[-1])
print('decoders',$d,'value',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1), d['e2e'].get('h2d_gbs_per_rank'), d['e2e'].get('d2h_gbs_per_rank'))
PY
done
