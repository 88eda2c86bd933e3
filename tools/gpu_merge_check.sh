#!/bin/bash
# N=2: headline + the C5 leg only (the tag merge over NCCL)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 --legs c2,c5 --no-cpu-baseline > gpurun_out/bench_merge.json 2> gpurun_out/bench_merge.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_merge.json').read().strip().splitlines()[-1])
print('value',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1))
c=d['legs']['c5_highcard']; print('c5', c['value']/1e6, c['e2e']['value']/1e6, c['merge'])
PY
