#!/bin/bash
# bench.py under torchrun on N GPUs of the box (the driver's launch line), both arms
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
    print('value',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1), 'n', d['n_gpus'], 'tag_merge', d['config'].get('tag_merge'))
    for k,v in d['legs'].items():
        print(k, {kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','ok','mismatches','checks','groups','error')}, 'e2e', v.get('e2e',{}).get('value'))
        if 'merge' in v: print('    merge', v['merge'], 'keys', v.get('flow_tag_keys_emitted_per_step_all_ranks'), 'unrec', v.get('flow_tag_unrecorded'))
except Exception as e:
    print('parse failed', e)
PY
