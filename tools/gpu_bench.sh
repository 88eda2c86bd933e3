#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
    print('value',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,1),'one_dec',round(d['config'].get('one_decoder_value',0)/1e6,1))
    r=d['roofline']; print('dominant',r['stage'],round(r['frac'],3),'path',round(r['path']['frac'],3), {k:round(v,2) for k,v in r['stage_ms_per_step_one_decoder'].items()})
    print('e2e mode', d['e2e'].get('mode'), d['e2e'].get('d2h_bytes_per_row'), 'containers', d.get('e2e_containers',{}).get('value'), 'native', d.get('e2e_native_strings',{}).get('value'), 'lz4', d.get('e2e_lz4',{}).get('value'))
    for k,v in d['legs'].items():
        print(k, {kk:(round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','ok','mismatches','checks','groups','error')}, 'e2e', v.get('e2e',{}).get('value'))
        if 'per_type' in v:
            for t,x in v['per_type'].items(): print('   ',t, round(x['solo_records_per_s']/1e6,1) if x['solo_records_per_s'] else None, {a:round(b,2) for a,b in x['stage_ms'].items()})
        if 'roofline' in v: print('    roof', {a:(round(b,3) if isinstance(b,float) else b) for a,b in v['roofline'].items() if a in ('frac','achieved','multiple_of_bound')})
        if 'merge' in v: print('    merge', v['merge'], 'keys', v.get('flow_tag_keys_emitted_per_step_all_ranks'), 'unrec', v.get('flow_tag_unrecorded'))
    print('cpu', d.get('cpu_baseline',{}).get('value'), 'clocks', d.get('clocks'))
except Exception as e:
    print('parse failed', e)
PY
