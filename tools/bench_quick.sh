timeout 300 python bench.py --records 1250000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_small.log 2>&1; echo rc=$?; python -c "
import json
l=open('gpurun_out/bench_small.log').read().strip().splitlines()[-1]
d=json.loads(l); print('value',round(d['value']/1e6,2),'e2e',round(d['e2e']['value']/1e6,2),'roof',round(d['roofline']['frac'],4),'emit_ms',round(d['roofline']['avg_launch_ms'],3),'share',round(d['roofline']['share_of_device_time'],3),'ms/step', round(d['ms_per_step'],3))
"
