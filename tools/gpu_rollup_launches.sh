#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_ru --csv --log-file gpurun_out/r02/launches_c3.csv python bench.py --steps 1 --warmup 1 --legs c3 --no-cpu-baseline > gpurun_out/r02/launches_c3.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02/launches_c3.csv')))
h=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
H=rows[h]; kn=H.index('Kernel Name'); mn=H.index('Metric Name'); mv=H.index('Metric Value'); mu=H.index('Metric Unit')
agg={}
for r in rows[h+1:]:
    if len(r)<=mv: continue
    n=r[kn].split('(')[0].replace('void ','').replace('unnamed>::','')
    v=float(r[mv].replace(',',''))
    agg.setdefault(n,{}).setdefault(r[mn]+' '+r[mu],[]).append(v)
for n,m in agg.items():
    print(n, {k:(len(v), round(sum(v)/len(v),3)) for k,v in m.items()})
PY
