#!/bin/bash
mkdir -p gpurun_out
N=10000000 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:refine_rows|assign_tc5h|reduce_partials|finalize_kernel|tc5h_prep" -c 120 --csv --log-file gpurun_out/launches_r2b_c5_kpp.csv python bench/c5_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_r2b_c5_kpp.csv')))
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
agg={}
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    name=r[kn].split('(')[0][:50]; v=float(r[mv].replace(',','')); u=r[mu]
    v = v/1e3 if u=='ns' else v*1e3 if u=='ms' else v
    agg.setdefault(name,[]).append(v)
for n,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print('%-52s n=%3d avg %10.1f us  first %10.1f last %10.1f' % (n,len(v),sum(v)/len(v), v[0], v[-1]))
PY
