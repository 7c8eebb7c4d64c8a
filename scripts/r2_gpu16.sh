#!/bin/bash
mkdir -p gpurun_out
C='python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 --no-other-configs --no-chained'
for v in base nounroll f2i dblcheck all; do
  export RDFE_LIB_PATH=$PWD/rd_vio_b200/lib_variants/$v/librdvio_fe.so
  timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum -k regex:lk_track --clock-control none --csv --log-file gpurun_out/r2_lkvar_$v.csv $C > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_lkvar_$v.csv')) if len(r)>5]
h=rows[0]; im,iv=h.index('Metric Name'),h.index('Metric Value')
t=[float(r[iv].replace(',','')) for r in rows[1:] if r[im]=='gpu__time_duration.sum']
i=[float(r[iv].replace(',','')) for r in rows[1:] if r[im]=='smsp__inst_executed.sum']
print('$v', 'launches',len(t),'us %.1f'%(sum(t)/len(t)/1e3 if t and t[0]>1e3 else sum(t)/len(t)),'Minst %.2f'%(sum(i)/len(i)/1e6))
PY
  timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs --no-chained 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   bench', round(b['value']), 'lk us', round(b['kernels']['lk_track']['us_per_launch'],1))"
done
