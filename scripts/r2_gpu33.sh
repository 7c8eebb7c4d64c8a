timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
C='python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 --no-other-configs --no-chained'
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum -k regex:lk_track --clock-control none --csv --log-file gpurun_out/r2_lknc.csv $C > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_lknc.csv')) if len(r)>5]
h=rows[0]; im,iv=h.index('Metric Name'),h.index('Metric Value')
t=[float(r[iv].replace(',','')) for r in rows[1:] if r[im]=='gpu__time_duration.sum']
i=[float(r[iv].replace(',','')) for r in rows[1:] if r[im]=='smsp__inst_executed.sum']
print('lk launches',len(t),'us %.1f'%(sum(t)/len(t)/1e3 if t[0]>1e3 else sum(t)/len(t)),'Minst %.2f'%(sum(i)/len(i)/1e6))
PY
for rep in 1 2; do timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), 'lk', round(d['kernels']['lk_track']['us_per_launch'],1), 'chained', {k:round(v['value']) for k,v in d['chained'].items() if isinstance(v,dict)}, 'advio', round(d['other_configs']['advio']['value']), 'hd', round(d['other_configs']['hd']['value']), d['other_configs']['hd']['kernels_us_per_launch']['lk_track'])"; done
