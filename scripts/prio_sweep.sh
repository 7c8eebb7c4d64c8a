B="python bench.py --no-cpu-baseline --no-e2e --steps 200 --warmup 10 --profile-steps 0"
i=0
for P in "-1,0,0,0,0" "-2,-1,-1,0,0" "-1,0,-1,0,-1" "-1,-1,0,0,0" "-2,0,0,-1,-1"; do
  RDFE_PRIO="$P" $B > gpurun_out/q$i.json 2>gpurun_out/q$i.err
  python -c "
import json; d=json.loads(open('gpurun_out/q$i.json').read().strip().splitlines()[-1]); print('prio $P', round(d['value']), d['ms_per_step'])"
  i=$((i+1))
done
