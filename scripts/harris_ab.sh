# Harris strip height / implementation in the pipelined step (round 2, after the footprint and priority changes)
B="python bench.py --no-cpu-baseline --no-e2e --steps 200 --warmup 10 --no-other-configs --no-chained"
for rep in 1 2; do
for R in 80 40 48 60 96 120 160; do
  RDFE_HARRIS_ROWS=$R timeout 120 $B 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rows $R', round(d['value']), 'harris us', round(d['kernels']['harris_nms']['us_per_launch'],1))"
done
RDFE_HARRIS_IMPL=1 timeout 120 $B 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('impl 1', round(d['value']), 'harris us', round(d['kernels']['harris_nms']['us_per_launch'],1), round(d['kernels'].get('harris_resolve',{}).get('us_per_launch',0),1))"
RDFE_HARRIS_IMPL=3 timeout 120 $B 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('impl 3', round(d['value']), 'harris us', round(d['kernels']['harris_nms']['us_per_launch'],1))"
done
