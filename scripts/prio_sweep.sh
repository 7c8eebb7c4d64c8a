# stream priorities of the pipelined step (pre,harris,select,lk,poisson); last sweep of round 2 (final kernels)
B="python bench.py --no-cpu-baseline --no-e2e --steps 200 --warmup 10 --profile-steps 0 --no-other-configs --no-chained"
for rep in 1 2; do
for P in "-1,0,0,0,-2" "-1,0,0,0,0" "-2,0,0,0,-3" "-2,-1,-1,-1,-3" "-1,0,0,0,-5" "-3,-1,-1,-1,-5"; do
  RDFE_PRIO="$P" timeout 120 $B 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prio $P', round(d['value']), round(d['ms_per_step'],4))"
done; done
