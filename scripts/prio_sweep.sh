# priority sweep of the pipelined step (run on a GPU box)
B="python bench.py --no-cpu-baseline --no-e2e --steps 200 --warmup 10 --profile-steps 0"
i=0
for M in 0 -1 -5; do
for P in "-1,0,0,0,0" "-2,0,-3,-1,-3" "-2,-1,-3,0,-3" "0,0,0,0,0" "-2,0,-3,0,-3"; do
  BENCH_MAIN_PRIO=$M RDFE_PRIO="$P" $B > gpurun_out/q$i.json 2>gpurun_out/q$i.err
  python -c "
import json; d=json.loads(open('gpurun_out/q$i.json').read().strip().splitlines()[-1]); print('main $M prio $P', round(d['value']), d['ms_per_step'])"
  i=$((i+1))
done
done
BENCH_MAIN_PRIO=-1 RDFE_PRIO="-2,0,-3,-1,-3" python bench.py --no-cpu-baseline --no-e2e --steps 100 --warmup 10 --profile-steps 0 --timeline gpurun_out/timeline4.json > /dev/null 2>&1
python -c "
import torch; print(torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream,'priority_range') else 'n/a')"
