#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 2
for rep in 1 2; do
  timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs --no-chained 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=b['kernels']; print('base', round(b['value']), 'lk', round(k['lk_track']['us_per_launch'],1), 'harris', round(k['harris_nms']['us_per_launch'],1))"
done
