#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_replay.py -x -q 2>&1 | tail -n 2
for rep in 1 2; do
for M in 0 1; do
  RDFE_SCHARR_SPLIT=$M timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-other-configs --no-chained 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split $M', round(b['value']), 'e2e', round(b['e2e']['value']))"
done; done
