#!/bin/bash
for rep in 1 2; do
for v in base hw2 lw2 hw2lw2; do
  if [ $v = base ]; then unset RDFE_LIB_PATH; else export RDFE_LIB_PATH=$PWD/rd_vio_b200/lib_variants/$v/librdvio_fe.so; fi
  timeout 200 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs --no-chained 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=b['kernels']; print('$v', round(b['value']), 'harris', round(k['harris_nms']['us_per_launch'],1), 'lk', round(k['lk_track']['us_per_launch'],1))"
done; done
for v in base l31c4; do
  if [ $v = base ]; then unset RDFE_LIB_PATH; else export RDFE_LIB_PATH=$PWD/rd_vio_b200/lib_variants/$v/librdvio_fe.so; fi
  timeout 200 python bench.py --workload hd --streams 32 --ring 4 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-other-configs --no-chained 2>/dev/null | python -c "import sys,json; b=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=b['kernels']; print('HD $v', round(b['value']), 'lk', round(k['lk_track']['us_per_launch'],1))"
done
