#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_template_cache.py tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q > gpurun_out/r2_pytest11.log 2>&1; echo "pytest: $(tail -n 1 gpurun_out/r2_pytest11.log)"
timeout 150 python scripts/lk_cache_ab.py 20 > gpurun_out/r2_lkab2.log 2>&1; tail -n 3 gpurun_out/r2_lkab2.log
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; echo "bench rc $?"; tail -n 4 gpurun_out/r2_bench11.err
