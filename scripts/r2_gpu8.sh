#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest: $(tail -n 1 gpurun_out/r2_pytest8.log)"
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-other-configs > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc $?"; tail -n 4 gpurun_out/r2_bench8.err
RDFE_LK_CTAS=5 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e --no-chained > gpurun_out/r2_bench8_c5.json 2> gpurun_out/r2_bench8_c5.err; echo "bench c5 rc $?"
