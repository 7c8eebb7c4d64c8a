#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest: $(tail -n 1 gpurun_out/r2_pytest7.log)"
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc $?"; tail -n 4 gpurun_out/r2_bench7.err
