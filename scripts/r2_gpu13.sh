#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest: $(tail -n 1 gpurun_out/r2_pytest13.log)"
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-chained > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; echo "bench rc $?"; tail -n 4 gpurun_out/r2_bench13.err
timeout 120 python scripts/latency_breakdown.py 60 > gpurun_out/r2_latency_breakdown13.txt 2>&1; cat gpurun_out/r2_latency_breakdown13.txt
