#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-other-configs --no-e2e > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc $?"; tail -n 4 gpurun_out/r2_bench9.err
python scripts/disagreement.py --frames 200 --hd-frames 60 --out gpurun_out/r2_disagreement.json > gpurun_out/r2_disagreement.log 2>&1; echo "disagreement rc $?"; tail -n 5 gpurun_out/r2_disagreement.log
python -m pytest tests/test_gpu_golden.py -x -q 2>&1 | tail -n 3
