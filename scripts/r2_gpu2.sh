#!/bin/bash
# round 2: all GPU tests with the default Harris kernel, the Harris/detect tests with the two alternatives, A/B benches, instruction counts
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_impl3.log 2>&1; echo "pytest impl3 rc=$?" | tee -a gpurun_out/r2_pytest_impl3.log
tail -3 gpurun_out/r2_pytest_impl3.log
for impl in 0 1; do
RDFE_HARRIS_IMPL=$impl python -m pytest tests/test_gpu_harris_prefilter.py tests/test_gpu_parity.py -m gpu -x -q -k "harris or detect or candidates or degenerate" > gpurun_out/r2_pytest_impl$impl.log 2>&1; echo "pytest impl$impl rc=$?" | tee -a gpurun_out/r2_pytest_impl$impl.log
tail -2 gpurun_out/r2_pytest_impl$impl.log
done
for impl in 3 1 0; do
RDFE_HARRIS_IMPL=$impl RDFE_HARRIS_MB=3 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/r2_b_impl$impl.json 2> gpurun_out/r2_b_impl$impl.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_b_impl$impl.json"))
    print("impl$impl", round(d["value"]), "f/s", {k: round(v["us_per_launch"],1) for k,v in d["kernels"].items() if v["us_per_launch"]>0})
except Exception as e: print("impl$impl failed", e)
PY
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/r2_plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.per_cycle_active -k regex:harris --clock-control none -c 6 --csv --log-file gpurun_out/r2_ncu_harris3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/r2_ncu2.log 2>&1
grep -E "duration|inst_executed|issue_active" gpurun_out/r2_ncu_harris3.csv | tail -6 | cut -d, -f5,13- | cut -c1-200
