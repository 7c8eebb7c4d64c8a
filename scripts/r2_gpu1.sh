#!/bin/bash
# round 2, GPU call 1: all GPU tests, A/B of the Harris paths, instruction counts of the Harris kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r2_b_pref4.json 2> gpurun_out/r2_b_pref4.err
RDFE_HARRIS_MB=3 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/r2_b_pref3.json 2> gpurun_out/r2_b_pref3.err
RDFE_HARRIS_EXACT=1 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/r2_b_exact.json 2> gpurun_out/r2_b_exact.err
for f in pref4 pref3 exact; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_b_$f.json"))
    print("$f", round(d["value"]), "f/s", {k: round(v["us_per_launch"],1) for k,v in d["kernels"].items()}, "e2e", d.get("e2e") and round(d["e2e"]["value"]))
except Exception as e: print("$f failed", e)
PY
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/r2_plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.per_cycle_active -k regex:harris --clock-control none -c 18 --csv --log-file gpurun_out/r2_ncu_harris.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-steps 0 > gpurun_out/r2_ncu1.log 2>&1
tail -15 gpurun_out/r2_ncu_harris.csv | cut -c1-300
