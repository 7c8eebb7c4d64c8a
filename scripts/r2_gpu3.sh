#!/bin/bash
mkdir -p gpurun_out
sel="tests/test_gpu_harris_prefilter.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_vs_cv2.py"
for cfg in "3 1" "3 0" "0 1" "1 1"; do set -- $cfg
RDFE_HARRIS_IMPL=$1 RDFE_HARRIS_F2D=$2 python -m pytest $sel -m gpu -q -k "harris or detect or candidates or degenerate or golden or cv2" > gpurun_out/r2_pytest_h$1$2.log 2>&1; echo "pytest impl$1 f2d$2 rc=$? $(tail -1 gpurun_out/r2_pytest_h$1$2.log)"
done
for cfg in "3 1" "3 0" "0 1" "1 1"; do set -- $cfg
RDFE_HARRIS_IMPL=$1 RDFE_HARRIS_F2D=$2 RDFE_HARRIS_MB=3 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/r2_b_h$1$2.json 2> gpurun_out/r2_b_h$1$2.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_b_h$1$2.json"))
    print("impl$1 f2d$2", round(d["value"]), "f/s", {k: round(v["us_per_launch"],1) for k,v in d["kernels"].items() if v["us_per_launch"]>0 and "harris" in k})
except Exception as e: print("impl$1 f2d$2 failed", e)
PY
done
