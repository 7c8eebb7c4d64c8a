#!/bin/bash
# GPU box, one GPU: A/B of the Harris strip geometry (narrow last tile on/off, strip height) -- tests first.
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest: $(tail -n 1 gpurun_out/pytest.log)"
RDFE_HARRIS_ROWS=120 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_cv2.py -m gpu -x -q -k "harris or detect or cv2" > gpurun_out/pytest_rows120.log 2>&1; echo "pytest rows=120: $(tail -n 1 gpurun_out/pytest_rows120.log)"
B='python bench.py --no-cpu-baseline --no-e2e'
RDFE_HARRIS_NARROW=0 $B > gpurun_out/ab_narrow0.json 2> gpurun_out/ab_narrow0.err; echo "narrow0 rc $?"
$B > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err; echo "default rc $?"
RDFE_HARRIS_ROWS=120 $B > gpurun_out/ab_rows120.json 2> gpurun_out/ab_rows120.err; echo "rows120 rc $?"
RDFE_HARRIS_ROWS=160 $B > gpurun_out/ab_rows160.json 2> gpurun_out/ab_rows160.err; echo "rows160 rc $?"
for f in narrow0 default rows120 rows160; do python - "$f" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
print(sys.argv[1], round(d["value"]), "frames/s; harris us", round(d["kernels"]["harris_nms"]["us_per_launch"], 1), "lk us", round(d["kernels"]["lk_track"]["us_per_launch"], 1))
PY
done
