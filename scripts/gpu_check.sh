#!/bin/bash
# GPU box: run the GPU tests and one bench line; compact summary on stdout, full logs under gpurun_out/.
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1
echo "pytest: $(tail -n 1 gpurun_out/pytest.log)"
python bench.py --steps 200 --warmup 10 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print("value", round(d['value']), "ms/step", round(d['ms_per_step'], 4), "e2e", d['e2e'] and round(d['e2e']['value']),
      "cpu", d.get('cpu_baseline') and round(d['cpu_baseline']['value']), "clocks", d['clocks'])
if d.get('kernels'):
    print({k: round(v['us_per_launch'] * v['launches_per_step'], 1) for k, v in d['kernels'].items()})
print("roofline", d.get('roofline') and {k: d['roofline'][k] for k in ('kernel', 'achieved', 'frac', 'traffic')})
PY
