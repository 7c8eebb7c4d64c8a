"""CPU: the JSON line of `bench.py --impl reference` (the reference's OpenCV path timed on the host cores) carries
the keys the driver reads.  Tiny run: 2 streams, 1 step."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--streams", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == pytest.approx(d["value"])
    ss = cb["single_stream"]          # BASELINE configs[0]: one stream, per-frame latency on the host (cv2 backend only)
    if cb["kind"] == "reference":
        assert ss["ms_per_frame_1_thread"] > 0 and ss["ms_per_frame_all_threads"] > 0 and ss["opencv_threads"] >= 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0 and e["value"] == pytest.approx(d["value"])
