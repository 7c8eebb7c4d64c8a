"""Synthetic multi-stream workload for benchmarks: a ring of `ring` frames per camera stream,
cached on local disk so that every process of one box (GPU ranks, CPU-baseline workers, the
reference arm) reads identical pixels.  SURVEY.md 8(d): 752x480 EuRoC-shaped streams,
seeds default_rng(stream_id*1000+7).
"""
from __future__ import annotations

import multiprocessing as mp
import os

import numpy as np

from .synthetic import SyntheticStream

CACHE_DIR = os.environ.get("RDVIO_SYNTH_CACHE", "/tmp/rdvio_synth_cache")

# BASELINE.json configs; "max_level" is OpenCV's maxLevel => max_level+1 images
WORKLOADS = {
    "euroc": dict(width=752, height=480, points=150, max_level=3, win=21),
    "advio": dict(width=1280, height=720, points=300, max_level=4, win=21),
    "hd": dict(width=1920, height=1080, points=1000, max_level=5, win=31),
}


def algorithmic_bytes(width, height, points, max_level, win, undistort=False):
    """SURVEY.md 8(d) stage-materialised byte model per frame. Returns (total, per-stage dict).
    undistort=True adds the ingest stage of 8(f) rank 1 (raw frame read + undistorted frame written = 2S; the 6 B/px
    fixed-point map is one per context and shared by every frame of a launch, so it is not charged per frame)."""
    sizes = []
    w, h = width, height
    for l in range(max_level + 1):
        sizes.append(w * h)
        nw, nh = (w + 1) // 2, (h + 1) // 2
        if nw <= win or nh <= win:
            break
        w, h = nw, nh
    S, P, SL = sizes[0], sum(sizes), sizes[-1]
    U = sum(min(points * (win + 1) ** 2, s) for s in sizes)
    stages = {
        "clahe_hist_lut": S,
        "clahe_apply": 2 * S,
        "pyrdown": (P - SL) + (P - S),
        "scharr": 5 * P,
        "harris_nms": S,
        "lk_track": 12 * U + 64 * points,
        "select": 0,
        "poisson_append": 0,
    }
    if undistort:
        stages["undistort"] = 2 * S
    return sum(stages.values()), stages


def _cache_path(stream_id, width, height, ring):
    return os.path.join(CACHE_DIR, f"s{stream_id}_{width}x{height}_T{ring}.npy")


def _make_one(args):
    stream_id, width, height, ring = args
    path = _cache_path(stream_id, width, height, ring)
    if os.path.exists(path):
        return path
    st = SyntheticStream(stream_id, width, height, period=ring)
    arr = np.stack([st.frame(k) for k in range(ring)], 0)
    tmp = f"{path}.{os.getpid()}.tmp.npy"
    np.save(tmp, arr)
    os.replace(tmp, path)
    return path


def ensure_rings(stream_ids, width, height, ring, workers=None):
    """Generate (in parallel, once) and cache the frame rings of the given streams."""
    os.makedirs(CACHE_DIR, exist_ok=True)
    todo = [(s, width, height, ring) for s in stream_ids if not os.path.exists(_cache_path(s, width, height, ring))]
    if todo:
        workers = workers or max(1, min(len(todo), (os.cpu_count() or 1)))
        if workers == 1:
            for t in todo:
                _make_one(t)
        else:
            with mp.get_context("spawn").Pool(workers) as pool:
                pool.map(_make_one, todo, chunksize=1)
    return [_cache_path(s, width, height, ring) for s in stream_ids]


def load_ring(stream_id, width, height, ring, mmap=True):
    path = _cache_path(stream_id, width, height, ring)
    if not os.path.exists(path):
        _make_one((stream_id, width, height, ring))
    return np.load(path, mmap_mode="r" if mmap else None)
