"""CPU arm of the benchmark (TEST/BASELINE INFRASTRUCTURE ONLY): the reference's per-frame call
sequence FeatureTracker::run -> preprocess(new), track_keypoints(prev -> new), release(prev),
detect_keypoints(new)  (/root/reference/src/rdvio/src/feature_tracker.cpp:32-98) executed on host
cores for a set of independent camera streams.

backend "cv2"  : oracle.cv2_reference.Cv2Image -- the reference's own OpenCV calls (kind "reference")
backend "port" : oracle.fe_oracle (C restatement)                                   (kind "port")

Each worker process owns a fixed subset of streams (state = previous preprocessed image), runs
OpenCV single-threaded, and is driven step by step over a pipe so that the parent can time whole
batches ("one step = one new frame for every stream").
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def pick_backend():
    from .cv2_reference import HAVE_CV2
    return "cv2" if HAVE_CV2 else "port"


class _StreamState:
    def __init__(self, sid, wl, ring, backend):
        from rd_vio_b200.synthetic import SyntheticStream
        from rd_vio_b200.workload import load_ring
        self.sid, self.wl, self.ring, self.backend = sid, wl, ring, backend
        self.frames = load_ring(sid, wl["width"], wl["height"], ring)
        self.synth = SyntheticStream(sid, wl["width"], wl["height"], period=ring)
        self.kp, self.pred = [], []
        for k in range(ring):                      # untimed: the carried keypoints of every ring frame
            im = self._pre(self.frames[k])
            kp = self._detect(im, np.zeros((0, 2)))[:wl["points"]]
            self.kp.append(kp)
            self.pred.append(self.synth.predict(k, kp))
        self.prev = self._pre(self.frames[0])
        self.last = None

    # -- backend dispatch
    def _pre(self, frame):
        if self.backend == "cv2":
            from .cv2_reference import Cv2Image
            im = Cv2Image(np.asarray(frame), level_num=self.wl["max_level"])
            im.WIN = self.wl["win"]
            im.preprocess(6.0, 8, 8)
            return im
        from . import fe_oracle as orc
        pre = orc.clahe(np.asarray(frame), 6.0, 8, 8)
        return (pre, orc.Pyramid(pre, self.wl["win"], self.wl["max_level"]))

    def _detect(self, im, existing):
        if self.backend == "cv2":
            return im.detect_keypoints(existing, self.wl["points"], 20.0)
        from . import fe_oracle as orc
        return orc.detect_keypoints(im[0], existing, self.wl["points"], 20.0)[0]

    def _track(self, a, b, curr, pred):
        if self.backend == "cv2":
            return a.track_keypoints(b, curr, pred)
        from . import fe_oracle as orc
        nxt, st, _ = orc.track_keypoints(a[1], b[1], curr, pred, self.wl["win"], self.wl["max_level"])
        return nxt, st

    def step(self, t):
        k = t % self.ring
        new = self._pre(self.frames[(k + 1) % self.ring])
        nxt, st = self._track(self.prev, new, self.kp[k], self.pred[k])
        out = self._detect(new, nxt[np.asarray(st) != 0])     # frame.cpp:160-170: only tracked points are carried
        self.prev = new
        self.last = (nxt, st, out)


def _worker(conn, stream_ids, wl, ring, backend):
    try:
        if backend == "cv2":
            import cv2
            cv2.setNumThreads(1)
        states = [_StreamState(s, wl, ring, backend) for s in stream_ids]
        conn.send(("ready", len(states)))
        while True:
            msg = conn.recv()
            if msg[0] == "stop":
                break
            if msg[0] == "step":
                t0 = time.perf_counter()
                for st in states:
                    st.step(msg[1])
                conn.send(("done", time.perf_counter() - t0))
            elif msg[0] == "result":
                st = states[msg[1]]
                conn.send(("result", st.last))
    except Exception as e:   # pragma: no cover
        conn.send(("error", repr(e)))


class CpuFrontEndPool:
    """`workers` processes, each owning a share of `stream_ids`."""

    def __init__(self, stream_ids, wl, ring, backend=None, workers=None):
        self.backend = backend or pick_backend()
        self.workers = max(1, min(workers or (os.cpu_count() or 1), len(stream_ids)))
        ctx = mp.get_context("spawn")
        self.conns, self.procs, self.shares = [], [], []
        for w in range(self.workers):
            share = list(stream_ids[w::self.workers])
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(b, share, wl, ring, self.backend), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
            self.shares.append(share)
        self.n_streams = len(stream_ids)
        for c in self.conns:
            m = c.recv()
            if m[0] != "ready":
                raise RuntimeError(f"cpu worker failed: {m}")

    def step(self, t):
        """One new frame for every stream; returns wall seconds."""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(("step", t))
        for c in self.conns:
            m = c.recv()
            if m[0] != "done":
                raise RuntimeError(f"cpu worker failed: {m}")
        return time.perf_counter() - t0

    def result(self, stream_index):
        w, j = stream_index % self.workers, stream_index // self.workers
        self.conns[w].send(("result", j))
        return self.conns[w].recv()[1]

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop",))
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()


def single_stream_latency(wl, ring, frames=60, stream_id=0):
    """BASELINE configs[0]: ONE stream through the reference's call sequence on the host, per-frame latency with one
    OpenCV thread and with OpenCV's default thread pool (cv2 backend only).  Median over `frames` frames."""
    from .cv2_reference import HAVE_CV2
    if not HAVE_CV2:
        return None
    import cv2
    st = _StreamState(stream_id, wl, ring, "cv2")
    out = {}
    for label, nthreads in (("ms_per_frame_1_thread", 1), ("ms_per_frame_all_threads", -1)):
        cv2.setNumThreads(nthreads)
        if nthreads < 0:
            out["opencv_threads"] = int(cv2.getNumThreads())
        ts = []
        for t in range(frames + 3):
            t0 = time.perf_counter()
            st.step(t)
            ts.append(time.perf_counter() - t0)
        out[label] = 1e3 * float(np.median(ts[3:]))
    cv2.setNumThreads(-1)
    return out

