"""The reference's own OpenCV call sequence, driven through Python cv2 -- TEST/BASELINE INFRASTRUCTURE ONLY.

rd_vio's hot path is 210 lines of glue (src/rdvio_extra/src/opencv_image.cpp) around five calls into
un-vendored OpenCV 4.x (CMakeLists.txt:30).  The C++ reference cannot be compiled in this image (no
OpenCV/Eigen/Ceres headers), but the image ships opencv-python-headless 4.13.0, which
exposes the very same functions.  `Cv2Image` below wraps them exactly as `OpenCvImage` does, so it
serves (a) as the ground truth the C oracle is pinned against (tests/test_oracle_vs_cv2.py,
tests/golden/make_golden.py) and (b) as the `--impl reference` CPU arm of bench.py.

One deviation forced by the binding: cv2.calcOpticalFlowPyrLK cannot take a pyramid list from Python
(SURVEY App. B7), so LK is given the level-0 images; OpenCV then rebuilds the same pyramid and
Scharr derivatives internally (identical results, slightly pessimistic timing).
"""
from __future__ import annotations

import numpy as np

try:
    import cv2
    HAVE_CV2 = True
except Exception:  # pragma: no cover - cv2 missing
    cv2 = None
    HAVE_CV2 = False

from . import fe_oracle as _orc   # PoissonDiskFilter restatement (host-side double math, not OpenCV)


class Cv2Image:
    """Mirror of rdvio::extra::OpenCvImage (opencv_image.h:9-56)."""

    WIN = 21

    def __init__(self, image, t=0.0, level_num=3):
        self.image = np.ascontiguousarray(image, np.uint8).copy()
        self.t = t
        self._level_num = level_num
        self.pyramid = None

    def level_num(self):
        return self._level_num

    def width(self):
        return self.image.shape[1]

    def height(self):
        return self.image.shape[0]

    # opencv_image.cpp:156-161
    def preprocess(self, clip_limit=6.0, width=8, height=8):
        self.image = cv2.createCLAHE(clipLimit=clip_limit, tileGridSize=(width, height)).apply(self.image)
        _, self.pyramid = cv2.buildOpticalFlowPyramid(self.image, (self.WIN, self.WIN), self._level_num, None, True)

    # opencv_image.cpp:38-73
    def detect_keypoints(self, keypoints, max_points=150, keypoint_distance=20.0):
        det = cv2.GFTTDetector_create(int(max_points), 1.0e-3, 20, 3, True)
        kps = det.detect(self.image)
        keypoints = np.asarray(keypoints, np.float64).reshape(-1, 2)
        if len(kps) == 0:
            return keypoints
        kps = sorted(kps, key=lambda k: -k.response)      # stable; ties keep GFTT order
        cand = np.array([[k.pt[0], k.pt[1]] for k in kps], np.float64)
        acc = _orc.poisson_filter(keypoints, cand, keypoint_distance)
        H, W = self.image.shape
        keep = ~((acc[:, 0] < 20) | (acc[:, 1] < 20) | (acc[:, 0] >= W - 20) | (acc[:, 1] >= H - 20))
        return np.concatenate([keypoints, acc[keep]], 0)

    # opencv_image.cpp:75-154
    def track_keypoints(self, next_image: "Cv2Image", curr_keypoints, next_keypoints=None):
        curr = np.asarray(curr_keypoints, np.float64).reshape(-1, 2)
        n = len(curr)
        c = curr.astype(np.float32)
        if next_keypoints is not None and len(next_keypoints) > 0:
            nxt = np.asarray(next_keypoints, np.float64).reshape(-1, 2).copy()
            q = nxt.astype(np.float32)
        else:
            nxt = np.zeros((n, 2), np.float64)
            q = c.copy()
        status = np.zeros(n, np.int8)
        if n == 0:
            return nxt, status
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
        H, W = self.image.shape
        q, st, _ = cv2.calcOpticalFlowPyrLK(self.image, next_image.image, c.reshape(-1, 1, 2), q.reshape(-1, 1, 2),
                                            winSize=(self.WIN, self.WIN), maxLevel=self._level_num, criteria=crit,
                                            flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        q = q.reshape(-1, 2)
        status[:] = st.reshape(-1)
        status[(q[:, 0] < 20) | (q[:, 0] >= W - 20) | (q[:, 1] < 20) | (q[:, 1] >= H - 20)] = 0
        d = (q - c).astype(np.float64)
        status[np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2) > (H // 4)] = 0
        r, rst, _ = cv2.calcOpticalFlowPyrLK(next_image.image, self.image, q.reshape(-1, 1, 2).copy(),
                                             c.reshape(-1, 1, 2).copy(), winSize=(self.WIN, self.WIN),
                                             maxLevel=self._level_num, criteria=crit,
                                             flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        r = r.reshape(-1, 2)
        e = (c - r).astype(np.float64)
        bad = (rst.reshape(-1) == 0) | (np.sqrt(e[:, 0] ** 2 + e[:, 1] ** 2) > 0.5)
        status[bad] = 0
        ok = status != 0
        nxt[ok] = q[ok].astype(np.float64)
        self.last_forward = q
        return nxt, status

    def release_image_buffer(self):
        self.image = None
        self.pyramid = None
