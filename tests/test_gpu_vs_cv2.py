"""GPU: the CUDA path directly against LIVE cv2 -- the reference's own OpenCV calls wrapped exactly like OpenCvImage
(oracle/cv2_reference.Cv2Image) -- on the same inputs, without the C restatement in between.  BASELINE configs 1-4
shapes: preprocess bit-exact, detect identical (plain float order, cv2.setUseOptimized(False)), tracked positions
within 0.01 px with status agreement >= 99.5 % (north_star tolerances)."""
import numpy as np
import pytest

from oracle.cv2_reference import HAVE_CV2, Cv2Image

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not HAVE_CV2, reason="cv2 not importable")]

CASES = [  # width, height, maxLevel, win, points
    (752, 480, 3, 21, 150),
    (1280, 720, 4, 21, 300),
    (1920, 1080, 5, 31, 1000),
]


@pytest.mark.parametrize("W,H,max_level,win,points", CASES)
def test_plugin_calls_against_cv2(W, H, max_level, win, points):
    import cv2
    from rd_vio_b200.frontend import FrontEnd
    from rd_vio_b200.synthetic import SyntheticStream
    st = SyntheticStream(11, W, H)
    f0, f1 = st.frame(0), st.frame(1)
    A, B = Cv2Image(f0, level_num=max_level), Cv2Image(f1, level_num=max_level)
    A.WIN = B.WIN = win
    A.preprocess(6.0, 8, 8)
    B.preprocess(6.0, 8, 8)
    with FrontEnd(W, H, max_level, win, num_slots=2, max_points=2 * points + 64) as fe:
        s0, s1 = fe.acquire(), fe.acquire()
        fe.preprocess([s0, s1], [f0, f1], 6.0, (8, 8))
        # preprocess: CLAHE output, every pyramid image and Scharr plane (opencv_image.cpp:156-161)
        assert fe.nlevels == len(A.pyramid) // 2
        for lvl in range(fe.nlevels):
            assert np.array_equal(fe.download_level(s0, lvl, 0), A.pyramid[2 * lvl]), f"image level {lvl}"
            assert np.array_equal(fe.download_level(s0, lvl, 1), A.pyramid[2 * lvl + 1]), f"Scharr level {lvl}"
        # detect_keypoints (opencv_image.cpp:38-73), plain float order
        cv2.setUseOptimized(False)
        try:
            ref_kp = A.detect_keypoints(np.zeros((0, 2)), points, 20.0)
        finally:
            cv2.setUseOptimized(True)
        kp = fe.detect([s0], [np.zeros((0, 2))], points, 20.0)[0]
        assert np.array_equal(kp, ref_kp), "detected keypoints differ from cv2"
        # track_keypoints (opencv_image.cpp:75-154) with the IMU-style prediction and without
        for pred in (st.predict(0, kp), None):
            r_next, r_st = A.track_keypoints(B, kp, pred)
            g_next, g_st = fe.track([s0], [s1], [kp], [pred] if pred is not None else None)
            g_next, g_st = g_next[0], g_st[0]
            # >= 99.5 % agreement; one flag of granularity for the small sets (live cv2's float32 LK sums depend on
            # the host's SIMD dispatch, so a point sitting on the 0.5-px round-trip gate may fall either way)
            mismatches = int((g_st != r_st).sum())
            assert mismatches <= max(1, int(0.005 * len(kp))), f"{mismatches} of {len(kp)} status flags differ"
            ok = (g_st != 0) & (r_st != 0)
            assert ok.sum() >= 0.5 * len(kp)
            assert np.abs(g_next[ok] - r_next[ok]).max() <= 0.01
        # second detect on top of the tracked points (what FeatureTracker::run does next)
        cv2.setUseOptimized(False)
        try:
            ref_kp2 = B.detect_keypoints(r_next[r_st != 0], points, 20.0)
        finally:
            cv2.setUseOptimized(True)
        # both sides get the SAME existing keypoints (cv2's survivors), so this holds whether or not a status flag differed
        kp2 = fe.detect([s1], [r_next[r_st != 0]], points, 20.0)[0]
        assert np.array_equal(kp2, ref_kp2)
