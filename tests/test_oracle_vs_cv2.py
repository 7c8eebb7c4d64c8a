"""CPU: the oracle against LIVE cv2 (the reference's own OpenCV calls) on seeded inputs, when cv2 is importable."""
import numpy as np
import pytest

from conftest import random_image
from oracle import fe_oracle as orc
from oracle.cv2_reference import HAVE_CV2, Cv2Image

pytestmark = pytest.mark.skipif(not HAVE_CV2, reason="cv2 not importable")


@pytest.mark.parametrize("shape", [(480, 752), (478, 750), (480, 750), (477, 752), (135, 241), (720, 1280)])
def test_clahe_vs_cv2(shape):
    import cv2
    img = random_image(*shape, seed=shape[0])
    assert np.array_equal(cv2.createCLAHE(6.0, (8, 8)).apply(img), orc.clahe(img))


@pytest.mark.parametrize("shape,max_level,win", [((480, 752), 3, 21), ((479, 751), 3, 21), ((100, 130), 3, 21), ((540, 960), 4, 31)])
def test_pyramid_vs_cv2(shape, max_level, win):
    import cv2
    img = random_image(*shape, seed=7)
    n, pyr = cv2.buildOpticalFlowPyramid(img, (win, win), max_level, None, True)
    P = orc.Pyramid(img, win, max_level)
    assert P.nlevels == n + 1
    for l in range(P.nlevels):
        assert np.array_equal(pyr[2 * l], P.image(l)) and np.array_equal(pyr[2 * l + 1], P.deriv(l))


def test_harris_and_detect_vs_cv2(frames0):
    import cv2
    cv2.setUseOptimized(False)
    try:
        A = Cv2Image(frames0[0])
        A.preprocess()
        assert np.array_equal(cv2.cornerHarris(A.image, 3, 3, 0.04), orc.harris(A.image))
        for ex, r, k in ((np.zeros((0, 2)), 20.0, 150), (np.array([[100.2, 100.7], [400.0, 300.0]]), 10.0, 200)):
            assert np.array_equal(A.detect_keypoints(ex, k, r), orc.detect_keypoints(A.image, ex, k, r)[0])
    finally:
        cv2.setUseOptimized(True)


def test_track_vs_cv2(stream0, frames0):
    A, B = Cv2Image(frames0[0]), Cv2Image(frames0[1])
    A.preprocess(); B.preprocess()
    pts = orc.detect_keypoints(A.image, np.zeros((0, 2)), 150, 20.0)[0]
    pts = np.concatenate([pts, [[5., 5.], [751., 479.], [-30., 100.], [760., 300.], [375.5, 240.25]]], 0)
    pred = stream0.predict(0, pts)
    PA, PB = orc.Pyramid(A.image), orc.Pyramid(B.image)
    for p in (pred, None):
        n_cv, s_cv = A.track_keypoints(B, pts, p)
        n_or, s_or, _ = orc.track_keypoints(PA, PB, pts, p)
        assert (s_cv == s_or).mean() >= 0.995
        ok = (s_cv != 0) & (s_or != 0)
        assert ok.sum() > 100 and np.abs(n_cv[ok] - n_or[ok]).max() <= 0.01


def test_track_negative_fourth_weight_vs_cv2(stream0, frames0):
    """Sub-pixel offsets whose three rounded Q14 bilinear weights sum to 2^14 + 1: OpenCV keeps the fourth
    weight (-1) signed, and so must the restatement (and the CUDA kernel, tests/test_gpu_parity.py)."""
    hits = [(1, 8184), (2, 4093), (3, 2728), (5, 1638), (8, 1023), (13, 630), (21, 390), (30, 273), (45, 182),
            (63, 130), (88, 93), (90, 91), (105, 78), (117, 70)]
    A, B = Cv2Image(frames0[0]), Cv2Image(frames0[1])
    A.preprocess(); B.preprocess()
    corners = orc.detect_keypoints(A.image, np.zeros((0, 2)), 150, 20.0)[0]
    pts = np.array([[c[0] + hits[k % len(hits)][k % 2] * 2.0 ** -14, c[1] + hits[k % len(hits)][1 - k % 2] * 2.0 ** -14]
                    for k, c in enumerate(corners)])
    PA, PB = orc.Pyramid(A.image), orc.Pyramid(B.image)
    n_cv, s_cv = A.track_keypoints(B, pts, None)
    n_or, s_or, _ = orc.track_keypoints(PA, PB, pts, None)
    assert np.array_equal(s_cv != 0, s_or != 0)
    ok = s_cv != 0
    assert ok.sum() > 100 and np.abs(n_cv[ok] - n_or[ok]).max() <= 0.01


@pytest.mark.parametrize("shape", [(480, 752), (720, 1280), (257, 331)])
def test_undistort_vs_cv2(shape):
    import cv2
    H, W = shape
    img = random_image(H, W, seed=W)
    s = W / 752.0
    K = np.array([[458.654 * s, 0, 367.215 * s], [0, 457.296 * s, 248.375 * H / 480.0], [0, 0, 1]], np.float32)
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
    assert np.array_equal(cv2.undistort(img, K, D), orc.undistort(img, K, D))


def test_gray_conversion_and_color_undistort_vs_cv2():
    import cv2
    rng = np.random.default_rng(5)
    H, W = 240, 320
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    bgra = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    assert np.array_equal(cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY), orc.bgr2gray(bgr))
    assert np.array_equal(cv2.cvtColor(bgra, cv2.COLOR_BGRA2GRAY), orc.bgr2gray(bgra))
    s = W / 752.0
    K = np.array([[458.654 * s, 0, 367.215 * s], [0, 457.296 * s, 248.375 * H / 480.0], [0, 0, 1]], np.float32)
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
    smooth = np.stack([random_image(H, W, seed=c) for c in range(3)], -1)
    want = cv2.cvtColor(cv2.undistort(smooth, K, D), cv2.COLOR_BGR2GRAY)
    assert np.array_equal(want, orc.bgr2gray(orc.undistort_color(smooth, K, D)))


def test_harris_map_residue_of_sliding_box_sums_is_rare_and_tiny():
    """OpenCV's boxFilter slides float64 row / column sums; where a derivative is a 1-ulp rounding residue the additions are
    inexact and the residue lingers (tests/golden/make_golden.py::harris_fresh_sums).  The oracle sums every window afresh.
    Pin how far the two can be apart at the EuRoC shape: a handful of pixels per frame, none of them a selected corner."""
    import cv2
    from rd_vio_b200.synthetic import SyntheticStream
    st = SyntheticStream(17)
    cv2.setUseOptimized(False)
    try:
        total = 0
        for k in range(3):
            pre = cv2.createCLAHE(6.0, (8, 8)).apply(st.frame(k))
            R, Ro = cv2.cornerHarris(pre, 3, 3, 0.04), orc.harris(pre, 0.04, mode=0)
            d = np.argwhere(R != Ro)
            total += len(d)
            assert len(d) <= 8, f"frame {k}: {len(d)} pixels differ"
            if len(d):     # absolute size: far below the selection threshold (1e-3 of the frame maximum)
                assert np.abs(R[R != Ro] - Ro[R != Ro]).max() <= 1e-6 * float(R.max())
            ref = Cv2Image(st.frame(k))
            ref.preprocess(6.0, 8, 8)
            assert np.array_equal(orc.detect_keypoints(pre, np.zeros((0, 2)), 150, 20.0)[0],
                                  ref.detect_keypoints(np.zeros((0, 2)), 150, 20.0))
        print(f"Harris map: {total} pixels of {3 * R.size} differ from cv2 (plain mode)")
    finally:
        cv2.setUseOptimized(True)
