"""Generates tests/golden/golden_v1.npz from the REAL reference arithmetic: Python cv2 (opencv-python-headless,
the same OpenCV functions rdvio::extra::OpenCvImage calls, /root/reference/src/rdvio_extra/src/opencv_image.cpp).
Run in the authoring container:  python tests/golden/make_golden.py
Bit-exact planes are stored as SHA-256 digests (inputs are stored verbatim), float/keypoint outputs as arrays.
cv2.setUseOptimized(False) is used for the Harris-dependent outputs: that is OpenCV's plain C++ float order,
the parity target (SURVEY.md 8(c) "Oracle modes"); the dispatched-mode keypoints are stored too, for reporting.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import random_image  # noqa: E402
from oracle.cv2_reference import Cv2Image  # noqa: E402


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def harris_fresh_sums(img):
    """cornerHarris restated from cv2's own Sobel output with every 3x3 box sum taken afresh in float64 (plain float
    order otherwise).  cv2's boxFilter instead SLIDES its float64 row / column sums (s += new - old), and a derivative that
    is a 1-ulp rounding residue instead of 0 makes products of ~1e-18 next to ~1e-3: such additions are inexact, the
    residue stays in the running sum for the rest of the row / column, and once in a few frames it flips the float
    rounding of a later window by an ulp or two.  Those pixels are listed in the fixture instead of being hashed."""
    sc = 1.0 / (4 * 3 * 255)
    Dx = cv2.Sobel(img, cv2.CV_32F, 1, 0, ksize=3, scale=sc)
    Dy = cv2.Sobel(img, cv2.CV_32F, 0, 1, ksize=3, scale=sc)

    def box(m):
        pad = np.pad(m.astype(np.float64), 1, mode="reflect")
        h, w = m.shape
        acc = np.zeros((h, w))
        for dy in range(3):
            for dx in range(3):
                acc += pad[dy:dy + h, dx:dx + w]
        return acc.astype(np.float32)

    A, B, C = box(Dx * Dx), box(Dx * Dy), box(Dy * Dy)
    k = np.float32(0.04)
    return (A * C - B * B) - (k * (A + C)) * (A + C)


def main(H=240, W=320, name="golden_v1.npz", seed=2024):
    base = random_image(H + 16, W + 16, seed=seed)
    f0 = np.ascontiguousarray(base[8:8 + H, 8:8 + W])
    # second frame: small shift + rotation + gain, rendered with cv2.warpAffine (stored verbatim)
    M = cv2.getRotationMatrix2D((W / 2 + 8, H / 2 + 8), 0.7, 1.004)
    M[:, 2] += (1.8, -1.1)
    warped = cv2.warpAffine(base, M, (W + 16, H + 16), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    f1 = np.clip(np.rint(warped[8:8 + H, 8:8 + W].astype(np.float32) * 1.05 + 3), 0, 255).astype(np.uint8)
    out = {"f0": f0, "f1": f1, "cv2_version": np.array(cv2.__version__)}

    # CLAHE incl. the padding quirk (odd crop)
    out["clahe_f0_sha"] = np.array(digest(cv2.createCLAHE(6.0, (8, 8)).apply(f0)))
    odd = np.ascontiguousarray(f0[:H - 3, :W - 5])
    out["clahe_odd_sha"] = np.array(digest(cv2.createCLAHE(6.0, (8, 8)).apply(odd)))
    out["clahe_46_sha"] = np.array(digest(cv2.createCLAHE(2.0, (4, 6)).apply(f0)))

    cv2.setUseOptimized(False)
    A, B = Cv2Image(f0), Cv2Image(f1)
    A.preprocess(6.0, 8, 8)
    B.preprocess(6.0, 8, 8)
    out["n_pyr_planes"] = np.array(len(A.pyramid))
    for i, p in enumerate(A.pyramid):
        out[f"pyr{i}_shape"] = np.array(p.shape)
        out[f"pyr{i}_sha"] = np.array(digest(p))
    R = cv2.cornerHarris(A.image, 3, 3, 0.04)
    out["harris_plain_sha"] = np.array(digest(R))
    # pixels where cv2's sliding box sums left a rounding residue (see harris_fresh_sums): listed, and masked out of a
    # second digest, so that a fresh-sum implementation can be checked exactly everywhere else
    res = np.argwhere(R != harris_fresh_sums(A.image))
    out["harris_residue_yx"] = res.astype(np.int32).reshape(-1, 2)
    out["harris_residue_cv2"] = np.array([R[y, x] for y, x in res], np.float32)
    Rm = R.copy()
    for y, x in res:
        Rm[y, x] = 0
    out["harris_plain_sha_masked"] = np.array(digest(Rm))
    out["harris_plain_rows"] = R[::40].copy()          # a few rows verbatim, for diagnostics
    det = cv2.GFTTDetector_create(150, 1.0e-3, 20, 3, True)
    kps = det.detect(A.image)
    out["gftt_xy"] = np.array([k.pt for k in kps], np.float32)
    out["gftt_resp"] = np.array([k.response for k in kps], np.float32)
    kp0 = A.detect_keypoints(np.zeros((0, 2)), 150, 20.0)
    out["detect_empty"] = kp0
    existing = kp0[::4] + 0.25
    out["existing"] = existing
    out["detect_existing_r20"] = A.detect_keypoints(existing, 150, 20.0)
    out["detect_existing_r10"] = A.detect_keypoints(existing, 150, 10.0)
    cv2.setUseOptimized(True)
    out["detect_empty_dispatched"] = A.detect_keypoints(np.zeros((0, 2)), 150, 20.0)

    # LK: tracked corners + edge cases; prediction = affine model + noise
    rng = np.random.default_rng(11)
    extra = np.array([[3., 3.], [W - 1., H - 1.], [-25., 50.], [150., -22.], [W + 5., 100.], [160.5, 120.25], [20., 20.]])
    pts = np.concatenate([kp0, extra], 0)
    Minv = cv2.invertAffineTransform(M)
    # f1(x) = base(Minv-mapped) => a point p of f0 (base coords p+8) appears in f1 at M*(p+8) - 8
    pred = (np.c_[pts + 8, np.ones(len(pts))] @ M.T) - 8 + rng.normal(0, 1.0, pts.shape)
    del Minv
    nxt, st = A.track_keypoints(B, pts, pred)
    out["lk_pts"], out["lk_pred"], out["track_next"], out["track_status"] = pts, pred, nxt, st
    out["track_forward_raw"] = A.last_forward.copy()
    nxt2, st2 = A.track_keypoints(B, pts, None)
    out["track_next_nopred"], out["track_status_nopred"] = nxt2, st2
    q, s, _ = cv2.calcOpticalFlowPyrLK(A.image, B.image, pts.astype(np.float32).reshape(-1, 1, 2),
                                       pred.astype(np.float32).reshape(-1, 1, 2), winSize=(21, 21), maxLevel=3,
                                       criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                       flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    out["lk_raw_xy"], out["lk_raw_status"] = q.reshape(-1, 2), s.reshape(-1)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(kp0), "corners,", int(st.sum()), "tracked")


if __name__ == "__main__":
    main()                                                     # 320x240 (round 1)
    main(480, 752, "golden_752x480_v1.npz", seed=2025)         # BASELINE configs[0]/[1] shape (EuRoC)
