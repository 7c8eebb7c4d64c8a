"""GPU: the Harris stage of the hot path (integer prefilter + exact evaluation of the flagged pixels, csrc/harris.cu)
must hand the selection exactly what cv::cornerHarris + goodFeaturesToTrack's "threshold, dilate, compare" would:
bit-identical 64-bit keys of every positive 3x3 local maximum off the 1-px frame and the bit-identical frame maximum.
Checked against the oracle's response map (oracle/fe_oracle.c, pinned to cv2) in both float orders, on ordinary,
border-heavy and adversarial images, and on degenerate frames (exact whole-frame fallback inside select_kernel).
Reference call site: OpenCvImage::detect_keypoints, /root/reference/src/rdvio_extra/src/opencv_image.cpp:44."""
import numpy as np
import pytest

from conftest import random_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    from oracle import fe_oracle
    return fe_oracle


def ref_keys(R):
    """Keys of the positive local maxima (>= all 8 neighbours) off the 1-px frame, sorted descending."""
    H, W = R.shape
    q = np.pad(R, 1, constant_values=-np.inf)
    m8 = np.max([q[i:i + H, j:j + W] for i in range(3) for j in range(3) if (i, j) != (1, 1)], axis=0)
    m = (R > 0) & (R >= m8)
    m[0, :] = m[-1, :] = False
    m[:, 0] = m[:, -1] = False
    ys, xs = np.nonzero(m)
    keys = (R[ys, xs].view(np.uint32).astype(np.uint64) << np.uint64(32)) | (ys * W + xs).astype(np.uint64)
    return np.sort(keys)[::-1]


def adversarial(H, W, kind, seed=3):
    rng = np.random.default_rng(seed)
    ii = np.indices((H, W))
    if kind == "noise":
        return rng.integers(0, 256, (H, W), dtype=np.uint8)
    if kind == "binary":
        return (rng.integers(0, 2, (H, W)) * 255).astype(np.uint8)
    if kind == "lowcontrast":
        return (100 + rng.integers(0, 2, (H, W))).astype(np.uint8)
    if kind == "half_saturated":
        a = rng.integers(0, 256, (H, W), dtype=np.uint8)
        a[:, W // 2:] = 255
        return a
    if kind == "blocks":
        return np.kron((rng.integers(0, 2, (H // 8 + 1, W // 8 + 1)) * 255).astype(np.uint8), np.ones((8, 8), np.uint8))[:H, :W].copy()
    if kind == "impulses":
        a = np.zeros((H, W), np.uint8)
        a[rng.integers(0, H, 200), rng.integers(0, W, 200)] = 255
        return a
    if kind == "checker7":
        return ((((ii[0] // 7) + (ii[1] // 7)) % 2) * 255).astype(np.uint8)
    raise ValueError(kind)


def run(fe, orc, img, clahe=True, modes=(0, 1)):
    """Returns the flagged fraction; asserts keys and frame maximum in both float orders."""
    s = fe.acquire()
    try:
        if clahe:
            fe.preprocess([s], [img])
            pre = orc.clahe(img)
        else:
            fe.upload_level0(s, img)                                     # level 0 as given: no CLAHE in between
            pre = img
        frac = 0.0
        for fma in modes:
            keys, fmax, nflag = fe.harris_candidates(s, harris_fma=fma)
            R = orc.harris(pre, 0.04, mode=fma)
            want = ref_keys(R)
            assert np.float32(fmax) == max(np.float32(R.max()), np.float32(0)), f"frame maximum differs (fma={fma})"
            assert len(keys) == len(want) and np.array_equal(keys, want), \
                f"candidate keys differ (fma={fma}): got {len(keys)}, want {len(want)}"
            frac = nflag / R.size
        return frac
    finally:
        fe.release(s)


@pytest.mark.parametrize("shape", [(480, 752), (478, 750), (300, 400), (100, 152), (241, 277), (130, 296), (97, 121), (64, 56), (720, 1280)])
def test_candidates_bit_exact_shapes(orc, shape):
    from rd_vio_b200.frontend import FrontEnd
    H, W = shape
    with FrontEnd(W, H, max_level=1 if min(H, W) > 90 else 0, win=21, num_slots=2, max_points=256) as fe:
        frac = run(fe, orc, random_image(H, W, seed=W * 5 + H))
        assert frac < 0.15, f"prefilter flagged {100 * frac:.1f} % of the pixels"      # 0 when an exact-everywhere kernel is selected


@pytest.mark.parametrize("kind", ["noise", "binary", "lowcontrast", "half_saturated", "blocks", "impulses", "checker7"])
def test_candidates_bit_exact_adversarial(orc, kind):
    from rd_vio_b200.frontend import FrontEnd
    H, W = 240, 328
    with FrontEnd(W, H, max_level=1, win=21, num_slots=2, max_points=256) as fe:
        run(fe, orc, adversarial(H, W, kind), clahe=False)


def test_candidates_synthetic_stream(orc, frames0):
    from rd_vio_b200.frontend import FrontEnd
    with FrontEnd(752, 480, max_level=3, win=21, num_slots=2, max_points=256) as fe:
        for f in frames0[:3]:
            frac = run(fe, orc, f)
            assert frac < 0.10


@pytest.mark.parametrize("kind", ["constant", "stripes"])
def test_degenerate_frames_fall_back_to_the_exact_path(orc, kind):
    """Frames whose threshold lies below the rounding residue (harris_exact.cuh: rho_s) are recomputed exactly inside
    select_kernel: a constant frame, and a frame of alternating ramps with per-row 1-px ripples that has NO integer
    gradient anywhere although cv::cornerHarris sees rounding residue (min R = -7e-32).  Corners must be identical (none)."""
    from rd_vio_b200.frontend import FrontEnd
    H, W = 96, 128
    if kind == "constant":
        img = np.full((H, W), 93, np.uint8)
    else:
        # rows p_y(x) = base + s_y x + beta_y (-1)^x with s_y = +-1 alternating: [-1 0 1] sees 2 s_y and [1 2 1] sees 4 s_y x
        # whatever beta_y is, so every integer Sobel gradient vanishes although no two rows are alike
        rng = np.random.default_rng(5)
        x = np.arange(W)
        img = np.empty((H, W), np.int64)
        for y in range(H):
            sy = 1 if y % 2 == 0 else -1
            img[y] = (60 if sy > 0 else 60 + W - 1) + sy * x + rng.integers(-20, 21) * ((-1) ** x)
        img = img.astype(np.uint8)
    with FrontEnd(W, H, max_level=0, win=21, num_slots=1, max_points=256) as fe:
        s = fe.acquire()
        fe.upload_level0(s, img)
        pre = img
        for fma in (0, 1):
            ref_kp, gxy_ref, gre_ref = orc.detect_keypoints(pre, np.zeros((0, 2)), 150, 20.0, harris_mode=fma)
            kp, gxy, gre = fe.detect([s], [np.zeros((0, 2))], 150, 20.0, return_gftt=True, harris_fma=fma)
            assert np.array_equal(gxy[0], gxy_ref) and np.array_equal(gre[0], gre_ref) and np.array_equal(kp[0], ref_kp), \
                f"{kind}, fma={fma}: {len(kp[0])} corners, want {len(ref_kp)}"
