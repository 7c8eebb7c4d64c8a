"""GPU: the CUDA path against the committed golden fixtures themselves (tests/golden/*.npz, generated from cv2 4.13.0
by the committed scripts next to them) -- the same vectors that pin the oracle (tests/test_oracle_golden.py), now with
no oracle in between: bit-exact planes by SHA-256, identical corner lists, LK within the north_star tolerances."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDENS = {n: np.load(os.path.join(HERE, "golden", n)) for n in ("golden_v1.npz", "golden_752x480_v1.npz")}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_harris_map(R, G):
    """Bit-exact against cv2's plain-order map, except at the listed pixels where cv2's SLIDING float64 box sums carry a
    rounding residue (tests/golden/make_golden.py::harris_fresh_sums): there the two must agree to <= 4 ulp."""
    assert np.array_equal(R[::40], G["harris_plain_rows"])
    res = G["harris_residue_yx"]
    print(f"Harris map: {len(res)} of {R.size} pixels carry a sliding-sum residue in cv2")
    assert len(res) <= 8
    Rm = R.copy()
    for (y, x), want in zip(res, G["harris_residue_cv2"]):
        o = lambda v: (lambda b: b if b >= 0 else -(b & 0x7FFFFFFF))(int(np.float32(v).view(np.int32)))   # monotone in the value
        ulps = abs(o(R[y, x]) - o(want))
        assert ulps <= 4, (y, x, R[y, x], want)
        Rm[y, x] = 0
    assert sha(Rm) == str(G["harris_plain_sha_masked"])
    if len(res) == 0:
        assert sha(R) == str(G["harris_plain_sha"])


@pytest.fixture(scope="module", params=sorted(GOLDENS))
def G(request):
    return GOLDENS[request.param]


@pytest.fixture(scope="module")
def fe(G):
    from rd_vio_b200.frontend import FrontEnd
    H, W = G["f0"].shape
    f = FrontEnd(W, H, max_level=3, win=21, num_slots=4, max_points=512)
    yield f
    f.close()


@pytest.fixture(scope="module")
def slots(fe, G):
    s0, s1 = fe.acquire(), fe.acquire()
    fe.preprocess([s0, s1], [G["f0"], G["f1"]], 6.0, (8, 8))
    return s0, s1


def test_clahe_and_pyramid_golden(fe, slots, G):
    s0, _ = slots
    assert sha(fe.download_level(s0, 0, 0)) == str(G["clahe_f0_sha"])
    assert 2 * fe.nlevels == int(G["n_pyr_planes"])
    for l in range(fe.nlevels):
        img, der = fe.download_level(s0, l, 0), fe.download_level(s0, l, 1)
        assert img.shape == tuple(G[f"pyr{2 * l}_shape"]) and der.shape == tuple(G[f"pyr{2 * l + 1}_shape"])
        assert sha(img) == str(G[f"pyr{2 * l}_sha"]), f"image level {l}"
        assert sha(der) == str(G[f"pyr{2 * l + 1}_sha"]), f"Scharr level {l}"


def test_clahe_other_parameters_golden(fe, G):
    from rd_vio_b200.frontend import FrontEnd
    s = fe.acquire()
    try:
        fe.preprocess([s], [G["f0"]], 2.0, (4, 6))
        assert sha(fe.download_level(s, 0, 0)) == str(G["clahe_46_sha"])
    finally:
        fe.release(s)
    H, W = G["f0"].shape
    odd = np.ascontiguousarray(G["f0"][:H - 3, :W - 5])           # CLAHE's padding quirk
    with FrontEnd(W - 5, H - 3, max_level=3, win=21, num_slots=1, max_points=64) as f2:
        s = f2.acquire()
        f2.preprocess([s], [odd], 6.0, (8, 8))
        assert sha(f2.download_level(s, 0, 0)) == str(G["clahe_odd_sha"])


def test_harris_and_detect_golden(fe, slots, G):
    s0, _ = slots
    check_harris_map(fe.harris_response(s0), G)
    kp, gxy, gre = fe.detect([s0], [np.zeros((0, 2))], 150, 20.0, return_gftt=True)
    assert np.array_equal(gxy[0], G["gftt_xy"]) and np.array_equal(gre[0], G["gftt_resp"])
    assert np.array_equal(kp[0], G["detect_empty"])
    assert np.array_equal(fe.detect([s0], [G["existing"]], 150, 20.0)[0], G["detect_existing_r20"])
    assert np.array_equal(fe.detect([s0], [G["existing"]], 150, 10.0)[0], G["detect_existing_r10"])


def test_track_golden(fe, slots, G):
    s0, s1 = slots
    for pred, kn, ks in ((G["lk_pred"], "track_next", "track_status"), (None, "track_next_nopred", "track_status_nopred")):
        nxt, st = fe.track([s0], [s1], [G["lk_pts"]], [pred] if pred is not None else None)
        nxt, st = nxt[0], st[0]
        assert (st == G[ks]).mean() >= 0.995
        ok = (st != 0) & (G[ks] != 0)
        assert ok.sum() > 50
        assert np.abs(nxt[ok] - G[kn][ok]).max() <= 0.01


def test_undistort_golden(fe, G):
    U = np.load(os.path.join(HERE, "golden", "golden_undistort_v1.npz"))
    if G["f0"].shape != (240, 320):
        pytest.skip("the undistortion fixture was generated for the 320x240 frame")
    s = fe.acquire()
    try:
        fe.set_undistort(U["K"], U["D"])
        fe.preprocess([s], [G["f0"]], 6.0, (8, 8))
        out = fe.download_level(s, 0, 3)
        assert np.array_equal(out[::30], U["rows"])
        assert sha(out) == str(U["undistorted_sha"])
    finally:
        fe.set_undistort(None, None)
        fe.release(s)
