"""CPU: the oracle (oracle/fe_oracle.c) against the committed golden vectors (tests/golden/golden_v1.npz,
generated from cv2 4.13.0 by tests/golden/make_golden.py).  These pin the oracle; the GPU tests then
compare the CUDA path with the oracle."""
import hashlib
import os

import numpy as np
import pytest

from oracle import fe_oracle as orc

GOLDENS = {n: np.load(os.path.join(os.path.dirname(__file__), "golden", n))
           for n in ("golden_v1.npz", "golden_752x480_v1.npz")}       # 320x240 and the EuRoC shape of BASELINE configs[0]/[1]


@pytest.fixture(params=sorted(GOLDENS))
def G(request):
    return GOLDENS[request.param]


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


def test_clahe_golden(G):
    f0 = G["f0"]
    assert sha(orc.clahe(f0, 6.0, 8, 8)) == str(G["clahe_f0_sha"])
    H, W = f0.shape
    assert sha(orc.clahe(np.ascontiguousarray(f0[:H - 3, :W - 5]), 6.0, 8, 8)) == str(G["clahe_odd_sha"])   # padding quirk
    assert sha(orc.clahe(f0, 2.0, 4, 6)) == str(G["clahe_46_sha"])


def test_pyramid_golden(G):
    P = orc.Pyramid(orc.clahe(G["f0"]), 21, 3)
    assert 2 * P.nlevels == int(G["n_pyr_planes"])
    for l in range(P.nlevels):
        assert tuple(G[f"pyr{2 * l}_shape"]) == P.image(l).shape
        assert sha(P.image(l)) == str(G[f"pyr{2 * l}_sha"]), f"image level {l}"
        assert sha(P.deriv(l)) == str(G[f"pyr{2 * l + 1}_sha"]), f"Scharr level {l}"


def test_harris_golden(G):
    check_harris_map(orc.harris(orc.clahe(G["f0"]), 0.04, mode=0), G)


def test_gftt_and_detect_golden(G):
    pre = orc.clahe(G["f0"])
    kp, gxy, gre = orc.detect_keypoints(pre, np.zeros((0, 2)), 150, 20.0)
    assert np.array_equal(gxy, G["gftt_xy"]) and np.array_equal(gre, G["gftt_resp"])
    assert np.array_equal(kp, G["detect_empty"])
    assert np.array_equal(orc.detect_keypoints(pre, G["existing"], 150, 20.0)[0], G["detect_existing_r20"])
    assert np.array_equal(orc.detect_keypoints(pre, G["existing"], 150, 10.0)[0], G["detect_existing_r10"])


def test_dispatched_mode_disagreement_is_reported(G):
    """OpenCV's AVX2-dispatched float order (FMA in Sobel) is NOT the parity target; report how far it is."""
    a, b = G["detect_empty"], G["detect_empty_dispatched"]
    sa, sb = set(map(tuple, a)), set(map(tuple, b))
    print(f"plain vs dispatched keypoint sets: {len(sa ^ sb)} of {len(sa)} differ")
    assert len(sa ^ sb) <= max(2, len(sa) // 10)


def test_lk_and_track_golden(G):
    PA, PB = orc.Pyramid(orc.clahe(G["f0"])), orc.Pyramid(orc.clahe(G["f1"]))
    q, st = orc.lk(PA, PB, G["lk_pts"].astype(np.float32), G["lk_pred"].astype(np.float32))
    assert (st == G["lk_raw_status"]).mean() >= 0.995
    ok = (st != 0) & (G["lk_raw_status"] != 0)
    assert np.abs(q[ok] - G["lk_raw_xy"][ok]).max() <= 0.01
    for pred, kn, ks in ((G["lk_pred"], "track_next", "track_status"), (None, "track_next_nopred", "track_status_nopred")):
        nxt, s, _ = orc.track_keypoints(PA, PB, G["lk_pts"], pred)
        assert (s == G[ks]).mean() >= 0.995
        ok = (s != 0) & (G[ks] != 0)
        assert ok.sum() > 50
        assert np.abs(nxt[ok] - G[kn][ok]).max() <= 0.01


def test_poisson_filter_literal_loop():
    """The reference's grid walk (poisson_disk_filter.h:73-94) vs brute force over 'visible' points."""
    rng = np.random.default_rng(3)
    for r in (10.0, 20.0, 33.3):
        ex = rng.uniform(0, 300, (60, 2))
        ex[10] = ex[3] + 0.5            # two presets in one cell: the later one hides the earlier one
        cand = rng.uniform(0, 300, (200, 2))
        got = orc.poisson_filter(ex, cand, r)
        g = r / np.sqrt(2.0)
        cells = np.floor(ex / g).astype(int)
        vis = [i for i in range(len(ex)) if not any((cells[j] == cells[i]).all() for j in range(i + 1, len(ex)))]
        pts = [ex[i] for i in vis]
        want = []
        for c in cand:
            if all(((c - p) ** 2).sum() >= r * r for p in pts):
                pts.append(c)
                want.append(c)
        assert np.array_equal(got, np.array(want).reshape(-1, 2))


def test_undistort_golden():
    """cv::undistort in front of the plugin (examples/dataset.hpp:232-236)."""
    U = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_undistort_v1.npz"))
    out = orc.undistort(GOLDENS["golden_v1.npz"]["f0"], U["K"], U["D"])
    assert np.array_equal(out[::30], U["rows"])
    assert sha(out) == str(U["undistorted_sha"])
