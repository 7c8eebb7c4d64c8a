"""GPU parity tests: the CUDA path (through the C ABI, librdvio_fe.so) against the CPU oracle
(oracle/fe_oracle.c, itself pinned against cv2 in test_oracle_*.py) on identical seeded inputs.

Bars (BASELINE.json north_star): preprocessed images, LUTs, pyramids, derivatives, halos and the
Harris response map bit-exact; selected keypoints identical (set AND order); tracked positions
within 0.01 px with status agreement >= 99.5 %.
"""
import numpy as np
import pytest

from conftest import random_image

pytestmark = pytest.mark.gpu

TOL_PX = 0.01          # north_star tolerance for tracked positions
MIN_STATUS_AGREE = 0.995


@pytest.fixture(scope="module")
def orc():
    from oracle import fe_oracle
    return fe_oracle


@pytest.fixture(scope="module")
def fe752():
    from rd_vio_b200.frontend import FrontEnd
    fe = FrontEnd(752, 480, max_level=3, win=21, num_slots=16, max_points=512)
    yield fe
    fe.close()


def _pre(fe, img, **kw):
    s = fe.acquire()
    fe.preprocess([s], [img], **kw)
    return s


@pytest.mark.parametrize("shape,tiles,clip", [
    ((480, 752), (8, 8), 6.0),      # EuRoC, tiles divide evenly
    ((478, 750), (8, 8), 6.0),      # both dims padded
    ((480, 750), (8, 8), 6.0),      # the "divisible dim still gets a full extra tile row" quirk
    ((477, 752), (8, 8), 6.0),
    ((135, 241), (8, 8), 6.0),
    ((200, 300), (4, 6), 2.0),
    ((720, 1280), (8, 8), 6.0),     # ADVIO-shaped
])
def test_clahe_bit_exact(orc, shape, tiles, clip):
    from rd_vio_b200.frontend import FrontEnd
    H, W = shape
    img = random_image(H, W, seed=H * 7 + W)
    ref, ref_lut = orc.clahe(img, clip, tiles[0], tiles[1], return_lut=True)
    with FrontEnd(W, H, max_level=1, win=21, num_slots=2, max_points=64) as fe:
        s = _pre(fe, img, clip_limit=clip, tiles=tiles)
        lut = fe.download_clahe_lut(0, tiles[0] * tiles[1])
        got = fe.download_level(s, 0, 0)
    assert np.array_equal(lut, ref_lut), f"LUT differs in {(lut != ref_lut).sum()} entries"
    assert np.array_equal(got, ref), f"CLAHE image differs in {(got != ref).sum()} px of {got.size}"


@pytest.mark.parametrize("shape,max_level,win", [
    ((480, 752), 3, 21),
    ((720, 1280), 4, 21),
    ((1080, 1920), 5, 31),
    ((479, 751), 3, 21),           # odd sizes: ceil halving + reflect at odd borders
    ((100, 130), 3, 21),           # pyramid truncates early (next level <= win)
])
def test_pyramid_bit_exact(orc, shape, max_level, win):
    from rd_vio_b200.frontend import FrontEnd
    H, W = shape
    img = random_image(H, W, seed=H + W)
    with FrontEnd(W, H, max_level=max_level, win=win, num_slots=2, max_points=64) as fe:
        s = _pre(fe, img)
        P = orc.Pyramid(orc.clahe(img), win, max_level)
        assert fe.nlevels == P.nlevels
        for l in range(P.nlevels):
            im, dv, halo = fe.download_level(s, l, 0), fe.download_level(s, l, 1), fe.download_level(s, l, 2)
            assert im.shape == P.level_shape(l)
            assert np.array_equal(im, P.image(l)), f"level {l} image differs ({(im != P.image(l)).sum()} px)"
            assert np.array_equal(dv, P.deriv(l)), f"level {l} Scharr differs ({(dv != P.deriv(l)).sum()} values)"
            want = np.pad(P.image(l), win, mode="reflect")          # numpy 'reflect' == BORDER_REFLECT_101
            assert np.array_equal(halo, want), f"level {l} halo differs ({(halo != want).sum()} px)"


@pytest.mark.parametrize("fma", [0, 1])
def test_harris_response_bit_exact(orc, fe752, frames0, fma):
    s = _pre(fe752, frames0[0])
    try:
        got = fe752.harris_response(s, harris_fma=fma)
        ref = orc.harris(orc.clahe(frames0[0]), 0.04, mode=fma)
        assert np.array_equal(got, ref), f"Harris map differs in {(got != ref).sum()} px, max |d|={np.abs(got - ref).max()}"
    finally:
        fe752.release(s)


@pytest.mark.parametrize("n_existing,radius,max_points", [(0, 20.0, 150), (47, 20.0, 150), (47, 10.0, 200), (0, 20.0, 1), (47, 30.0, 150), (20, 45.5, 150)])
def test_detect_identical(orc, fe752, frames0, n_existing, radius, max_points):
    s = _pre(fe752, frames0[1])
    try:
        pre = orc.clahe(frames0[1])
        base, _, _ = orc.detect_keypoints(pre, np.zeros((0, 2)), 150, 20.0)
        ex = base[::3][:n_existing] + 0.3 if n_existing else np.zeros((0, 2))
        ref, gxy_ref, gre_ref = orc.detect_keypoints(pre, ex, max_points, radius)
        got, gxy, gre = fe752.detect([s], [ex], max_points, radius, return_gftt=True)
        assert np.array_equal(gxy[0], gxy_ref), "GFTT corner list (set or order) differs"
        assert np.array_equal(gre[0], gre_ref), "GFTT responses differ"
        assert np.array_equal(got[0], ref), f"keypoints differ: got {len(got[0])}, want {len(ref)}"
    finally:
        fe752.release(s)


def test_detect_prefetch_is_transparent(orc, fe752, frames0):
    """rdfe_detect_prefetch never changes results: used when slots and GFTT parameters match, ignored otherwise,
    dropped when the slot is preprocessed again or released."""
    s = _pre(fe752, frames0[0])
    try:
        pre0 = orc.clahe(frames0[0])
        existing = orc.detect_keypoints(pre0, np.zeros((0, 2)), 40, 20.0)[0] + 3.25
        ref = orc.detect_keypoints(pre0, existing, 150, 20.0)[0]
        assert np.array_equal(fe752.detect([s], [existing], 150, 20.0)[0], ref)
        fe752.detect_prefetch([s], 150)
        assert np.array_equal(fe752.detect([s], [existing], 150, 20.0)[0], ref)           # consumed
        fe752.detect_prefetch([s], 150)
        assert np.array_equal(fe752.detect([s], [existing], 150, 12.0)[0],               # Poisson radius may differ
                              orc.detect_keypoints(pre0, existing, 150, 12.0)[0])
        fe752.detect_prefetch([s], 100)                                                   # other max_points: ignored
        assert np.array_equal(fe752.detect([s], [existing], 150, 20.0)[0], ref)
        fe752.detect_prefetch([s], 150)                                                   # stale after a new preprocess
        fe752.preprocess([s], [frames0[1]])
        ref1 = orc.detect_keypoints(orc.clahe(frames0[1]), np.zeros((0, 2)), 150, 20.0)[0]
        assert np.array_equal(fe752.detect([s], [np.zeros((0, 2))], 150, 20.0)[0], ref1)
        fe752.detect_prefetch([s], 150)                                                   # stale after release + reuse
    finally:
        fe752.release(s)
    s2 = _pre(fe752, frames0[2])
    try:
        ref2 = orc.detect_keypoints(orc.clahe(frames0[2]), np.zeros((0, 2)), 150, 20.0)[0]
        assert np.array_equal(fe752.detect([s2], [np.zeros((0, 2))], 150, 20.0)[0], ref2)
    finally:
        fe752.release(s2)


def _track_case(orc, fe, f0, f1, pts, pred, win=21, max_level=3):
    s0, s1 = _pre(fe, f0), _pre(fe, f1)
    try:
        PA, PB = orc.Pyramid(orc.clahe(f0), win, max_level), orc.Pyramid(orc.clahe(f1), win, max_level)
        ref_xy, ref_st, _ = orc.track_keypoints(PA, PB, pts, pred, win, max_level)
        got_xy, got_st = fe.track([s0], [s1], [pts], [pred] if pred is not None else None)
        got_xy, got_st = got_xy[0], got_st[0]
        agree = (got_st == ref_st).mean()
        both = (got_st != 0) & (ref_st != 0)
        err = np.abs(got_xy[both] - ref_xy[both]).max() if both.any() else 0.0
        return agree, err, int(ref_st.sum()), got_xy, got_st, ref_xy, ref_st
    finally:
        fe.release(s0)
        fe.release(s1)


def test_track_parity(orc, fe752, stream0, frames0):
    pre = orc.clahe(frames0[0])
    pts, _, _ = orc.detect_keypoints(pre, np.zeros((0, 2)), 150, 20.0)
    # plus border / outside / sub-pixel cases
    extra = np.array([[5., 5.], [751., 479.], [-30., 100.], [400., -25.], [760., 300.], [375.5, 240.25],
                      [20.0, 20.0], [731.9, 459.9], [21.3, 240.7]])
    pts = np.concatenate([pts, extra], 0)
    pred = stream0.predict(0, pts)
    agree, err, nok, *_ = _track_case(orc, fe752, frames0[0], frames0[1], pts, pred)
    assert nok > 100
    assert agree >= MIN_STATUS_AGREE, f"status agreement {agree:.4f}"
    assert err <= TOL_PX, f"max position error {err:.2e} px"
    # without prediction (next_keypoints empty branch, opencv_image.cpp:81-86)
    agree, err, nok, *_ = _track_case(orc, fe752, frames0[0], frames0[1], pts, None)
    assert agree >= MIN_STATUS_AGREE and err <= TOL_PX, (agree, err)


def test_track_negative_fourth_weight(orc, fe752, stream0, frames0):
    """Sub-pixel offsets at which the three rounded Q14 weights sum to 2^14 + 1, so the fourth is -1 (OpenCV
    keeps it signed).  Found by a 200-frame replay: treating it as 65535 made a few points diverge."""
    hits = [(1, 8184), (2, 4093), (3, 2728), (5, 1638), (8, 1023), (13, 630), (21, 390), (30, 273), (45, 182),
            (63, 130), (88, 93), (90, 91), (105, 78), (117, 70)]
    f = np.float32
    for i, j in hits:
        a, b = f(i) * f(2.0 ** -14), f(j) * f(2.0 ** -14)
        w = [np.rint((f(1) - a) * (f(1) - b) * f(16384)), np.rint(a * (f(1) - b) * f(16384)), np.rint((f(1) - a) * b * f(16384))]
        assert 16384 - sum(w) == -1
    pre = orc.clahe(frames0[0])
    corners, _, _ = orc.detect_keypoints(pre, np.zeros((0, 2)), 150, 20.0)
    pts = []
    for k, c in enumerate(corners):
        i, j = hits[k % len(hits)]
        if k % 2:
            i, j = j, i
        pts.append([c[0] + i * 2.0 ** -14, c[1] + j * 2.0 ** -14])
    pts = np.array(pts)
    for pred in (None, stream0.predict(0, pts)):
        agree, err, nok, got_xy, got_st, ref_xy, ref_st = _track_case(orc, fe752, frames0[0], frames0[1], pts, pred)
        assert nok > 100
        assert np.array_equal(got_st, ref_st)
        assert err == 0.0


def test_track_large_motion_and_failures(orc, fe752, stream0, frames0):
    """Frames 3 apart with a poor guess: exercises restaging of the search region, the 0.5-px
    round-trip gate and LK failures."""
    pre = orc.clahe(frames0[0])
    pts, _, _ = orc.detect_keypoints(pre, np.zeros((0, 2)), 150, 20.0)
    rng = np.random.default_rng(5)
    pred = pts + rng.normal(0, 6.0, pts.shape)
    agree, err, nok, *_ = _track_case(orc, fe752, frames0[0], frames0[3], pts, pred)
    assert agree >= MIN_STATUS_AGREE, f"status agreement {agree:.4f} (ok in ref: {nok})"
    assert err <= TOL_PX, f"max position error {err:.2e} px"


def test_batched_equals_single(orc, fe752, stream0, frames0):
    """n images in one call give exactly what n single calls give (streams share no state)."""
    n = 4
    slots = [fe752.acquire() for _ in range(n + 1)]
    try:
        fe752.preprocess(slots, frames0[:n + 1])
        kps = fe752.detect(slots[:n], [np.zeros((0, 2))] * n, 150, 20.0)
        nxt, st = fe752.track(slots[:n], slots[1:], kps, None)
        for i in range(n):
            k1 = fe752.detect([slots[i]], [np.zeros((0, 2))], 150, 20.0)[0]
            assert np.array_equal(k1, kps[i])
            n1, s1 = fe752.track([slots[i]], [slots[i + 1]], [kps[i]], None)
            assert np.array_equal(s1[0], st[i]) and np.array_equal(n1[0], nxt[i])
            ref = orc.detect_keypoints(orc.clahe(frames0[i]), np.zeros((0, 2)), 150, 20.0)[0]
            assert np.array_equal(kps[i], ref)
    finally:
        for s in slots:
            fe752.release(s)


def test_large_window_config(orc):
    """Config 4 shape: 1920x1080, maxLevel 5, 31x31 window (GPU-only setting; oracle = same restatement)."""
    from rd_vio_b200.frontend import FrontEnd
    from rd_vio_b200.synthetic import SyntheticStream
    st = SyntheticStream(3, 1920, 1080)
    f0, f1 = st.frame(0), st.frame(1)
    with FrontEnd(1920, 1080, max_level=5, win=31, num_slots=2, max_points=1200) as fe:
        pre = orc.clahe(f0)
        pts, _, _ = orc.detect_keypoints(pre, np.zeros((0, 2)), 1000, 20.0)
        got = fe_detect = None
        s0, s1 = _pre(fe, f0), _pre(fe, f1)
        got = fe.detect([s0], [np.zeros((0, 2))], 1000, 20.0)[0]
        assert np.array_equal(got, pts), f"detect differs: {len(got)} vs {len(pts)}"
        pred = st.predict(0, pts)
        PA, PB = orc.Pyramid(pre, 31, 5), orc.Pyramid(orc.clahe(f1), 31, 5)
        ref_xy, ref_st, _ = orc.track_keypoints(PA, PB, pts, pred, 31, 5)
        got_xy, got_st = fe.track([s0], [s1], [pts], [pred])
        agree = (got_st[0] == ref_st).mean()
        both = (got_st[0] != 0) & (ref_st != 0)
        err = np.abs(got_xy[0][both] - ref_xy[both]).max()
        assert agree >= MIN_STATUS_AGREE and err <= TOL_PX, (agree, err)


def test_fused_step_equals_separate_calls(orc, stream0, frames0):
    """rdfe_frontend_step_dev (GFTT selection overlapped with LK on a second stream) == preprocess + track + detect."""
    import ctypes as C
    import torch
    from rd_vio_b200 import _native as N
    from rd_vio_b200.frontend import FrontEnd
    L = N.lib()
    n, stride = 3, 300
    ts = torch.cuda.Stream()
    with torch.cuda.stream(ts):
        fe = FrontEnd(752, 480, 3, 21, num_slots=2 * n, max_points=512, stream=ts.cuda_stream)
        prev = np.array([fe.acquire() for _ in range(n)], np.int32)
        new = np.array([fe.acquire() for _ in range(n)], np.int32)
        fe.preprocess(list(prev), frames0[0:n])
        kps = fe.detect(list(prev), [np.zeros((0, 2))] * n, 150, 20.0)
        preds = [stream0.predict(i, kps[i]) for i in range(n)]
        # separate calls (host API) as the expectation
        fe.preprocess(list(new), frames0[1:n + 1])
        nxt, st = fe.track(list(prev), list(new), kps, preds)
        want = []
        for i in range(n):      # Frame::track_keypoints carries only the status != 0 points into detect (frame.cpp:160-170)
            want.append(fe.detect([int(new[i])], [nxt[i][st[i] != 0]], 150, 20.0, stride=stride)[0])
        # fused device call
        curr = torch.zeros((n, stride, 2), dtype=torch.float64)
        work = torch.zeros((n, stride, 2), dtype=torch.float64)
        cnt = torch.zeros(n, dtype=torch.int32)
        for i in range(n):
            curr[i, :len(kps[i])] = torch.from_numpy(kps[i])
            work[i, :len(kps[i])] = torch.from_numpy(preds[i])
            cnt[i] = len(kps[i])
        curr, work, cnt = curr.cuda(), work.cuda(), cnt.cuda()
        kcnt = cnt.clone()
        status = torch.zeros((n, stride), dtype=torch.int8, device="cuda")
        imgs = torch.from_numpy(np.stack(frames0[1:n + 1])).cuda()
        ptrs = (C.c_void_p * n)(*[imgs[i].data_ptr() for i in range(n)])
        tp, dp = fe.track_params(has_prediction=1), fe.detect_params(max_points=150, keypoint_distance=20.0)
        N.check(L.rdfe_frontend_step_dev(fe.handle, prev.ctypes.data, new.ctypes.data, n, ptrs, 752, 6.0, 8, 8, C.byref(tp),
                                         C.c_void_p(curr.data_ptr()), C.c_void_p(work.data_ptr()), C.c_void_p(cnt.data_ptr()),
                                         C.c_void_p(status.data_ptr()), C.byref(dp), C.c_void_p(kcnt.data_ptr()), stride),
                "frontend_step")
        fe.sync()
        got_xy, got_cnt, got_st = work.cpu().numpy(), kcnt.cpu().numpy(), status.cpu().numpy()
        for i in range(n):
            assert np.array_equal(got_st[i, :len(kps[i])], st[i])
            assert got_cnt[i] == len(want[i])
            assert np.array_equal(got_xy[i, :got_cnt[i]], want[i])
        fe.close()


def test_pipelined_host_step_equals_separate_calls(orc, stream0, frames0):
    """rdfe_frontend_step_submit/_wait with three steps in flight (the pipeline depth) == the per-call host API."""
    import ctypes as C
    from rd_vio_b200 import _native as N
    from rd_vio_b200.frontend import FrontEnd
    L = N.lib()
    n, stride = 2, 300
    with FrontEnd(752, 480, 3, 21, num_slots=4 * n, max_points=512) as fe:
        # four slot sets: no step rewrites a set the expectation phase still needs
        sets = [np.array([fe.acquire() for _ in range(n)], np.int32) for _ in range(4)]
        fe.preprocess(list(sets[0]), frames0[0:n])
        kps = fe.detect(list(sets[0]), [np.zeros((0, 2))] * n, 150, 20.0)
        # expectation: three consecutive steps through the separate host calls
        want = []
        carried = kps
        for stp in range(3):
            a, b = sets[stp], sets[stp + 1]
            fe.preprocess(list(b), frames0[stp + 1:stp + 1 + n])
            nxt, st = fe.track(list(a), list(b), carried, None)
            merged = []
            for i in range(n):  # only tracked points are carried (frame.cpp:160-170)
                merged.append(fe.detect([int(b[i])], [nxt[i][st[i] != 0]], 150, 20.0, stride=stride)[0])
            want.append((merged, st))
            carried = [m[:stride] for m in merged]
        # pipelined: submit steps 0, 1, 2 back to back (each step's input = the previous step's expected output), then wait
        imgs = [np.stack(frames0[s + 1:s + 1 + n]) for s in range(3)]
        tp, dp = fe.track_params(), fe.detect_params(max_points=150, keypoint_distance=20.0)
        tickets, bufs = [], []
        inputs = [kps, want[0][0], want[1][0]]
        for stp in range(3):
            a, b = sets[stp], sets[stp + 1]
            curr = np.zeros((n, stride, 2)); cnt = np.zeros(n, np.int32)
            for i in range(n):
                c = inputs[stp][i][:stride]
                curr[i, :len(c)] = c; cnt[i] = len(c)
            ptrs = (C.c_void_p * n)(*[imgs[stp][i].ctypes.data for i in range(n)])
            tk = C.c_int()
            N.check(L.rdfe_frontend_step_submit(fe.handle, a.ctypes.data, b.ctypes.data, n, ptrs, 752, 6.0, 8, 8, C.byref(tp),
                                                curr.ctypes.data, None, cnt.ctypes.data, C.byref(dp), stride, C.byref(tk)), "submit")
            tickets.append(tk.value); bufs.append((curr, cnt, ptrs))
        for stp in range(3):
            out = np.zeros((n, stride, 2)); oc = np.zeros(n, np.int32); ost = np.zeros((n, stride), np.int8)
            N.check(L.rdfe_frontend_step_wait(fe.handle, tickets[stp], out.ctypes.data, oc.ctypes.data, ost.ctypes.data), "wait")
            for i in range(n):
                m, st = want[stp][0][i], want[stp][1][i]
                assert oc[i] == len(m), (stp, i, oc[i], len(m))
                assert np.array_equal(out[i, :oc[i]], m)
                assert np.array_equal(ost[i, :len(st)], st)


def test_cross_step_pipelining_equals_serial(stream0, frames0):
    """rdfe_set_pipelining(1) with three slot sets: preprocess of step t+1 overlaps step t; results must equal
    the serial schedule bit for bit."""
    import ctypes as C
    import torch
    from rd_vio_b200 import _native as N
    from rd_vio_b200.frontend import FrontEnd
    L = N.lib()
    n, stride, steps = 4, 300, 5
    ts = torch.cuda.Stream()
    results = {}
    with torch.cuda.stream(ts):
        for mode in (0, 1):
            fe = FrontEnd(752, 480, 3, 21, num_slots=3 * n, max_points=512, stream=ts.cuda_stream)
            N.check(L.rdfe_set_pipelining(fe.handle, mode), "set_pipelining")
            sets = [np.array([fe.acquire() for _ in range(n)], np.int32) for _ in range(3)]
            imgs = torch.from_numpy(np.stack([np.stack([frames0[(s + i) % 6] for i in range(n)]) for s in range(steps + 1)])).cuda()
            fe.preprocess(list(sets[0]), [frames0[i % 6] for i in range(n)])
            kps = fe.detect(list(sets[0]), [np.zeros((0, 2))] * n, 150, 20.0)
            curr = torch.zeros((n, stride, 2), dtype=torch.float64)
            cnt = torch.zeros(n, dtype=torch.int32)
            for i in range(n):
                curr[i, :len(kps[i])] = torch.from_numpy(kps[i]); cnt[i] = len(kps[i])
            curr, cnt = curr.cuda(), cnt.cuda()
            tp, dp = fe.track_params(has_prediction=0), fe.detect_params(max_points=150, keypoint_distance=20.0)
            outs = []
            for s in range(steps):          # all steps are enqueued back to back: no host sync in between
                work = curr.clone(); kcnt = cnt.clone()
                status = torch.zeros((n, stride), dtype=torch.int8, device="cuda")
                ptrs = (C.c_void_p * n)(*[imgs[s + 1, i].data_ptr() for i in range(n)])
                prev, new = sets[s % 3], sets[(s + 1) % 3]
                N.check(L.rdfe_frontend_step_dev(fe.handle, prev.ctypes.data, new.ctypes.data, n, ptrs, 752, 6.0, 8, 8, C.byref(tp),
                                                 C.c_void_p(curr.data_ptr()), C.c_void_p(work.data_ptr()), C.c_void_p(cnt.data_ptr()),
                                                 C.c_void_p(status.data_ptr()), C.byref(dp), C.c_void_p(kcnt.data_ptr()), stride), "step")
                outs.append((work, kcnt, status, ptrs))
                curr, cnt = work, torch.clamp(kcnt, max=stride)     # next step tracks everything detected so far
            fe.sync()
            results[mode] = [(w.cpu().numpy(), k.cpu().numpy(), st.cpu().numpy()) for w, k, st, _ in outs]
            fe.close()
    for s in range(steps):
        a, b = results[0][s], results[1][s]
        assert np.array_equal(a[1], b[1]), f"step {s}: counts differ"
        assert np.array_equal(a[2], b[2]), f"step {s}: status differs"
        for i in range(n):
            assert np.array_equal(a[0][i, :a[1][i]], b[0][i, :b[1][i]]), f"step {s} image {i}"


@pytest.mark.parametrize("shape", [(480, 752), (257, 331)])
def test_undistort_bit_exact(orc, shape):
    """rdfe_set_undistort: cv::undistort's fixed-point remap in front of preprocess."""
    from rd_vio_b200.frontend import FrontEnd
    H, W = shape
    img = random_image(H, W, seed=3 * W)
    s = W / 752.0
    K = np.array([[458.654 * s, 0, 367.215 * s], [0, 457.296 * s, 248.375 * H / 480.0], [0, 0, 1]], np.float32)
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
    want = orc.undistort(img, K, D)
    with FrontEnd(W, H, max_level=2, win=21, num_slots=2, max_points=64) as fe:
        fe.set_undistort(K, D)
        s0 = fe.acquire()
        fe.preprocess([s0], [img])
        assert np.array_equal(fe.download_level(s0, 0, 3), want), "undistorted frame differs"
        assert np.array_equal(fe.download_level(s0, 0, 0), orc.clahe(want)), "CLAHE(undistort) differs"
        fe.set_undistort(None, None)
        fe.preprocess([s0], [img])
        assert np.array_equal(fe.download_level(s0, 0, 0), orc.clahe(img))


@pytest.mark.parametrize("channels,undistort", [(3, False), (4, False), (3, True), (4, True)])
def test_color_ingest_bit_exact(orc, channels, undistort):
    """rdfe_set_input_format: cvtColor(BGR/BGRA -> gray) of Odometry::addFrame on the device, optionally after the
    per-channel undistortion (reader order: cv::undistort on the loaded frame, then addFrame converts)."""
    from rd_vio_b200.frontend import FrontEnd
    H, W = 240, 320
    img = np.stack([random_image(H, W, seed=10 + c) for c in range(channels)], -1)
    s = W / 752.0
    K = np.array([[458.654 * s, 0, 367.215 * s], [0, 457.296 * s, 248.375 * H / 480.0], [0, 0, 1]], np.float32)
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
    gray = orc.bgr2gray(orc.undistort_color(img, K, D) if undistort else img)
    with FrontEnd(W, H, max_level=2, win=21, num_slots=2, max_points=64) as fe:
        fe.set_input_format(channels)
        if undistort:
            fe.set_undistort(K, D)
        s0 = fe.acquire()
        fe.preprocess([s0], [img])
        assert np.array_equal(fe.download_level(s0, 0, 3), gray), "ingest (gray) frame differs"
        assert np.array_equal(fe.download_level(s0, 0, 0), orc.clahe(gray))


@pytest.mark.parametrize("shape", [(478, 750), (300, 400), (100, 152), (241, 277), (130, 296), (97, 121), (64, 56)])
def test_harris_narrow_last_tile(orc, shape):
    """Widths whose last 120-column tile holds <= 32 / <= 56 columns: that tile is walked by 3 / 2 lane groups, one
    strip each (harris.cu), including bottom strips shorter than the others and groups without a strip.  Response map
    bit-exact in both float orders, corners identical."""
    from rd_vio_b200.frontend import FrontEnd
    H, W = shape
    img = random_image(H, W, seed=W * 7 + H)
    pre = orc.clahe(img)
    with FrontEnd(W, H, max_level=1 if min(H, W) > 90 else 0, win=21, num_slots=1, max_points=256) as fe:
        s = _pre(fe, img)
        for fma in (0, 1):
            got, ref = fe.harris_response(s, harris_fma=fma), orc.harris(pre, 0.04, mode=fma)
            assert np.array_equal(got, ref), f"fma={fma}: Harris map differs in {(got != ref).sum()} px"
        ref_kp, gxy_ref, gre_ref = orc.detect_keypoints(pre, np.zeros((0, 2)), 150, 20.0)
        kp, gxy, gre = fe.detect([s], [np.zeros((0, 2))], 150, 20.0, return_gftt=True)
        assert np.array_equal(gxy[0], gxy_ref) and np.array_equal(gre[0], gre_ref) and np.array_equal(kp[0], ref_kp)


def test_detect_full_height_strips_batch(orc):
    """48 frames of 752x480 in one call: the Harris strips have their full-batch height (adaptive_strip_rows) and the
    32-column last tile is walked three strips per warp; every frame's corners against the oracle."""
    from rd_vio_b200.frontend import FrontEnd
    n = 48
    imgs = [random_image(480, 752, 900 + i) for i in range(n)]
    with FrontEnd(752, 480, max_level=3, win=21, num_slots=n, max_points=256) as fe:
        slots = [fe.acquire() for _ in range(n)]
        fe.preprocess(slots, imgs)
        kps = fe.detect(slots, [np.zeros((0, 2))] * n, 150, 20.0)
        for i in range(n):
            ref = orc.detect_keypoints(orc.clahe(imgs[i]), np.zeros((0, 2)), 150, 20.0)[0]
            assert np.array_equal(kps[i], ref), f"frame {i}"
