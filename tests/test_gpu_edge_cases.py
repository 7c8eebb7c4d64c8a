"""GPU edge cases the reference's call sites can produce: empty inputs, flat images (no corners, LK failures),
full keypoint buffers, maximum batch, invalid arguments (error codes, no crashes)."""
import ctypes as C

import numpy as np
import pytest

from conftest import random_image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orc():
    from oracle import fe_oracle
    return fe_oracle


@pytest.fixture(scope="module")
def fe():
    from rd_vio_b200.frontend import FrontEnd
    f = FrontEnd(320, 240, max_level=3, win=21, num_slots=130, max_points=400)
    yield f
    f.close()


def test_flat_and_saturated_images(fe, orc):
    """Constant image: Harris response is 0 everywhere -> no corner; LK: minEig below threshold -> status 0."""
    flat = np.full((240, 320), 77, np.uint8)
    sat = random_image(240, 320, 5)
    sat[60:180, 80:240] = 255                      # saturated block
    a, b, c = fe.acquire(), fe.acquire(), fe.acquire()
    try:
        fe.preprocess([a, b, c], [flat, flat, sat])
        out = fe.detect([a], [np.zeros((0, 2))], 150, 20.0)[0]
        assert len(out) == 0 == len(orc.detect_keypoints(orc.clahe(flat), np.zeros((0, 2)), 150, 20.0)[0])
        pts = np.array([[100.0, 100.0], [200.5, 120.25]])
        nxt, st = fe.track([a], [b], [pts], None)
        assert st[0].sum() == 0
        got = fe.detect([c], [np.zeros((0, 2))], 150, 20.0)[0]
        assert np.array_equal(got, orc.detect_keypoints(orc.clahe(sat), np.zeros((0, 2)), 150, 20.0)[0])
    finally:
        for s in (a, b, c):
            fe.release(s)


def test_empty_and_full_keypoint_buffers(fe, orc):
    img = random_image(240, 320, 9)
    a, b = fe.acquire(), fe.acquire()
    try:
        fe.preprocess([a, b], [img, img])
        # zero keypoints to track: nothing happens, nothing crashes
        nxt, st = fe.track([a], [b], [np.zeros((0, 2))], None)
        assert len(nxt[0]) == 0 and len(st[0]) == 0
        # identical frames: every good point tracks onto itself
        kp = fe.detect([a], [np.zeros((0, 2))], 150, 20.0)[0]
        nxt, st = fe.track([a], [b], [kp], None)
        ok = st[0] != 0
        assert ok.sum() >= 0.9 * len(kp) and np.abs(nxt[0][ok] - kp[ok]).max() < 1e-3
        # The reference's keypoint vector is unbounded: a buffer too small for existing + new corners is an error
        # (RDFE_ERR_OVERFLOW, "keypoint list truncated"), never a silently shorter list ...
        from rd_vio_b200._native import FrontEndError
        full = kp[:10]
        ref = orc.detect_keypoints(orc.clahe(img), full, 150, 20.0)[0]
        assert len(ref) > 13
        for too_small in (10, 13):
            with pytest.raises(FrontEndError, match="truncated"):
                fe.detect([a], [full], 150, 20.0, stride=too_small)
        # ... the flag is cleared by the failing call, and a buffer that fits exactly is fine
        out = fe.detect([a], [full], 150, 20.0, stride=len(ref))[0]
        assert np.array_equal(out, ref)
    finally:
        fe.release(a)
        fe.release(b)


def test_maximum_batch(fe, orc):
    """128 images in one call (RDFE_MAX_BATCH), each different; spot-check a few against the oracle."""
    n = 128
    imgs = [random_image(240, 320, 100 + i) for i in range(n)]
    slots = [fe.acquire() for _ in range(n)]
    try:
        fe.preprocess(slots, imgs)
        kps = fe.detect(slots, [np.zeros((0, 2))] * n, 100, 20.0)
        for i in (0, 63, 127):
            assert np.array_equal(fe.download_level(slots[i], 0, 0), orc.clahe(imgs[i]))
            assert np.array_equal(kps[i], orc.detect_keypoints(orc.clahe(imgs[i]), np.zeros((0, 2)), 100, 20.0)[0])
    finally:
        for s in slots:
            fe.release(s)


def test_invalid_arguments_return_errors(fe):
    from rd_vio_b200 import _native as N
    L = N.lib()
    img = random_image(240, 320, 1)
    with pytest.raises(N.FrontEndError):                     # slot never acquired
        fe.preprocess([129], [img])
    a = fe.acquire()
    try:
        with pytest.raises(ValueError):                      # wrong image size caught on the host side
            fe.preprocess([a], [img[:100]])
        fe.preprocess([a], [img])
        with pytest.raises(N.FrontEndError):                 # max_points above the context capacity
            fe.detect([a], [np.zeros((0, 2))], 100000, 20.0, stride=10)
        sl = np.array([a], np.int32)
        ptr = (C.c_void_p * 1)(img.ctypes.data)
        assert L.rdfe_preprocess_batch(fe.handle, sl.ctypes.data, 0, ptr, 320, 6.0, 8, 8) < 0      # n = 0
        assert L.rdfe_preprocess_batch(fe.handle, sl.ctypes.data, 1, ptr, 100, 6.0, 8, 8) < 0      # pitch < width
        assert L.rdfe_preprocess_batch(fe.handle, sl.ctypes.data, 1, ptr, 320, 6.0, 0, 8) < 0      # zero tiles
        assert L.rdfe_last_error()
        # the context is still usable afterwards
        fe.preprocess([a], [img])
    finally:
        fe.release(a)
    with pytest.raises(N.FrontEndError):                     # double release
        fe.release(a)


def test_unsupported_window_is_refused():
    from rd_vio_b200 import _native as N
    from rd_vio_b200.frontend import FrontEnd
    with pytest.raises(N.FrontEndError):
        FrontEnd(320, 240, max_level=3, win=15)


def test_random_odd_shapes_preprocess_bit_exact(orc):
    """Sizes that are not multiples of the tile grid / of 2 / of 4: CLAHE padding quirk, ceil halving, REFLECT_101
    at odd borders, partial 4-pixel groups, byte-granular fallbacks."""
    from rd_vio_b200.frontend import FrontEnd
    rng = np.random.default_rng(2024)
    for _ in range(10):
        H, W = int(rng.integers(60, 300)), int(rng.integers(70, 400))
        tiles = (int(rng.integers(2, 9)), int(rng.integers(2, 9)))
        img = random_image(H, W, seed=H * 1000 + W)
        ref = orc.clahe(img, 6.0, tiles[0], tiles[1])
        P = orc.Pyramid(ref, 21, 3)
        with FrontEnd(W, H, max_level=3, win=21, num_slots=1, max_points=64) as f:
            s = f.acquire()
            f.preprocess([s], [img], tiles=tiles)
            assert f.nlevels == P.nlevels, (H, W)
            for l in range(P.nlevels):
                assert np.array_equal(f.download_level(s, l, 0), P.image(l)), (H, W, tiles, l)
                assert np.array_equal(f.download_level(s, l, 1), P.deriv(l)), (H, W, tiles, l)
                assert np.array_equal(f.download_level(s, l, 2), np.pad(P.image(l), 21, mode="reflect")), (H, W, l)
            R = f.harris_response(s)
            assert np.array_equal(R, orc.harris(ref)), (H, W)


def test_contexts_on_two_devices(orc):
    """The > 48 KB dynamic shared memory opt-in is a per-device function attribute: a second context on another GPU of the
    same process must get its own (round-1 bug: a process-wide cache skipped it).  1080p selection needs ~133 KB."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from rd_vio_b200.frontend import FrontEnd
    img = random_image(1080, 1920, 77)
    ref = orc.detect_keypoints(orc.clahe(img), np.zeros((0, 2)), 400, 20.0)[0]
    for dev in (0, 1):
        with FrontEnd(1920, 1080, max_level=3, win=21, num_slots=2, max_points=1024, device=dev) as f:
            s = f.acquire()
            f.preprocess([s], [img])
            got = f.detect([s], [np.zeros((0, 2))], 400, 20.0)[0]
            assert np.array_equal(got, ref), f"device {dev}"
