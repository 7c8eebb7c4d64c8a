"""CPU: size-independent properties of the path, checked on the oracle (the GPU tests compare the CUDA path with the
oracle bit for bit, so what holds here holds there): identities of the integer stages, LK on an image against
itself and against an integer translation of itself, and the geometric guarantees of detect_keypoints."""
import numpy as np
import pytest

from conftest import random_image
from oracle import fe_oracle as orc


def test_integer_stages_on_constant_and_ramp_images():
    c = np.full((97, 131), 77, np.uint8)
    assert np.array_equal(orc.pyrdown(c), np.full((49, 66), 77, np.uint8))       # kernel sums to 256, (v*256+128)>>8 = v
    assert not orc.scharr(c).any()
    ramp = np.tile(np.arange(131, dtype=np.uint8), (97, 1))                          # I(x, y) = x
    d = orc.scharr(ramp)
    assert (d[:, 1:-1, 0] == 32).all() and not d[..., 1].any()                       # (3 + 10 + 3) * (I[x+1] - I[x-1])
    assert not d[:, 0, 0].any() and not d[:, -1, 0].any()                            # REFLECT_101: I[-1] = I[1]
    flat = orc.clahe(c)                                                              # one grey level: every LUT maps it to 255
    assert (flat == flat[0, 0]).all()


@pytest.mark.parametrize("shape,win,max_level", [((480, 752), 21, 3), ((360, 640), 31, 3)])
def test_lk_identity_and_integer_translation(shape, win, max_level):
    H, W = shape
    big = orc.clahe(random_image(H + 32, W + 32, seed=H))
    a = np.ascontiguousarray(big[16:16 + H, 16:16 + W])
    PA = orc.Pyramid(a, win, max_level)
    pts = orc.detect_keypoints(a, np.zeros((0, 2)), 150, 20.0)[0]
    inner = pts[(pts[:, 0] > 60) & (pts[:, 0] < W - 60) & (pts[:, 1] > 60) & (pts[:, 1] < H - 60)]
    assert len(inner) > 40
    # an image tracked onto itself: the first iteration finds b = 0 and stops; positions are returned unchanged
    nxt, st, _ = orc.track_keypoints(PA, PA, inner, None, win, max_level)
    assert st.all() and np.abs(nxt - inner).max() <= 1e-3
    # the same scene shifted by whole pixels: level 0 sees an exact translation, the result is the shift
    for dx, dy in ((3, -2), (-5, 4)):
        b = np.ascontiguousarray(big[16 - dy:16 - dy + H, 16 - dx:16 - dx + W])      # b(x, y) = a(x - dx, y - dy)
        PB = orc.Pyramid(b, win, max_level)
        nxt, st, _ = orc.track_keypoints(PA, PB, inner, None, win, max_level)
        assert st.mean() >= 0.95
        err = np.abs(nxt[st != 0] - (inner[st != 0] + [dx, dy]))
        assert np.median(err) <= 0.02 and err.max() <= 0.25
        # forward-backward consistency is what status certifies (opencv_image.cpp:127-134)
        back, st2, _ = orc.track_keypoints(PB, PA, nxt[st != 0], None, win, max_level)
        assert np.abs(back[st2 != 0] - inner[st != 0][st2 != 0]).max() <= 0.5


@pytest.mark.parametrize("n_existing,radius,max_points", [(0, 20.0, 150), (60, 20.0, 150), (60, 35.0, 80), (200, 10.0, 300)])
def test_detect_geometric_guarantees(n_existing, radius, max_points):
    H, W = 480, 752
    img = orc.clahe(random_image(H, W, seed=n_existing + 1))
    rng = np.random.default_rng(n_existing)
    ex = rng.uniform([0, 0], [W, H], (n_existing, 2))
    out, gxy, gre = orc.detect_keypoints(img, ex, max_points, radius)
    new = out[n_existing:]
    assert np.array_equal(out[:n_existing], ex)                                      # existing keypoints untouched, in place
    assert len(gxy) <= max_points and np.all(np.diff(gre) <= 0)                      # GFTT order: response descending
    if len(gxy) > 1:                                                                 # GFTT's own minDistance 20
        d = np.sqrt(((gxy[:, None, :] - gxy[None, :, :]) ** 2).sum(-1)) + np.eye(len(gxy)) * 1e9
        assert d.min() >= 20.0
    assert np.all(new == np.rint(new))                                               # integer pixel coordinates
    assert np.all((new[:, 0] >= 20) & (new[:, 0] < W - 20) & (new[:, 1] >= 20) & (new[:, 1] < H - 20))   # :61-68
    # the Poisson filter sees one point per cell (later presets hide earlier ones), so the guarantee is against the
    # visible existing points and against each other
    g = radius / np.sqrt(2.0)
    cells = np.floor(ex / g).astype(int)
    visible = [i for i in range(n_existing) if not any((cells[j] == cells[i]).all() for j in range(i + 1, n_existing))]
    ref = np.concatenate([ex[visible], new], 0) if len(new) else ex[visible]
    for p in new:
        d = np.sqrt(((ref - p) ** 2).sum(-1))
        assert np.sort(d)[1] >= radius - 1e-9 if len(ref) > 1 else True              # d[0] = the point itself
