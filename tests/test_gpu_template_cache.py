"""GPU: the LK template cache (rdfe_set_template_cache) must never change a result.  The backward pass of track(A -> B)
leaves its per-level templates at the tracked positions in B; track(B -> C) on the carried points loads them instead of
rebuilding.  Reference semantics: OpenCvImage::track_keypoints (opencv_image.cpp:75-154) called frame after frame by
Frame::track_keypoints (frame.cpp:74-172) with next frame's curr = this frame's tracked points."""
import numpy as np
import pytest

from conftest import random_image

pytestmark = pytest.mark.gpu


def _chain(fe, frames, n_detect, cache):
    """detect on frame 0, then track 0->1->2->... carrying the survivors; returns every (next, status) pair."""
    fe.set_template_cache(cache)
    slots = [fe.acquire() for _ in frames]
    out = []
    try:
        fe.preprocess(slots, frames)
        pts = fe.detect([slots[0]], [np.zeros((0, 2))], n_detect, 20.0)[0]
        for a, b in zip(slots[:-1], slots[1:]):
            nxt, st = fe.track([a], [b], [pts], None)
            out.append((nxt[0].copy(), st[0].copy()))
            pts = nxt[0][st[0] != 0]
    finally:
        for s in slots:
            fe.release(s)
    return out


@pytest.mark.parametrize("cfg", [(752, 480, 3, 21, 150), (640, 480, 3, 31, 120)])
def test_cache_is_transparent_and_hits(cfg):
    from rd_vio_b200.frontend import FrontEnd
    from rd_vio_b200.synthetic import SyntheticStream
    W, H, lv, win, npts = cfg
    st = SyntheticStream(3, W, H)
    frames = [st.frame(k) for k in range(4)]
    with FrontEnd(W, H, lv, win, num_slots=4, max_points=512) as fe:
        ref = _chain(fe, frames, npts, cache=False)
        assert fe.template_cache_stats() == (0, 0)
        got = _chain(fe, frames, npts, cache=True)
        lookups, hits = fe.template_cache_stats(reset=True)
    for (rn, rs), (gn, gs) in zip(ref, got):
        assert np.array_equal(rs, gs)
        assert np.array_equal(rn, gn)           # bit-identical positions, also where status == 0 (left untouched)
    carried = sum(int((s != 0).sum()) for _, s in got[:-1])
    assert lookups == sum(len(s) for _, s in got)
    # every carried point finds the template its backward pass left (first step: detections, nothing cached)
    assert hits == carried, (lookups, hits, carried)


def test_cache_invalidated_by_new_pixels():
    """Rewriting a slot's image must not let an old template through: same positions, different pixels."""
    from rd_vio_b200.frontend import FrontEnd
    imgs = [random_image(240, 320, s) for s in (11, 12, 13, 14)]
    with FrontEnd(320, 240, 3, 21, num_slots=3, max_points=256) as fe:
        fe.set_template_cache(True)
        a, b, c = fe.acquire(), fe.acquire(), fe.acquire()
        fe.preprocess([a, b, c], [imgs[0], imgs[0], imgs[0]])
        pts = fe.detect([a], [np.zeros((0, 2))], 100, 20.0)[0]
        n1, s1 = fe.track([a], [b], [pts], None)          # leaves templates for slot b at n1
        carried = n1[0][s1[0] != 0]
        fe.preprocess([b], [imgs[1]])                       # new pixels in slot b: records are stale
        fe.template_cache_stats(reset=True)
        got = fe.track([b], [c], [carried], None)
        lookups, hits = fe.template_cache_stats(reset=True)
        assert lookups == len(carried) and hits == 0
        fe.set_template_cache(False)
        want = fe.track([b], [c], [carried], None)
        assert np.array_equal(got[1][0], want[1][0]) and np.array_equal(got[0][0], want[0][0])
