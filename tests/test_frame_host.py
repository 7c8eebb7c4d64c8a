"""CPU: the host-side mirror of the plugin's callers (oracle/frame_host.py) against the
oracle's independent restatements and against properties the reference's formulas guarantee."""
import numpy as np
import pytest

from oracle import frame_host as F


def test_apply_remove_k_round_trip_and_formula(stream0):
    K = stream0.K()
    rng = np.random.default_rng(1)
    px = rng.uniform([0, 0], [752, 480], (200, 2))
    b = F.remove_k(px, K)
    assert np.allclose(np.linalg.norm(b, axis=1), 1.0, atol=1e-15)
    assert np.abs(F.apply_k(b, K) - px).max() < 1e-10
    # stereo.h:7-9 literally, one point
    p = b[7]
    assert F.apply_k(p, K)[0] == p[0] / p[2] * K[0, 0] + K[0, 2]
    assert F.apply_k(p, K)[1] == p[1] / p[2] * K[1, 1] + K[1, 2]


def test_quaternion_helpers_match_matrices():
    rng = np.random.default_rng(2)
    from rd_vio_b200.synthetic import _rot
    for _ in range(20):
        Ra, Rb = _rot(*rng.uniform(-3, 3, 3)), _rot(*rng.uniform(-3, 3, 3))
        qa, qb = F.q_from_matrix(Ra), F.q_from_matrix(Rb)
        v = rng.normal(size=(5, 3))
        assert np.allclose(F.q_rotate(qa, v), v @ Ra.T, atol=1e-12)
        assert np.allclose(F.q_rotate(F.q_mul(qa, qb), v), v @ (Ra @ Rb).T, atol=1e-12)
        assert np.allclose(F.q_rotate(F.q_conj(qa), v), v @ Ra, atol=1e-12)


def test_rotation_prediction_equals_true_flow_for_distant_points(stream0):
    """frame.cpp:82-93: the prediction rotates bearings by the gyro increment; for points on the far plane of the
    synthetic scene (8 m) it must be within the parallax of the true flow, and exact for a pure rotation."""
    K = stream0.K()
    a = F.Frame(None, K)
    a.bearings = F.remove_k(np.array([[100.0, 100.0], [376.0, 240.0], [700.0, 400.0]]), K)
    b = F.Frame(None, K, delta_q=F.q_from_matrix(stream0.gyro_delta(3)))
    pred = F.apply_k(F.q_rotate(a.predicted_rotation(b), a.bearings), K)
    R0, _ = stream0.pose(3)
    R1, _ = stream0.pose(4)
    rays = a.bearings @ R0.T @ R1                       # world ray expressed in camera 4
    assert np.abs(pred - F.apply_k(rays, K)).max() < 1e-9
    true = stream0.flow(3, a.keypoints())
    assert np.abs(pred - true).max() < 12.0              # translation parallax only


@pytest.mark.parametrize("radius", [20.0, 10.0, 45.5])
def test_poisson_filter_matches_the_oracle_restatement(radius):
    from oracle import fe_oracle as orc
    rng = np.random.default_rng(int(radius))
    for trial in range(20):
        ex = rng.uniform([0, 0], [752, 480], (rng.integers(0, 120), 2))
        ca = np.rint(rng.uniform([0, 0], [752, 480], (300, 2)))
        f = F.PoissonDiskFilter(radius)
        f.preset_points(ex)
        got = f.insert_points(ca)
        assert np.array_equal(got, orc.poisson_filter(ex, ca, radius)), trial


def test_poisson_filter_insert_only_equals_brute_force():
    """poisson_disk_filter.h:80-92 skips cell (ix-2, iy-2) and visits (ix-2, iy+3) instead; the skipped cell is at
    least r away, so for insert-only use (one point per cell) the filter equals the brute-force distance test."""
    rng = np.random.default_rng(5)
    f, kept = F.PoissonDiskFilter(20.0), []
    for p in rng.uniform([0, 0], [300, 300], (3000, 2)):
        brute = all((p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 >= 400.0 for q in kept)
        assert f.insert_point(p) == brute
        if brute:
            kept.append(p)
    # preset_point overwrites the single slot of a cell (:23-27): the earlier point of that cell is forgotten
    f = F.PoissonDiskFilter(20.0)
    f.preset_point((100.0, 100.0))
    f.preset_point((112.0, 112.0))                       # same cell (cell size 14.14: both in cell 7)
    assert f.permit_point((90.0, 90.0))                  # 14.1 px from the forgotten point, 31 px from the kept one


class _FakeImage:
    """Plugin stub: tracks every point by a fixed shift, fails the points listed in `fail`; detects nothing new."""

    def __init__(self, shift=(1.0, 0.0), fail=(), new=None):
        self.shift, self.fail, self.new, self.calls = np.asarray(shift), set(fail), new, []

    def preprocess(self, *a):
        self.calls.append("preprocess")

    def detect_keypoints(self, kp, max_points, dist):
        self.calls.append("detect")
        kp = np.asarray(kp, np.float64).reshape(-1, 2)
        return kp if self.new is None else np.vstack([kp, self.new])

    def track_keypoints(self, nxt_img, curr, pred=None):
        self.calls.append("track")
        st = np.ones(len(curr), np.int8)
        out = np.asarray(pred if pred is not None else curr, np.float64).copy()
        for i in range(len(curr)):
            if i in self.fail:
                st[i] = 0
            else:
                out[i] = curr[i] + self.shift
        return out, st

    def release_image_buffer(self):
        self.calls.append("release")


def test_feature_tracker_call_order_and_track_bookkeeping(stream0):
    K = stream0.K()
    new0 = np.array([[100.0, 100.0], [300.0, 200.0], [305.0, 203.0], [500.0, 300.0]])
    ft = F.FeatureTracker()
    im0, im1, im2 = _FakeImage(new=new0), _FakeImage(fail={1}), _FakeImage()
    f0 = ft.track_frame(F.Frame(im0, K, 0))
    assert im0.calls == ["preprocess", "detect"] and f0.keypoint_num() == 4 and f0.tracks == [None] * 4
    f1 = ft.track_frame(F.Frame(im1, K, 1))
    # feature_tracker.cpp:32-98: preprocess(new), last.track(new), last.release, new.detect
    assert im1.calls == ["preprocess", "detect"] and im0.calls == ["preprocess", "detect", "track", "release"]
    # no tracks existed yet -> nothing is Poisson-filtered (frame.cpp:140-144 skips track == nullptr): 4 survive
    assert f1.keypoint_num() == 4 and [t.keypoint_num for t in f1.tracks] == [2, 2, 2, 2]
    assert np.allclose(f1.keypoints(), new0 + [1.0, 0.0], atol=1e-9)
    f2 = ft.track_frame(F.Frame(im2, K, 2))
    # im1 fails point 1; points 1 and 2 of f1 are 5.8 px apart: with point 1 gone point 2 survives the filter
    assert f2.keypoint_num() == 3 and [t.id for t in f2.tracks] == [0, 2, 3]
    assert [t.keypoint_num for t in f2.tracks] == [3, 3, 3]


def test_track_length_ordered_poisson_filter(stream0):
    """frame.cpp:134-158: of two tracked points closer than min_keypoint_distance the longer track wins."""
    K = stream0.K()
    alloc = F.TrackAllocator()
    a, b = F.Frame(_FakeImage(), K, 0), F.Frame(_FakeImage(), K, 1)
    a.bearings = F.remove_k(np.array([[200.0, 200.0], [205.0, 200.0], [400.0, 300.0]]), K)
    a.tracks = [alloc.create_track() for _ in range(3)]
    for t, n in zip(a.tracks, (2, 9, 1)):
        t.keypoint_num = n
    st = a.track_keypoints(b, alloc)
    assert st.tolist() == [0, 1, 1] and [t.id for t in b.tracks] == [1, 2]
    # a TT_TRASH track is dropped even when the filter permits it
    a2, b2 = F.Frame(_FakeImage(), K, 0), F.Frame(_FakeImage(), K, 1)
    a2.bearings, a2.tracks = a.bearings.copy(), list(a.tracks)
    a.tracks[2].trash = True
    assert a2.track_keypoints(b2, alloc).tolist() == [0, 1, 0]


def test_oracle_replay_with_imu_prediction_keeps_tracks(stream0):
    """FeatureTracker over 5 synthetic frames on the CPU oracle: tracks persist and grow."""
    from frame_helpers import OracleImage, replay
    out = replay(stream0, 5, lambda im, t: OracleImage(im, t))
    n0 = len(out[0][0])
    assert 100 <= n0 <= 150
    ids_last = set(out[-1][1]) - {-1}
    assert len(ids_last) >= 0.6 * n0                     # most of the first frame's corners are still tracked
    for kp, ids in out[1:]:
        tracked = [i for i in ids if i >= 0]
        assert len(tracked) == len(set(tracked))
        d = kp[:, None, :] - kp[None, :, :]
        dist = np.sqrt((d ** 2).sum(-1)) + np.eye(len(kp)) * 1e9
        assert dist.min() >= 20.0 - 1e-9                 # both Poisson filters held (radius 20)
