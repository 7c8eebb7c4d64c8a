"""TEST INFRASTRUCTURE ONLY (lives beside the oracle; the product package never imports it).
Host-side mirror of the callers of the Image plugin:

  Frame::detect_keypoints / Frame::track_keypoints   /root/reference/src/rdvio_map/src/frame.cpp:55-172
  apply_k / remove_k                                  src/rdvio_geometry/include/rdvio/geometry/stereo.h:7-14
  extra::PoissonDiskFilter<2>                         src/rdvio_extra/include/rdvio/extra/poisson_disk_filter.h:8-113
  FeatureTracker::run (plugin call order)             src/rdvio/src/feature_tracker.cpp:26-111

In a C++ deployment these functions are the reference's own and stay untouched (the drop-in replaces only the
class behind `rdvio::Image`); this module exists so that replays, tests and benchmarks can drive the plugin
with exactly the inputs the reference's FeatureTracker would hand it: pixel positions obtained from unit
bearings through K, predictions obtained by rotating the bearings with the pre-integrated gyro rotation, and
survivors thinned by the track-length-ordered Poisson filter.  Everything here is float64 host arithmetic on a
few hundred points per frame; the pixels never come back to the host.

Not mirrored (host back-end, out of scope): the 5-point essential-matrix RANSAC and the
2-point rotation RANSAC of frame.cpp:106-132 (Eigen SVD solvers).  They enter through the optional
`geometric_check` hook, whose default accepts every point.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np


# ---------------------------------------------------------------- stereo.h:7-14
def apply_k(p, K):
    """bearing(s) (...,3) -> pixel(s) (...,2): p0/p2*K00+K02, p1/p2*K11+K12 (same operation order)."""
    p = np.asarray(p, np.float64)
    K = np.asarray(K, np.float64)
    return np.stack([p[..., 0] / p[..., 2] * K[0, 0] + K[0, 2], p[..., 1] / p[..., 2] * K[1, 1] + K[1, 2]], -1)


def remove_k(p, K):
    """pixel(s) (...,2) -> unit bearing(s) (...,3)."""
    p = np.asarray(p, np.float64)
    K = np.asarray(K, np.float64)
    v = np.stack([(p[..., 0] - K[0, 2]) / K[0, 0], (p[..., 1] - K[1, 2]) / K[1, 1], np.ones(p.shape[:-1])], -1)
    return v / np.sqrt((v * v).sum(-1, keepdims=True))


# ---------------------------------------------------------------- quaternions (w, x, y, z), Hamilton product
def q_mul(a, b):
    aw, ax, ay, az = a
    bw, bx, by, bz = b
    return np.array([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw])


def q_conj(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def q_rotate(q, v):
    """q * v for unit q (Eigen's QuaternionBase::_transformVector: v + w*uv + qv x uv, uv = 2 qv x v)."""
    v = np.asarray(v, np.float64)
    qv = np.asarray(q[1:], np.float64)
    uv = 2.0 * np.cross(qv, v)
    return v + q[0] * uv + np.cross(qv, uv)


def q_from_matrix(R):
    """Unit quaternion of a rotation matrix (w >= 0 branch first, Shepperd's method)."""
    R = np.asarray(R, np.float64)
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


_Q_ID = np.array([1.0, 0.0, 0.0, 0.0])


# ---------------------------------------------------------------- poisson_disk_filter.h:8-113
class PoissonDiskFilter:
    """extra::PoissonDiskFilter<2>: hash grid with ONE slot per cell (a later point in the same cell overwrites
    the slot, :23-27), cell = r/sqrt(2), span 2; test_point walks the 5x5 block in the reference's literal order,
    which skips the first cell (ix-2, iy-2) and visits (ix-2, iy+3) instead (:80-92); reject iff d^2 < r^2."""

    def __init__(self, radius: float):
        self.radius = float(radius)
        self.radius_squared = self.radius * self.radius
        self.grid_size = self.radius / math.sqrt(2.0)
        self.grid_span = int(math.ceil(math.sqrt(2.0)))
        self.points: List[tuple] = []
        self.sparse_grid: Dict[tuple, int] = {}

    def clear(self):
        self.points.clear()
        self.sparse_grid.clear()

    def _to_index(self, p):
        return (int(math.floor(p[0] / self.grid_size)), int(math.floor(p[1] / self.grid_size)))

    def preset_point(self, p):
        self.sparse_grid[self._to_index(p)] = len(self.points)
        self.points.append((float(p[0]), float(p[1])))

    def preset_points(self, pts):
        for p in pts:
            self.preset_point(p)

    def _test_point(self, p):
        index = self._to_index(p)
        b0, b1 = index[0] - self.grid_span, index[1] - self.grid_span
        e0, e1 = index[0] + self.grid_span, index[1] + self.grid_span
        c0, c1 = b0, b1
        while c1 <= e1:
            c0 += 1
            if c0 > e0:
                c0 = b0
                c1 += 1
            k = self.sparse_grid.get((c0, c1))
            if k is not None:
                q = self.points[k]
                dx, dy = p[0] - q[0], p[1] - q[1]
                if dx * dx + dy * dy < self.radius_squared:
                    return False, index
        return True, index

    def permit_point(self, p) -> bool:
        return self._test_point(p)[0]

    def insert_point(self, p) -> bool:
        ok, index = self._test_point(p)
        if ok:
            self.sparse_grid[index] = len(self.points)
            self.points.append((float(p[0]), float(p[1])))
        return ok

    def insert_points(self, candidates):
        """Returns the accepted candidates, in order (the reference swaps them into its argument)."""
        n0 = len(self.points)
        for p in candidates:
            self.insert_point(p)
        return np.array(self.points[n0:], np.float64).reshape(-1, 2)


# ---------------------------------------------------------------- config.cpp:23-37, 57-59
class Config:
    feature_tracker_min_keypoint_distance = 20.0
    feature_tracker_max_keypoint_detection = 150
    feature_tracker_clahe_clip_limit = 6.0
    feature_tracker_clahe_width = 8
    feature_tracker_clahe_height = 8
    feature_tracker_predict_keypoints = True
    sliding_window_tracker_frequent = 1


class Track:
    """The two members of rdvio::Track this path reads: keypoint_num() and tag(TT_TRASH)."""
    __slots__ = ("id", "keypoint_num", "trash")

    def __init__(self, tid):
        self.id, self.keypoint_num, self.trash = tid, 0, False


class TrackAllocator:
    """Map::create_track as far as the front-end needs it."""

    def __init__(self):
        self.tracks: List[Track] = []

    def create_track(self) -> Track:
        t = Track(len(self.tracks))
        self.tracks.append(t)
        return t


class Frame:
    """rdvio::Frame restricted to what detect_keypoints / track_keypoints touch: `image` (anything with the
    Image plugin's methods), K, unit `bearings`, `tracks`, the camera / IMU extrinsic rotations `camera_q_cs`,
    `imu_q_cs`, and `delta_q` = preintegration.delta.q, the gyro rotation from the previous frame to this one."""

    def __init__(self, image, K, frame_id=0, delta_q=None, camera_q_cs=None, imu_q_cs=None):
        self.image = image
        self.K = np.asarray(K, np.float64).reshape(3, 3)
        self.id = frame_id
        self.delta_q = _Q_ID if delta_q is None else np.asarray(delta_q, np.float64)
        self.camera_q_cs = _Q_ID if camera_q_cs is None else np.asarray(camera_q_cs, np.float64)
        self.imu_q_cs = _Q_ID if imu_q_cs is None else np.asarray(imu_q_cs, np.float64)
        self.bearings = np.zeros((0, 3), np.float64)
        self.tracks: List[Optional[Track]] = []
        self.no_translation = False          # tag(FT_NO_TRANSLATION); only the geometric_check hook sets it

    def keypoint_num(self):
        return len(self.bearings)

    def keypoints(self):
        return apply_k(self.bearings, self.K) if len(self.bearings) else np.zeros((0, 2))

    def append_keypoint(self, bearing):                                   # frame.cpp:37-41
        self.bearings = np.vstack([self.bearings, np.asarray(bearing, np.float64).reshape(1, 3)])
        self.tracks.append(None)

    def get_track(self, i, allocator: TrackAllocator) -> Track:           # frame.cpp:43-53
        if self.tracks[i] is None:
            t = allocator.create_track()
            t.keypoint_num += 1
            self.tracks[i] = t
        return self.tracks[i]

    # ------------------------------------------------------------ frame.cpp:55-72
    def detect_keypoints(self, config=Config):
        pk = self.keypoints()
        out = self.image.detect_keypoints(pk, config.feature_tracker_max_keypoint_detection,
                                          config.feature_tracker_min_keypoint_distance)
        out = np.asarray(out, np.float64).reshape(-1, 2)
        old = len(self.bearings)
        if len(out) > old:
            self.bearings = np.vstack([self.bearings, remove_k(out[old:], self.K)])
            self.tracks.extend([None] * (len(out) - old))

    # ------------------------------------------------------------ frame.cpp:74-172
    def predicted_rotation(self, next_frame):
        """delta_key_q of frame.cpp:82-87: rotates a bearing of this camera into the next camera."""
        q = q_mul(q_conj(self.camera_q_cs), self.imu_q_cs)
        q = q_mul(q, next_frame.delta_q)
        q = q_mul(q, q_conj(next_frame.imu_q_cs))
        q = q_mul(q, next_frame.camera_q_cs)
        return q_conj(q)

    def track_keypoints(self, next_frame: "Frame", allocator: TrackAllocator, config=Config,
                        geometric_check: Optional[Callable] = None):
        n = len(self.bearings)
        curr = self.keypoints()
        pred = None
        if config.feature_tracker_predict_keypoints and n:
            dq = self.predicted_rotation(next_frame)
            pred = apply_k(q_rotate(dq, self.bearings), next_frame.K)
        nxt, status = self.image.track_keypoints(next_frame.image, curr, pred)
        nxt = np.asarray(nxt, np.float64).reshape(-1, 2)
        status = np.array(status, np.int8).reshape(-1).copy()
        next_bearings = remove_k(nxt, next_frame.K) if n else np.zeros((0, 3))

        if geometric_check is not None:     # find_essential_matrix / find_rotation_matrix (:106-132), host back-end
            mask = np.asarray(geometric_check(self, next_frame, self.bearings, next_bearings), bool)
            status[~mask] = 0

        # filter keypoints based on track length (:134-158); std::sort leaves the order of equal lengths
        # unspecified in the reference, here it is the stable order (keypoint index ascending)
        order = [(i, self.tracks[i].keypoint_num) for i in range(n) if status[i] and self.tracks[i] is not None]
        order.sort(key=lambda a: -a[1])
        filt = PoissonDiskFilter(config.feature_tracker_min_keypoint_distance)
        for i, _ in order:
            if filt.permit_point(nxt[i]) and not self.tracks[i].trash:
                filt.preset_point(nxt[i])
            else:
                status[i] = 0

        for i in range(n):                                                 # :160-171
            if status[i]:
                j = next_frame.keypoint_num()
                next_frame.append_keypoint(next_bearings[i])
                t = self.get_track(i, allocator)
                next_frame.tracks[j] = t
                t.keypoint_num += 1
        return status


class FeatureTracker:
    """The plugin-facing part of FeatureTracker::run (feature_tracker.cpp:26-111): preprocess(new) ->
    last.track_keypoints(new) -> last.image.release_image_buffer() -> new.detect_keypoints()."""

    def __init__(self, config=Config, geometric_check=None):
        self.config = config
        self.allocator = TrackAllocator()
        self.last: Optional[Frame] = None
        self.geometric_check = geometric_check

    def track_frame(self, frame: Frame) -> Frame:
        c = self.config
        frame.image.preprocess(c.feature_tracker_clahe_clip_limit, c.feature_tracker_clahe_width,
                               c.feature_tracker_clahe_height)
        if self.last is not None:
            self.last.track_keypoints(frame, self.allocator, c, self.geometric_check)
            self.last.image.release_image_buffer()
        if frame.id % c.sliding_window_tracker_frequent == 0:
            frame.detect_keypoints(c)
        self.last = frame
        return frame
