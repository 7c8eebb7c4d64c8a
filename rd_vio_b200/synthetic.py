"""EuRoC-shaped synthetic camera streams ("Synthetic inputs").

A stream is a pinhole camera (EuRoC cam0 intrinsics, /root/reference/configs/euroc_sensor.yaml:43,
scaled for other resolutions) moving smoothly in front of three fronto-parallel textured planes at
depths 2 / 4 / 8 m (parallax), with per-frame gain drift (+-10 %) and additive Gaussian noise
(sigma = 2 grey levels) so that CLAHE has something to do.  The trajectory is periodic with period
`period` frames, so a ring of `period` frames can be replayed forever with small inter-frame motion
(also across the wrap-around).  Frames are "already undistorted", as the reference's dataset reader
undistorts before the Image plugin sees pixels (/root/reference/examples/dataset.hpp:591).

Pure numpy + scipy (no cv2) so it runs anywhere the package runs.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

_EUROC_K = (458.654, 457.296, 367.215, 248.375)   # fx fy cx cy at 752x480


def _texture(rng, size):
    """Multi-octave band-limited noise: blurred uniform noise at sigma 1.5, 3, 6, 12 px, normalised to 0..255."""
    acc = np.zeros((size, size), np.float32)
    for sigma, wgt in ((1.5, 1.0), (3.0, 1.0), (6.0, 1.0), (12.0, 1.0)):
        n = rng.random((size, size), dtype=np.float32) - 0.5
        b = ndimage.gaussian_filter(n, sigma, mode="wrap")
        acc += wgt * b / (b.std() + 1e-12)
    acc -= acc.min()
    acc *= 255.0 / acc.max()
    return acc


def _rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


class SyntheticStream:
    """One independent camera stream. frame(k) -> HxW uint8; flow(k, pts) -> true positions in frame k+1."""

    # plane depth [m], half-extent in world x/y [m] (the farthest plane is unbounded)
    PLANES = ((2.0, 0.55), (4.0, 1.9), (8.0, np.inf))
    TEX_SIZE = 1024
    TEXELS_PER_M = (420.0, 170.0, 70.0)

    def __init__(self, stream_id=0, width=752, height=480, period=16, traj_seed=648):
        self.W, self.H, self.period = int(width), int(height), int(period)
        s = width / 752.0
        self.fx, self.fy, self.cx, self.cy = (_EUROC_K[0] * s, _EUROC_K[1] * s,
                                              _EUROC_K[2] * s, _EUROC_K[3] * height / 480.0)
        self._stream_id = stream_id
        self._tex_cache = None           # textures are built lazily: flow()/predict() never need them
        self._noise_seed = stream_id * 1000 + 7
        trng = np.random.default_rng(traj_seed + 31 * stream_id)
        # smooth periodic trajectory: <=1.5 deg/frame rotation, <=~4 px/frame flow at 4 m
        self._phase = trng.uniform(0, 2 * np.pi, 6)
        self._amp_t = trng.uniform(0.6, 1.0, 3) * np.array([0.035, 0.025, 0.05]) * self.period / (2 * np.pi)
        self._amp_r = trng.uniform(0.5, 1.0, 3) * np.deg2rad([0.5, 0.5, 1.0]) * self.period / (2 * np.pi)
        ys, xs = np.mgrid[0:self.H, 0:self.W]
        self._rays = np.stack([(xs - self.cx) / self.fx, (ys - self.cy) / self.fy, np.ones_like(xs, float)], -1)

    @property
    def _tex(self):
        if self._tex_cache is None:
            rng = np.random.default_rng(self._stream_id * 1000 + 7)
            self._tex_cache = [_texture(rng, self.TEX_SIZE) for _ in self.PLANES]
        return self._tex_cache

    # -- camera model
    def pose(self, k):
        w = 2 * np.pi * (k % self.period) / self.period
        t = self._amp_t * np.sin(w + self._phase[:3])
        r = self._amp_r * np.sin(w + self._phase[3:])
        return _rot(*r), t            # camera-to-world rotation, camera centre

    def K(self):
        """3x3 intrinsic matrix of the stream (Frame::K)."""
        return np.array([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]])

    def gyro_delta(self, k):
        """Rotation increment from frame k to k+1 as IMU pre-integration reports it (delta.q as a matrix, R_k^T R_k+1;
        camera and IMU frames coincide in the synthetic rig)."""
        R0, _ = self.pose(k)
        R1, _ = self.pose(k + 1)
        return R0.T @ R1

    def _hit(self, R, c, rays):
        """Intersect rays with the planes; returns world XY, plane index per ray."""
        d = rays @ R.T
        shape = d.shape[:-1]
        X = np.zeros(shape + (2,))
        which = np.full(shape, len(self.PLANES) - 1, np.int32)
        done = np.zeros(shape, bool)
        for i, (z, half) in enumerate(self.PLANES):
            tt = (z - c[2]) / d[..., 2]
            px, py = c[0] + tt * d[..., 0], c[1] + tt * d[..., 1]
            ok = (~done) & (np.abs(px) <= half) & (np.abs(py) <= half)
            X[ok, 0], X[ok, 1] = px[ok], py[ok]
            which[ok] = i
            done |= ok
        return X, which

    def frame(self, k):
        R, c = self.pose(k)
        X, which = self._hit(R, c, self._rays)
        img = np.zeros((self.H, self.W), np.float32)
        for i in range(len(self.PLANES)):
            m = which == i
            if not m.any():
                continue
            u = X[m, 0] * self.TEXELS_PER_M[i] + self.TEX_SIZE / 2
            v = X[m, 1] * self.TEXELS_PER_M[i] + self.TEX_SIZE / 2
            img[m] = ndimage.map_coordinates(self._tex[i], [v, u], order=1, mode="wrap")
        rng = np.random.default_rng((self._noise_seed, k % self.period))
        gain = 1.0 + 0.1 * np.sin(2 * np.pi * (k % self.period) / self.period + self._phase[0])
        img = img * gain * 0.8 + 20.0 + rng.normal(0.0, 2.0, img.shape).astype(np.float32)
        return np.clip(np.rint(img), 0, 255).astype(np.uint8)

    def flow(self, k, pts):
        """True correspondence of pixel positions `pts` (N,2) of frame k in frame k+1."""
        pts = np.asarray(pts, np.float64).reshape(-1, 2)
        R0, c0 = self.pose(k)
        R1, c1 = self.pose(k + 1)
        rays = np.stack([(pts[:, 0] - self.cx) / self.fx, (pts[:, 1] - self.cy) / self.fy, np.ones(len(pts))], -1)
        X, which = self._hit(R0, c0, rays)
        z = np.array([p[0] for p in self.PLANES])[which]
        Pw = np.stack([X[:, 0], X[:, 1], z], -1)
        Pc = (Pw - c1) @ R1           # world -> camera 1 (R1 is camera-to-world)
        return np.stack([self.fx * Pc[:, 0] / Pc[:, 2] + self.cx, self.fy * Pc[:, 1] / Pc[:, 2] + self.cy], -1)

    def predict(self, k, pts, noise_px=1.0):
        """IMU-style prediction: true flow + N(0, noise_px)."""
        rng = np.random.default_rng((self._noise_seed, 7919, k % self.period))
        q = self.flow(k, pts)
        return q + rng.normal(0.0, noise_px, q.shape)
