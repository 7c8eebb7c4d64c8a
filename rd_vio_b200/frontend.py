"""Python host side of the B200 front-end: a batched `FrontEnd` context over the C ABI and
`GpuImage`, the mirror of the reference's Image plugin interface
(rdvio::Image, /root/reference/src/rdvio/include/rdvio/types.h:153-177, as implemented by
rdvio::extra::OpenCvImage, src/rdvio_extra/src/opencv_image.cpp) -- same member names, argument
meaning and failure behaviour, so parity tests read like tests of the reference class.

All compute happens in librdvio_fe.so (CUDA, sm_100a).  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _native as N


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class FrontEnd:
    """One context per (GPU, host thread): slot pool + batched preprocess / detect / track."""

    def __init__(self, width=752, height=480, max_level=3, win=21, num_slots=2, max_points=512, device=0,
                 stream: Optional[int] = None):
        self._h = C.c_void_p()
        self._L = N.lib()
        cfg = N.Config(device, width, height, max_level, win, num_slots, max_points, stream)
        N.check(self._L.rdfe_create(C.byref(cfg), C.byref(self._h)), "rdfe_create")
        self.width, self.height, self.max_level, self.win = width, height, max_level, win
        self.num_slots, self.max_points, self.device = num_slots, max_points, device
        self.nlevels = self._L.rdfe_num_levels(self._h)

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.rdfe_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def handle(self):
        return self._h

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        N.check(self._L.rdfe_level_size(self._h, level, C.byref(w), C.byref(h)), "rdfe_level_size")
        return w.value, h.value

    def acquire(self) -> int:
        s = C.c_int()
        N.check(self._L.rdfe_slot_acquire(self._h, C.byref(s)), "rdfe_slot_acquire")
        return s.value

    def release(self, slot: int):
        N.check(self._L.rdfe_slot_release(self._h, slot), "rdfe_slot_release")

    def sync(self):
        N.check(self._L.rdfe_sync(self._h), "rdfe_sync")

    def kernel_launches(self) -> int:
        return int(self._L.rdfe_kernel_launches(self._h))

    def set_undistort(self, K=None, D=None):
        """cv::undistort(img, out, K, D) in front of every preprocess (examples/dataset.hpp:232-236); None switches off."""
        if K is None or D is None:
            N.check(self._L.rdfe_set_undistort(self._h, None, None), "rdfe_set_undistort")
            return
        K = np.ascontiguousarray(K, np.float32).reshape(9)
        D = np.ascontiguousarray(np.asarray(D, np.float32).reshape(-1)[:4])
        N.check(self._L.rdfe_set_undistort(self._h, _vp(K), _vp(D)), "rdfe_set_undistort")

    def set_template_cache(self, on: bool = True):
        """LK template cache (rdfe_set_template_cache): the backward pass's templates serve the next forward pass."""
        N.check(self._L.rdfe_set_template_cache(self._h, 1 if on else 0), "rdfe_set_template_cache")

    def template_cache_stats(self, reset=False):
        """(lookups, hits) of the LK template cache since creation / the last reset."""
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        N.check(self._L.rdfe_template_cache_stats(self._h, C.byref(a), C.byref(b), 1 if reset else 0), "rdfe_template_cache_stats")
        return int(a.value), int(b.value)

    def set_input_format(self, channels: int):
        """1 = gray (default), 3 = BGR, 4 = BGRA: cv::cvtColor of Odometry::addFrame (rdvio.hpp:42-49) on the device."""
        N.check(self._L.rdfe_set_input_format(self._h, int(channels)), "rdfe_set_input_format")
        self.channels = int(channels)

    # -- batched host-pointer API
    def preprocess(self, slots: Sequence[int], images: Sequence[np.ndarray], clip_limit=6.0, tiles=(8, 8)):
        n = len(slots)
        ch = getattr(self, "channels", 1)
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        want = (self.height, self.width) if ch == 1 else (self.height, self.width, ch)
        for im in imgs:
            if im.shape != want:
                raise ValueError(f"image shape {im.shape} != {want}")
        sl = np.asarray(slots, np.int32)
        ptrs = (C.c_void_p * n)(*[im.ctypes.data for im in imgs])
        N.check(self._L.rdfe_preprocess_batch(self._h, _vp(sl), n, ptrs, self.width * ch, float(clip_limit),
                                              int(tiles[0]), int(tiles[1])), "rdfe_preprocess_batch")

    def detect_params(self, **kw) -> N.DetectParams:
        p = N.DetectParams()
        self._L.rdfe_default_detect_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def track_params(self, **kw) -> N.TrackParams:
        p = N.TrackParams()
        self._L.rdfe_default_track_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def detect(self, slots: Sequence[int], existing: Sequence[np.ndarray], max_points=150, keypoint_distance=20.0,
               stride: Optional[int] = None, return_gftt=False, **params):
        """OpenCvImage::detect_keypoints for n images. existing[i]: (k_i,2) float64 already-tracked keypoints.
        Returns list of (k_i + new_i, 2) arrays (existing first, new appended)."""
        n = len(slots)
        ex = [np.asarray(e, np.float64).reshape(-1, 2) for e in existing]
        if stride is None:
            stride = max(len(e) for e in ex) + max_points
        if stride > self.max_points:       # the reference's vector is unbounded: never clamp silently
            raise ValueError(f"existing + max_points = {stride} keypoints exceed the context capacity {self.max_points}")
        xy = np.zeros((n, stride, 2), np.float64)
        counts = np.zeros(n, np.int32)
        for i, e in enumerate(ex):
            if len(e) > stride:
                raise ValueError("more existing keypoints than capacity")
            xy[i, :len(e)] = e
            counts[i] = len(e)
        p = self.detect_params(max_points=max_points, keypoint_distance=keypoint_distance, **params)
        gxy = np.zeros((n, max_points, 2), np.float32) if return_gftt else None
        gre = np.zeros((n, max_points), np.float32) if return_gftt else None
        gcn = np.zeros(n, np.int32) if return_gftt else None
        sl = np.asarray(slots, np.int32)
        N.check(self._L.rdfe_detect_batch(self._h, _vp(sl), n, C.byref(p), _vp(xy), _vp(counts), stride,
                                          _vp(gxy), _vp(gre), _vp(gcn)), "rdfe_detect_batch")
        out = [xy[i, :counts[i]].copy() for i in range(n)]
        if return_gftt:
            return out, [gxy[i, :gcn[i]].copy() for i in range(n)], [gre[i, :gcn[i]].copy() for i in range(n)]
        return out

    def detect_prefetch(self, slots: Sequence[int], max_points=150, **params):
        """Start Harris + GFTT selection of preprocessed slots early (rdfe_detect_prefetch); a following detect() on
        the same slots with the same GFTT parameters only runs the Poisson append."""
        p = self.detect_params(max_points=max_points, **params)
        sl = np.asarray(slots, np.int32)
        N.check(self._L.rdfe_detect_prefetch(self._h, _vp(sl), len(sl), C.byref(p)), "rdfe_detect_prefetch")

    def track(self, curr_slots, next_slots, curr: Sequence[np.ndarray], pred: Optional[Sequence[np.ndarray]] = None,
              **params):
        """OpenCvImage::track_keypoints for n image pairs. Returns (next list, status list)."""
        n = len(curr_slots)
        cu = [np.asarray(c, np.float64).reshape(-1, 2) for c in curr]
        stride = max(1, max(len(c) for c in cu))
        if stride > self.max_points:
            raise ValueError("more keypoints than capacity")
        cxy = np.zeros((n, stride, 2), np.float64)
        nxy = np.zeros((n, stride, 2), np.float64)
        counts = np.zeros(n, np.int32)
        for i, c in enumerate(cu):
            cxy[i, :len(c)] = c
            counts[i] = len(c)
            if pred is not None:
                nxy[i, :len(c)] = np.asarray(pred[i], np.float64).reshape(-1, 2)
        st = np.zeros((n, stride), np.int8)
        p = self.track_params(has_prediction=1 if pred is not None else 0, **params)
        a, b = np.asarray(curr_slots, np.int32), np.asarray(next_slots, np.int32)
        N.check(self._L.rdfe_track_batch(self._h, _vp(a), _vp(b), n, C.byref(p), _vp(cxy), _vp(nxy), _vp(counts),
                                         stride, _vp(st)), "rdfe_track_batch")
        return [nxy[i, :counts[i]].copy() for i in range(n)], [st[i, :counts[i]].copy() for i in range(n)]

    # -- parity taps
    def download_level(self, slot, level, plane=0):
        w, h = self.level_size(level)
        if plane == 0:
            out = np.empty((h, w), np.uint8)
        elif plane == 1:
            out = np.empty((h, w, 2), np.int16)
        elif plane == 3:
            out = np.empty((h, w), np.uint8)
        else:
            out = np.empty((h + 2 * self.win, w + 2 * self.win), np.uint8)
        N.check(self._L.rdfe_download_level(self._h, slot, level, plane, _vp(out), out.nbytes), "rdfe_download_level")
        return out

    def upload_level0(self, slot, image):
        """Test tap: replace level 0 of `slot` (normally the CLAHE output) by `image` with its REFLECT_101 halo."""
        img = np.ascontiguousarray(image, np.uint8)
        if img.shape != (self.height, self.width):
            raise ValueError(f"image shape {img.shape} != {(self.height, self.width)}")
        padded = np.ascontiguousarray(np.pad(img, self.win, mode="reflect"))      # numpy 'reflect' == BORDER_REFLECT_101
        N.check(self._L.rdfe_upload_level0(self._h, slot, _vp(padded), padded.nbytes), "rdfe_upload_level0")

    def download_clahe_lut(self, batch_index=0, tiles=64):
        out = np.empty((tiles, 256), np.uint8)
        N.check(self._L.rdfe_download_clahe_lut(self._h, batch_index, _vp(out), out.nbytes), "rdfe_download_clahe_lut")
        return out

    def harris_response(self, slot, **params):
        out = np.empty((self.height, self.width), np.float32)
        p = self.detect_params(**params)
        N.check(self._L.rdfe_harris_response(self._h, slot, C.byref(p), _vp(out), out.nbytes), "rdfe_harris_response")
        return out

    def harris_candidates(self, slot, **params):
        """Keys (response bits << 32 | y*w+x) the Harris stage of the hot path emits, the frame maximum and the number of
        pixels its prefilter flagged for exact evaluation: (keys uint64 sorted descending, frame_max, flagged)."""
        cap = self.width * self.height // 2
        keys = np.empty(cap, np.uint64)
        n, fm, nf = C.c_uint(0), C.c_float(0.0), C.c_uint(0)
        p = self.detect_params(**params)
        N.check(self._L.rdfe_harris_candidates(self._h, slot, C.byref(p), _vp(keys), cap, C.byref(n), C.byref(fm), C.byref(nf)),
                "rdfe_harris_candidates")
        return np.sort(keys[:n.value])[::-1].copy(), float(fm.value), int(nf.value)


class GpuImage:
    """Mirror of rdvio::extra::OpenCvImage (opencv_image.h:9-56) on top of a FrontEnd context.

    Like the reference class it owns `image` (8-bit gray) and `t`, is preprocessed once, tracked
    from once, detected on once and then released (FeatureTracker::run, feature_tracker.cpp:32-98).
    CLAHE / GFTT parameters are frozen by the FIRST call in the process, exactly like the reference's
    function-local static singletons (opencv_image.cpp:179-188).
    """

    _frozen_clahe = None      # (clip, width, height) of the first preprocess() in the process
    _frozen_max_points = None

    def __init__(self, fe: FrontEnd, image: np.ndarray, t: float = 0.0):
        self.fe = fe
        self.image = np.ascontiguousarray(image, np.uint8)
        self.raw = self.image
        self.t = t
        self._slot = None
        self._w, self._h = self.image.shape[1], self.image.shape[0]

    def width(self):
        return self._w

    def height(self):
        return self._h

    def level_num(self):
        return self.fe.max_level

    def get_rawdata(self):
        return self.raw

    def preprocess(self, clipLimit: float, width: int, height: int):
        if GpuImage._frozen_clahe is None:
            GpuImage._frozen_clahe = (clipLimit, width, height)
        clip, tx, ty = GpuImage._frozen_clahe
        if self._slot is None:
            self._slot = self.fe.acquire()
        self.fe.preprocess([self._slot], [self.image], clip, (tx, ty))
        if GpuImage._frozen_max_points is not None:      # like the C++ class: selection runs beside the tracking call
            self.fe.detect_prefetch([self._slot], GpuImage._frozen_max_points)

    def detect_keypoints(self, keypoints, max_points=1000, keypoint_distance=10.0):
        """Appends new corners to `keypoints` (list/array of (x,y)); returns the new array."""
        if GpuImage._frozen_max_points is None:
            GpuImage._frozen_max_points = int(max_points)
        kp = np.asarray(keypoints, np.float64).reshape(-1, 2)
        if self._slot is None:
            return kp
        return self.fe.detect([self._slot], [kp], GpuImage._frozen_max_points, keypoint_distance)[0]

    def track_keypoints(self, next_image, curr_keypoints, next_keypoints=None):
        """Returns (next_keypoints, result_status). A next_image of another type or without pixels yields
        all-zero status, like the failed dynamic_cast in the reference (opencv_image.cpp:88-92)."""
        curr = np.asarray(curr_keypoints, np.float64).reshape(-1, 2)
        has_pred = next_keypoints is not None and len(next_keypoints) > 0
        nxt = np.asarray(next_keypoints, np.float64).reshape(-1, 2).copy() if has_pred else np.zeros_like(curr)
        status = np.zeros(len(curr), np.int8)
        if (not isinstance(next_image, GpuImage) or next_image._slot is None or self._slot is None
                or len(curr) == 0):
            return nxt, status
        res, st = self.fe.track([self._slot], [next_image._slot], [curr], [nxt] if has_pred else None)
        ok = st[0] != 0
        nxt[ok] = res[0][ok]
        return nxt, st[0]

    def release_image_buffer(self):
        if self._slot is not None:
            self.fe.release(self._slot)
            self._slot = None
        self.image = None
        self.raw = None

    @classmethod
    def reset_frozen_parameters(cls):
        """Test hook: forget the process-wide frozen CLAHE/GFTT parameters."""
        cls._frozen_clahe = None
        cls._frozen_max_points = None
