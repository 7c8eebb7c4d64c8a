"""ctypes binding of the CPU ORACLE (oracle/fe_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (rd_vio_b200) never does.

Reference being restated: rdvio::extra::OpenCvImage
(/root/reference/src/rdvio_extra/src/opencv_image.cpp) and the OpenCV 4.x calls
it makes; see the header of fe_oracle.c for the per-function citations.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfe_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/fe_oracle.c with gcc (building the checker is not using it)."""
    src = os.path.join(_HERE, "fe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, i16p, f32p, f64p, i32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int16),
                                       C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int))
        L.orc_clahe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                C.c_void_p, C.c_int, C.c_void_p]
        L.orc_clahe.restype = C.c_int
        L.orc_pyrdown.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_scharr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_build_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_build_pyramid.restype = C.c_void_p
        L.orc_pyramid_free.argtypes = [C.c_void_p]
        L.orc_pyramid_free.restype = None
        for nm in ("orc_pyramid_levels",):
            getattr(L, nm).argtypes = [C.c_void_p]
        for nm in ("orc_pyramid_width", "orc_pyramid_height"):
            getattr(L, nm).argtypes = [C.c_void_p, C.c_int]
        L.orc_pyramid_image.argtypes = [C.c_void_p, C.c_int]
        L.orc_pyramid_image.restype = C.c_void_p
        L.orc_pyramid_deriv.argtypes = [C.c_void_p, C.c_int]
        L.orc_pyramid_deriv.restype = C.c_void_p
        L.orc_harris.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int]
        L.orc_gftt_select.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_poisson_filter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.orc_detect_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.orc_track_keypoints.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        del u8p, i16p, f32p, f64p, i32p
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


# ------------------------------------------------------------------ stages
def clahe(img, clip_limit=6.0, tiles_x=8, tiles_y=8, return_lut=False):
    img = _u8(img)
    H, W = img.shape
    dst = np.empty_like(img)
    lut = np.empty((tiles_y * tiles_x, 256), np.uint8)
    rc = lib().orc_clahe(_p(img), W, H, W, float(clip_limit), tiles_x, tiles_y, _p(dst), W, _p(lut))
    if rc != 0:
        raise RuntimeError(f"orc_clahe rc={rc}")
    return (dst, lut) if return_lut else dst


def pyrdown(img):
    img = _u8(img)
    H, W = img.shape
    dst = np.empty(((H + 1) // 2, (W + 1) // 2), np.uint8)
    lib().orc_pyrdown(_p(img), W, H, W, _p(dst), dst.shape[1])
    return dst


def scharr(img):
    img = _u8(img)
    H, W = img.shape
    dst = np.empty((H, W, 2), np.int16)
    lib().orc_scharr(_p(img), W, H, W, _p(dst), 2 * W)
    return dst


class Pyramid:
    """orc_pyramid handle: what cv::buildOpticalFlowPyramid(..., withDerivatives=true) yields."""

    def __init__(self, img, win=21, max_level=3):
        img = _u8(img)
        H, W = img.shape
        self.shape = (H, W)
        self.win, self.max_level = win, max_level
        self._h = lib().orc_build_pyramid(_p(img), W, H, W, win, max_level)
        if not self._h:
            raise RuntimeError("orc_build_pyramid failed")
        self.nlevels = lib().orc_pyramid_levels(self._h)

    def level_shape(self, l):
        return lib().orc_pyramid_height(self._h, l), lib().orc_pyramid_width(self._h, l)

    def image(self, l):
        h, w = self.level_shape(l)
        buf = (C.c_uint8 * (h * w)).from_address(lib().orc_pyramid_image(self._h, l))
        return np.frombuffer(buf, np.uint8).reshape(h, w).copy()

    def deriv(self, l):
        h, w = self.level_shape(l)
        buf = (C.c_int16 * (h * w * 2)).from_address(lib().orc_pyramid_deriv(self._h, l))
        return np.frombuffer(buf, np.int16).reshape(h, w, 2).copy()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_pyramid_free(self._h)
            self._h = None


def harris(img, k=0.04, mode=0):
    img = _u8(img)
    H, W = img.shape
    R = np.empty((H, W), np.float32)
    lib().orc_harris(_p(img), W, H, W, C.c_float(k), _p(R), mode)
    return R


def gftt_select(R, max_corners, quality=1e-3, min_distance=20.0):
    R = np.ascontiguousarray(R, np.float32)
    H, W = R.shape
    cap = max_corners if max_corners > 0 else H * W
    xy = np.empty((cap, 2), np.float32)
    resp = np.empty(cap, np.float32)
    nc = C.c_int(0)
    n = lib().orc_gftt_select(_p(R), W, H, max_corners, quality, min_distance, _p(xy), _p(resp), C.byref(nc))
    return xy[:n].copy(), resp[:n].copy(), nc.value


def poisson_filter(existing_xy, cand_xy, radius):
    ex = np.ascontiguousarray(existing_xy, np.float64).reshape(-1, 2)
    ca = np.ascontiguousarray(cand_xy, np.float64).reshape(-1, 2)
    out = np.empty_like(ca)
    n = lib().orc_poisson_filter(_p(ex), len(ex), _p(ca), len(ca), float(radius), _p(out))
    return out[:n].copy()


def detect_keypoints(img, existing_xy, max_points=150, keypoint_distance=20.0, harris_mode=0):
    """OpenCvImage::detect_keypoints on a preprocessed image. Returns (all_keypoints, gftt_xy, gftt_resp)."""
    img = _u8(img)
    H, W = img.shape
    ex = np.ascontiguousarray(existing_xy, np.float64).reshape(-1, 2)
    kp = np.empty((len(ex) + max_points, 2), np.float64)
    kp[:len(ex)] = ex
    gxy = np.empty((max_points, 2), np.float32)
    gre = np.empty(max_points, np.float32)
    gn = C.c_int(0)
    tot = lib().orc_detect_keypoints(_p(img), W, H, W, _p(kp), len(ex), max_points, float(keypoint_distance),
                                     harris_mode, _p(gxy), _p(gre), C.byref(gn))
    return kp[:tot].copy(), gxy[:gn.value].copy(), gre[:gn.value].copy()


def lk(prev: Pyramid, nxt: Pyramid, prev_xy, init_xy, win=21, max_level=3, max_count=30, eps=0.01,
       return_iters=False):
    p = np.ascontiguousarray(prev_xy, np.float32).reshape(-1, 2)
    q = np.ascontiguousarray(init_xy, np.float32).reshape(-1, 2).copy()
    n = len(p)
    st = np.zeros(n, np.uint8)
    it = np.zeros((n, max_level + 1), np.int32) if return_iters else None
    lib().orc_lk(prev._h, nxt._h, _p(p), _p(q), _p(st), n, win, max_level, max_count, float(eps), _p(it))
    return (q, st, it) if return_iters else (q, st)


def track_keypoints(curr: Pyramid, nxt: Pyramid, curr_xy, pred_xy=None, win=21, max_level=3):
    """OpenCvImage::track_keypoints. Returns (next_xy float64, status int8, raw forward LK xy float32)."""
    c = np.ascontiguousarray(curr_xy, np.float64).reshape(-1, 2)
    n = len(c)
    has_pred = pred_xy is not None and len(pred_xy) > 0
    q = np.ascontiguousarray(pred_xy, np.float64).reshape(-1, 2).copy() if has_pred else np.zeros((n, 2), np.float64)
    st = np.zeros(n, np.int8)
    fwd = np.zeros((n, 2), np.float32)
    H, W = curr.shape
    lib().orc_track_keypoints(curr._h, nxt._h, W, H, _p(c), _p(q), int(has_pred), _p(st), n, win, max_level, _p(fwd))
    return q, st, fwd


# ------------------------------------------------------------------ undistort
def undistort_map(W, H, K, D):
    """Fixed-point map of cv::undistort(src, dst, K, D) (K 3x3, D = k1 k2 p1 p2; both as float32 like the reference)."""
    L = lib()
    L.orc_undistort_map.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_remap_bilinear.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    K = np.ascontiguousarray(K, np.float32).reshape(9)
    D = np.ascontiguousarray(D, np.float32).reshape(-1)[:4]
    mxy = np.empty((H, W, 2), np.int16)
    mf = np.empty((H, W), np.uint16)
    L.orc_undistort_map(W, H, _p(K), _p(D), _p(mxy), _p(mf))
    return mxy, mf


def undistort(img, K, D):
    img = _u8(img)
    H, W = img.shape
    mxy, mf = undistort_map(W, H, K, D)
    dst = np.empty_like(img)
    lib().orc_remap_bilinear(_p(img), W, H, W, _p(mxy), _p(mf), _p(dst), W)
    return dst


# ------------------------------------------------------------------ gray conversion
def bgr2gray(img):
    """cv::cvtColor(img, COLOR_BGR2GRAY / COLOR_BGRA2GRAY) for 8-bit (Odometry::addFrame, rdvio.hpp:42-49):
    15-bit fixed point (B*3735 + G*19235 + R*9798 + 2^14) >> 15."""
    img = np.ascontiguousarray(img, np.uint8)
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def undistort_color(img, K, D):
    """cv::undistort on a 3/4-channel frame: the same remap per channel (alpha, if any, is dropped here)."""
    img = np.ascontiguousarray(img, np.uint8)
    return np.stack([undistort(np.ascontiguousarray(img[..., c]), K, D) for c in range(3)], -1)
