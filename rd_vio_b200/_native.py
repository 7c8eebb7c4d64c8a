"""ctypes binding of librdvio_fe.so (C ABI declared in include/rdvio_fe.h).

The product path has NO CPU fallback: if the CUDA library is missing or fails to load, importing
this module's `lib()` raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C rd_vio_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RDFE_LIB_PATH: an alternative build of the same library (kernel A/B experiments, scripts/lk_variants.sh); never a fallback
LIB_PATH = os.environ.get("RDFE_LIB_PATH") or os.path.join(_HERE, "lib", "librdvio_fe.so")

RDFE_MAX_BATCH = 128
RDFE_MAX_LEVELS = 8


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("width", C.c_int), ("height", C.c_int), ("max_level", C.c_int),
                ("win", C.c_int), ("num_slots", C.c_int), ("max_points", C.c_int), ("stream", C.c_void_p)]


class DetectParams(C.Structure):
    _fields_ = [("max_points", C.c_int), ("quality_level", C.c_double), ("min_distance", C.c_double),
                ("harris_k", C.c_double), ("keypoint_distance", C.c_double), ("border", C.c_int),
                ("harris_fma", C.c_int)]


class TrackParams(C.Structure):
    _fields_ = [("max_count", C.c_int), ("epsilon", C.c_double), ("min_eig_threshold", C.c_double),
                ("border", C.c_int), ("max_round_trip", C.c_double), ("has_prediction", C.c_int)]


# every symbol include/rdvio_fe.h declares: name -> (restype, argtypes)
_vp, _i, _sz, _d = C.c_void_p, C.c_int, C.c_size_t, C.c_double
SYMBOLS = {
    "rdfe_last_error": (C.c_char_p, []),
    "rdfe_abi_version": (_i, []),
    "rdfe_default_detect_params": (None, [C.POINTER(DetectParams)]),
    "rdfe_default_track_params": (None, [C.POINTER(TrackParams)]),
    "rdfe_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "rdfe_destroy": (None, [_vp]),
    "rdfe_num_levels": (_i, [_vp]),
    "rdfe_level_size": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i)]),
    "rdfe_sync": (_i, [_vp]),
    "rdfe_stream": (_vp, [_vp]),
    "rdfe_kernel_launches": (C.c_int64, [_vp]),
    "rdfe_slot_acquire": (_i, [_vp, C.POINTER(_i)]),
    "rdfe_slot_release": (_i, [_vp, _i]),
    "rdfe_preprocess_batch": (_i, [_vp, _vp, _i, _vp, _sz, _d, _i, _i]),
    "rdfe_preprocess_batch_dev": (_i, [_vp, _vp, _i, _vp, _sz, _d, _i, _i]),
    "rdfe_detect_batch": (_i, [_vp, _vp, _i, C.POINTER(DetectParams), _vp, _vp, _i, _vp, _vp, _vp]),
    "rdfe_detect_batch_dev": (_i, [_vp, _vp, _i, C.POINTER(DetectParams), _vp, _vp, _i, _vp, _vp, _vp]),
    "rdfe_track_batch": (_i, [_vp, _vp, _vp, _i, C.POINTER(TrackParams), _vp, _vp, _vp, _i, _vp]),
    "rdfe_track_batch_dev": (_i, [_vp, _vp, _vp, _i, C.POINTER(TrackParams), _vp, _vp, _vp, _i, _vp]),
    "rdfe_frontend_step_dev": (_i, [_vp, _vp, _vp, _i, _vp, _sz, _d, _i, _i, C.POINTER(TrackParams), _vp, _vp, _vp, _vp,
                                    C.POINTER(DetectParams), _vp, _i]),
    "rdfe_set_undistort": (_i, [_vp, _vp, _vp]),
    "rdfe_set_input_format": (_i, [_vp, _i]),
    "rdfe_set_pipelining": (_i, [_vp, _i]),
    "rdfe_set_step_compaction": (_i, [_vp, _i]),
    "rdfe_predict_rotation_dev": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp]),
    "rdfe_set_template_cache": (_i, [_vp, _i]),
    "rdfe_template_cache_stats": (_i, [_vp, _vp, _vp, _i]),
    "rdfe_frontend_step_submit": (_i, [_vp, _vp, _vp, _i, _vp, _sz, _d, _i, _i, C.POINTER(TrackParams), _vp, _vp, _vp,
                                       C.POINTER(DetectParams), _i, C.POINTER(_i)]),
    "rdfe_frontend_step_wait": (_i, [_vp, _i, _vp, _vp, _vp]),
    "rdfe_upload_only": (_i, [_vp, _vp, _i, _vp, _sz, _i]),
    "rdfe_download_level": (_i, [_vp, _i, _i, _i, _vp, _sz]),
    "rdfe_download_clahe_lut": (_i, [_vp, _i, _vp, _sz]),
    "rdfe_upload_level0": (_i, [_vp, _i, _vp, _sz]),
    "rdfe_harris_response": (_i, [_vp, _i, C.POINTER(DetectParams), _vp, _sz]),
    "rdfe_harris_candidates": (_i, [_vp, _i, C.POINTER(DetectParams), _vp, _sz, C.POINTER(C.c_uint), C.POINTER(C.c_float),
                                    C.POINTER(C.c_uint)]),
    "rdfe_harris_prefilter_constants": (None, [C.POINTER(C.c_float)]),
    "rdfe_dev_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "rdfe_dev_free": (_i, [_vp, _vp]),
    "rdfe_host_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "rdfe_host_free": (_i, [_vp, _vp]),
    "rdfe_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz, _i]),
    "rdfe_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz, _i]),
    "rdfe_profile_num_kernels": (_i, []),
    "rdfe_profile_kernel_name": (C.c_char_p, [_i]),
    "rdfe_profile_enable": (_i, [_vp, _i]),
    "rdfe_profile_collect": (_i, [_vp, _vp, _vp]),
    "rdfe_profile_timeline": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "rdfe_detect_prefetch": (_i, [_vp, _vp, _i, _vp]),
    "rdfe_set_host_sync": (_i, [_vp, _i]),
    "rdfe_timer_start": (_i, [_vp]),
    "rdfe_timer_stop": (_i, [_vp]),
    "rdfe_timer_elapsed_ms": (_i, [_vp, C.POINTER(C.c_float)]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load librdvio_fe.so; raises loudly (never falls back to a CPU path)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not built: run `make -C rd_vio_b200/csrc` (or __graft_entry__.build()); "
                "rd_vio_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class FrontEndError(RuntimeError):
    def __init__(self, rc, where):
        msg = lib().rdfe_last_error()
        super().__init__(f"{where}: rc={rc}: {msg.decode() if msg else ''}")
        self.rc = rc


def check(rc, where):
    if rc != 0:
        raise FrontEndError(rc, where)
