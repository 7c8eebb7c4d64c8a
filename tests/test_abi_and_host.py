"""CPU: the C-ABI library loads without a GPU and exports every symbol include/rdvio_fe.h declares; host-side logic."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rdvio_fe.h")).read()
    return sorted(set(re.findall(r"RDFE_API[^;(]*?\b(rdfe_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rd_vio_b200 import _native
    assert os.path.exists(_native.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_native.LIB_PATH)
    declared = header_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"not exported: {missing}"
    assert sorted(_native.SYMBOLS) == declared, "python binding table and header disagree"
    lib.rdfe_abi_version.restype = ctypes.c_int
    assert lib.rdfe_abi_version() == 1


def test_no_gpu_calls_fail_loudly_not_silently():
    """Without a CUDA device rdfe_create must return an error (there is no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rd_vio_b200 import _native as N
    from rd_vio_b200.frontend import FrontEnd
    with pytest.raises(N.FrontEndError):
        FrontEnd(752, 480)


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rd_vio_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "fe_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_byte_model_matches_survey():
    from rd_vio_b200.workload import algorithmic_bytes
    assert algorithmic_bytes(752, 480, 150, 3, 21)[0] == 6523440
    assert algorithmic_bytes(1280, 720, 300, 4, 21)[0] == 15765600
    assert algorithmic_bytes(1920, 1080, 1000, 5, 31)[0] == 46207360


def test_synthetic_stream_is_deterministic_and_periodic():
    from rd_vio_b200.synthetic import SyntheticStream
    a, b = SyntheticStream(5, 188, 120, period=4), SyntheticStream(5, 188, 120, period=4)
    assert np.array_equal(a.frame(1), b.frame(1)) and np.array_equal(a.frame(1), a.frame(5))
    pts = np.array([[50.0, 60.0], [100.0, 30.0]])
    assert np.abs(a.flow(1, pts) - pts).max() < 8.0
    assert not np.array_equal(SyntheticStream(6, 188, 120, period=4).frame(1), a.frame(1))


def test_cpp_plugin_header_compiles_against_shims():
    """The drop-in C++ Image subclass keeps the reference signatures (types.h:153-177): syntax-check it
    against minimal Eigen/cv::Mat shims (the real headers are absent from this image)."""
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "cpp", "shim"), os.path.join(ROOT, "tests", "cpp", "plugin_replay.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_header_is_plain_c_and_links(tmp_path):
    """include/rdvio_fe.h is a C ABI: it must compile as C99 (no C++ or torch types in any signature), a C program
    must link against librdvio_fe.so, and the entry points that need no GPU must answer; with no device,
    rdfe_create must fail with RDFE_ERR_CUDA / RDFE_ERR_UNSUPPORTED and leave a message in rdfe_last_error()."""
    from rd_vio_b200 import _native
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "rdvio_fe.h"
int main(void) {
    rdfe_detect_params d; rdfe_track_params t; rdfe_config cfg; rdfe_ctx *ctx = 0;
    if (rdfe_abi_version() != RDFE_ABI_VERSION) return 1;
    rdfe_default_detect_params(&d); rdfe_default_track_params(&t);
    if (d.max_points != 150 || d.quality_level != 1e-3 || d.min_distance != 20.0 || d.harris_k != 0.04 ||
        d.keypoint_distance != 20.0 || d.border != 20 || d.harris_fma != 0) return 2;      /* opencv_image.cpp:184-188, :61-68 */
    if (t.max_count != 30 || t.epsilon != 0.01 || t.min_eig_threshold != 1e-4 || t.border != 20 ||
        t.max_round_trip != 0.5) return 3;                                                   /* opencv_image.cpp:94-134 */
    if (rdfe_profile_num_kernels() < 8 || !rdfe_profile_kernel_name(0)) return 4;
    memset(&cfg, 0, sizeof cfg);
    cfg.width = 752; cfg.height = 480; cfg.max_level = 3; cfg.win = 21; cfg.num_slots = 2; cfg.max_points = 256;
    int rc = rdfe_create(&cfg, &ctx);
    printf("create rc=%d err=%s\n", rc, rdfe_last_error());
    if (rc == RDFE_OK) { rdfe_destroy(ctx); return 0; }
    if (ctx != 0 || !rdfe_last_error() || !rdfe_last_error()[0]) return 5;
    return 0;
}
''')
    exe = tmp_path / "abi"
    lib_dir = os.path.dirname(_native.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                        str(src), "-o", str(exe), "-L", lib_dir, "-lrdvio_fe", f"-Wl,-rpath,{lib_dir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    import torch
    if not torch.cuda.is_available():
        assert "create rc=-" in r.stdout and "err=" in r.stdout and len(r.stdout.strip().split("err=")[1]) > 3, r.stdout
