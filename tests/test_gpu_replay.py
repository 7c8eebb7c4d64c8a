"""GPU: FeatureTracker-level replay (config 5 proxy): the plugin call sequence of
FeatureTracker::run over a synthetic sequence, (a) through the Python mirror GpuImage and (b) through the C++
drop-in class rdvio::extra::GpuImage (include/rdvio_b200/gpu_image.hpp) compiled against the test shims,
compared per frame with the same loop run on the CPU oracle.  Identical keypoint streams into the (untouched,
deterministic-seeded) back-end imply an identical trajectory."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_FRAMES = 12


def oracle_replay(frames, use_prediction=None):
    from oracle import fe_oracle as orc
    out, last, last_kp = [], None, None
    for i, f in enumerate(frames):
        pre = orc.clahe(f)
        pyr = orc.Pyramid(pre, 21, 3)
        kp = np.zeros((0, 2))
        if last is not None:
            nxt, st, _ = orc.track_keypoints(last, pyr, last_kp, None)
            kp = nxt[st != 0]
        kp = orc.detect_keypoints(pre, kp, 150, 20.0)[0]
        out.append(kp)
        last, last_kp = pyr, kp
    return out


@pytest.fixture(scope="module")
def seq(stream0):
    return [stream0.frame(k) for k in range(N_FRAMES)]


@pytest.fixture(scope="module")
def ref(seq):
    return oracle_replay(seq)


def test_python_plugin_replay(seq, ref):
    from rd_vio_b200.frontend import FrontEnd, GpuImage
    GpuImage.reset_frozen_parameters()
    with FrontEnd(752, 480, 3, 21, num_slots=4, max_points=1024) as fe:
        last, last_kp = None, None
        for i, f in enumerate(seq):
            img = GpuImage(fe, f, t=0.05 * i)
            img.preprocess(6.0, 8, 8)
            kp = np.zeros((0, 2))
            if last is not None:
                nxt, st = last.track_keypoints(img, last_kp, None)
                kp = nxt[st != 0]
                last.release_image_buffer()
            kp = img.detect_keypoints(kp, 150, 20.0)
            assert kp.shape == ref[i].shape, f"frame {i}: {kp.shape} vs {ref[i].shape}"
            assert np.abs(kp - ref[i]).max() <= 0.01, f"frame {i}"
            last, last_kp = img, kp


def test_frozen_parameters_like_static_singletons(seq):
    """opencv_image.cpp:179-188: later clip/tile/max_points values are ignored."""
    from oracle import fe_oracle as orc
    from rd_vio_b200.frontend import FrontEnd, GpuImage
    GpuImage.reset_frozen_parameters()
    with FrontEnd(752, 480, 3, 21, num_slots=4, max_points=1024) as fe:
        a = GpuImage(fe, seq[0]); a.preprocess(6.0, 8, 8)
        b = GpuImage(fe, seq[0]); b.preprocess(2.0, 4, 4)          # ignored: frozen at (6.0, 8, 8)
        assert np.array_equal(fe.download_level(b._slot, 0, 0), orc.clahe(seq[0], 6.0, 8, 8))
        k1 = a.detect_keypoints(np.zeros((0, 2)), 100, 20.0)
        k2 = b.detect_keypoints(np.zeros((0, 2)), 150, 20.0)      # max_points frozen at 100
        assert len(k1) == len(k2) <= 100
        # wrong dynamic type / released image => all-zero status, no throw (opencv_image.cpp:88-92)
        nxt, st = a.track_keypoints(object(), k1, None)
        assert st.sum() == 0
        b.release_image_buffer()
        nxt, st = a.track_keypoints(b, k1, None)
        assert st.sum() == 0
    GpuImage.reset_frozen_parameters()


def _run_cpp_replay(seq, ref, tmp_path):
    exe = os.path.join(ROOT, "tests", "cpp", "plugin_replay")
    src = exe + ".cpp"
    lib_dir = os.path.join(ROOT, "rd_vio_b200", "lib")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I",
                        os.path.join(ROOT, "tests", "cpp", "shim"), src, "-o", exe, "-L", lib_dir, "-lrdvio_fe",
                        f"-Wl,-rpath,{lib_dir}"], check=True)
    fb, fo = tmp_path / "frames.bin", tmp_path / "out.txt"
    with open(fb, "wb") as f:
        np.array([len(seq), 480, 752], np.int32).tofile(f)
        for im in seq:
            im.tofile(f)
    env = dict(os.environ, LD_LIBRARY_PATH=lib_dir + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe, str(fb), str(fo)], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = open(fo).read().split("\n")
    i = 0
    for fr in range(len(seq)):
        tag, idx, n = lines[i].split()
        assert tag == "frame" and int(idx) == fr
        n = int(n)
        kp = np.array([[float(x) for x in lines[i + 1 + j].split()] for j in range(n)]).reshape(-1, 2)
        i += 1 + n
        assert kp.shape == ref[fr].shape, f"frame {fr}: {kp.shape} vs {ref[fr].shape}"
        assert np.abs(kp - ref[fr]).max() <= 0.01, f"frame {fr}"
    return r.stderr


def test_cpp_plugin_replay(seq, ref, tmp_path):
    _run_cpp_replay(seq, ref, tmp_path)


def test_cpp_plugin_replay_200_frames(tmp_path):
    """SURVEY.md 8(d) config 5 (as far as it can go without Eigen/Ceres): >= 200 frames of FeatureTracker's call
    sequence through the C++ drop-in class, every frame's keypoints against the oracle replay; the per-frame
    front-end latency the binary reports is written to gpurun_out/plugin_latency.txt."""
    from rd_vio_b200.synthetic import SyntheticStream
    st = SyntheticStream(3, 752, 480, period=200)
    frames = [st.frame(k) for k in range(200)]
    err = _run_cpp_replay(frames, oracle_replay(frames), tmp_path)
    lat = [l for l in err.splitlines() if l.startswith("latency_us")]
    assert lat, err
    print(lat[0])
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "plugin_latency.txt"), "w") as f:
            f.write(lat[0] + "\n")


def test_feature_tracker_replay_with_imu_prediction():
    """the plugin driven by the reference's callers (oracle/frame_host.py: pixels from unit
    bearings through K, predictions = bearings rotated by the gyro increment, track-length-ordered Poisson filter,
    detect on what survived), 40 frames, GPU plugin against the same loop on the CPU oracle: identical track ids
    and keypoint counts in every frame, positions within 0.01 px."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from frame_helpers import OracleImage, replay
    from rd_vio_b200.frontend import FrontEnd, GpuImage
    from rd_vio_b200.synthetic import SyntheticStream
    st = SyntheticStream(5, 752, 480, period=64)
    ref = replay(st, 40, lambda im, t: OracleImage(im, t))
    GpuImage.reset_frozen_parameters()
    with FrontEnd(752, 480, 3, 21, num_slots=4, max_points=1024) as fe:
        got = replay(st, 40, lambda im, t: GpuImage(fe, im, t))
    GpuImage.reset_frozen_parameters()
    long_tracks = 0
    for i, ((kp, ids), (rkp, rids)) in enumerate(zip(got, ref)):
        assert ids == rids, f"frame {i}: track ids differ"
        assert kp.shape == rkp.shape and np.abs(kp - rkp).max() <= 0.01, f"frame {i}"
        long_tracks = max(long_tracks, sum(1 for t in ids if t >= 0))
    assert long_tracks >= 80          # the prediction path really carried tracks from frame to frame
