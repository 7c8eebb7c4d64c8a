"""Per-kernel device time of the single-stream plugin path (n = 1), serialised with CUDA events:
python scripts/latency_breakdown.py [frames]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rd_vio_b200 import _native as N
from rd_vio_b200.frontend import FrontEnd, GpuImage
from rd_vio_b200.synthetic import SyntheticStream

nfr = int(sys.argv[1]) if len(sys.argv) > 1 else 60
st = SyntheticStream(5, 752, 480, period=64)
frames = [st.frame(k) for k in range(nfr)]
L = N.lib()
with FrontEnd(752, 480, 3, 21, num_slots=4, max_points=1024) as fe:
    GpuImage.reset_frozen_parameters()
    last, kp = None, np.zeros((0, 2))
    lat = []
    for i, f in enumerate(frames):
        if i == 10:
            N.check(L.rdfe_profile_enable(fe.handle, 1), "enable")
        t0 = time.perf_counter()
        img = GpuImage(fe, f, t=0.05 * i)
        img.preprocess(6.0, 8, 8)
        if last is not None:
            nxt, s = last.track_keypoints(img, kp, None)
            kp = nxt[s != 0]
            last.release_image_buffer()
        kp = img.detect_keypoints(kp, 150, 20.0)
        lat.append(time.perf_counter() - t0)
        last = img
    nk = L.rdfe_profile_num_kernels()
    ms = (C.c_double * nk)(); cnt = (C.c_int64 * nk)()
    N.check(L.rdfe_profile_collect(fe.handle, ms, cnt), "collect")
    n = nfr - 10
    print("python-plugin wall latency per frame: median %.1f us" % (1e6 * float(np.median(lat[10:]))))
    tot = 0.0
    for k in range(nk):
        if cnt[k]:
            print("  %-16s %7.1f us/frame (%d launches/frame)" % (L.rdfe_profile_kernel_name(k).decode(), 1e3 * ms[k] / n, cnt[k] // n))
            tot += 1e3 * ms[k] / n
    print("  kernels total   %7.1f us/frame; tracked points at end %d" % (tot, len(kp)))
