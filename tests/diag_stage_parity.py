"""Diagnostic (run by hand on a GPU box, not collected by pytest): per-stage parity statistics against the oracle + rough timings."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fe_oracle as orc
from rd_vio_b200.frontend import FrontEnd
from rd_vio_b200.synthetic import SyntheticStream

def main():
    s = SyntheticStream(0)
    f = [s.frame(k) for k in range(4)]
    fe = FrontEnd(752, 480, 3, 21, num_slots=8, max_points=512)
    print("levels", fe.nlevels, [fe.level_size(l) for l in range(fe.nlevels)])
    sl = [fe.acquire() for _ in range(4)]
    t = time.time(); fe.preprocess(sl, f); print("preprocess 4 frames (first call) %.1f ms" % ((time.time()-t)*1e3))
    t = time.time(); fe.preprocess(sl, f); print("preprocess 4 frames %.2f ms" % ((time.time()-t)*1e3))
    for i in range(2):
        ref, ref_lut = orc.clahe(f[i], return_lut=True)
        lut = fe.download_clahe_lut(i, 64)
        got = fe.download_level(sl[i], 0, 0)
        print(f"frame {i}: LUT neq {(lut != ref_lut).sum()}  CLAHE neq {(got != ref).sum()} / {got.size}")
        P = orc.Pyramid(ref, 21, 3)
        for l in range(fe.nlevels):
            im, dv, halo = fe.download_level(sl[i], l, 0), fe.download_level(sl[i], l, 1), fe.download_level(sl[i], l, 2)
            want = np.pad(P.image(l), 21, mode="reflect")
            print(f"  level {l}: img neq {(im != P.image(l)).sum()}  deriv neq {(dv != P.deriv(l)).sum()}  halo neq {(halo != want).sum()}")
    pre = orc.clahe(f[0])
    for fma in (0, 1):
        R = fe.harris_response(sl[0], harris_fma=fma); Rr = orc.harris(pre, 0.04, fma)
        print(f"harris fma={fma}: neq {(R != Rr).sum()} max|d| {np.abs(R-Rr).max():.3e}")
    ref, gxy_r, gre_r = orc.detect_keypoints(pre, np.zeros((0, 2)), 150, 20.0)
    t = time.time(); got, gxy, gre = fe.detect([sl[0]], [np.zeros((0, 2))], 150, 20.0, return_gftt=True); dt = time.time()-t
    print("detect: n got %d ref %d equal %s (gftt equal %s) %.2f ms" % (len(got[0]), len(ref), np.array_equal(got[0], ref), np.array_equal(gxy[0], gxy_r), dt*1e3))
    if not np.array_equal(gxy[0], gxy_r):
        n = min(len(gxy[0]), len(gxy_r)); d = np.nonzero((gxy[0][:n] != gxy_r[:n]).any(1))[0]
        print("  first diffs at", d[:10], gxy[0][d[:3]], gxy_r[d[:3]])
    pts = ref
    pred = s.predict(0, pts)
    PA, PB = orc.Pyramid(pre, 21, 3), orc.Pyramid(orc.clahe(f[1]), 21, 3)
    rxy, rst, _ = orc.track_keypoints(PA, PB, pts, pred)
    t = time.time(); gxy2, gst = fe.track([sl[0]], [sl[1]], [pts], [pred]); dt = time.time()-t
    gxy2, gst = gxy2[0], gst[0]
    both = (gst != 0) & (rst != 0)
    print("track: status agree %.4f  n_ok got %d ref %d  max err %.3e  median err %.3e  %.2f ms" % (
        (gst == rst).mean(), gst.sum(), rst.sum(), np.abs(gxy2[both]-rxy[both]).max() if both.any() else -1,
        np.median(np.abs(gxy2[both]-rxy[both])) if both.any() else -1, dt*1e3))
    # batch of 64 timing through the host API
    fe.close()
    fe = FrontEnd(752, 480, 3, 21, num_slots=128, max_points=512)
    A = [fe.acquire() for _ in range(64)]; B = [fe.acquire() for _ in range(64)]
    imgsA = [f[0]] * 64; imgsB = [f[1]] * 64
    fe.preprocess(A, imgsA); fe.preprocess(B, imgsB)
    for rep in range(3):
        t = time.time(); fe.preprocess(B, imgsB); t1 = time.time()-t
        t = time.time(); k = fe.detect(B, [pts] * 64, 150, 20.0); t2 = time.time()-t
        t = time.time(); n_, s_ = fe.track(A, B, [pts] * 64, [pred] * 64); t3 = time.time()-t
        print("host-API batch64: preprocess %.2f ms detect %.2f ms track %.2f ms -> %.0f frames/s (incl. python+copies)" % (
            t1*1e3, t2*1e3, t3*1e3, 64/(t1+t2+t3)))
    print("launches", fe.kernel_launches())
    fe.close()

if __name__ == "__main__":
    main()
