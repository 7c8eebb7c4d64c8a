"""Disagreement report against the reference's DEFAULT OpenCV mode (north_star: "any disagreement is reported").

The parity target of the CUDA path is OpenCV's plain C++ float order (cv2.setUseOptimized(False)), where detect is
identical and LK agrees to < 0.01 px.  A deployed rd_vio runs OpenCV with its SIMD dispatch ON (FMA in Sobel, a
re-associated Harris formula, float32 lane-striped LK sums).  This script counts, frame by frame at each BASELINE
shape, how the GPU results differ from cv2 with setUseOptimized(True):

  detect : keypoints in the symmetric difference of the two sets, for (i) the plain GPU path (harris_fma = 0) and
           (ii) the GPU path with OpenCV's FMA placement (harris_fma = 1);
  track  : status flags that differ and the largest position difference among points both sides accept
           (the same keypoints and IMU-style predictions are fed to both sides).

Run on a GPU box:  python scripts/disagreement.py [--frames 200] [--out profiles/r2_disagreement.json]
The reference calls are the ones OpenCvImage makes (oracle/cv2_reference.py wraps exactly those, opencv_image.cpp:38-161).
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
from oracle.cv2_reference import Cv2Image
from rd_vio_b200.frontend import FrontEnd
from rd_vio_b200.synthetic import SyntheticStream
from rd_vio_b200.workload import WORKLOADS

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--hd-frames", type=int, default=100)
ap.add_argument("--out", default="profiles/r2_disagreement.json")
ap.add_argument("--workloads", default="euroc,advio,hd")
a = ap.parse_args()


def ordered(v):
    """float32 -> integer that is monotone in the value (so that differences count ulps across the sign change too)"""
    b = int(np.float32(v).view(np.int32))
    return b if b >= 0 else -(b & 0x7FFFFFFF)


def setdiff(p, q):
    ps = {(float(x), float(y)) for x, y in p}
    qs = {(float(x), float(y)) for x, y in q}
    return len(ps ^ qs)


cv2.setUseOptimized(True)
report = {"opencv": cv2.__version__, "use_optimized": bool(cv2.useOptimized()), "workloads": {}}
for name in a.workloads.split(","):
    wl = WORKLOADS[name]
    W, H, NP, ML, WIN = wl["width"], wl["height"], wl["points"], wl["max_level"], wl["win"]
    F = a.hd_frames if name == "hd" else a.frames
    st = SyntheticStream(5, W, H, period=max(F + 1, 16))
    acc = dict(frames=0, keypoints=0, detect_plain_diff=0, detect_fma_diff=0, frames_detect_plain_differ=0,
               frames_detect_fma_differ=0, tracked=0, status_diff=0, max_pos_diff=0.0,
               detect_diff_vs_plain_cv2=0, harris_px_diff_vs_plain_cv2=0, harris_max_ulp_vs_plain_cv2=0, harris_px=0)
    t0 = time.time()
    with FrontEnd(W, H, ML, WIN, num_slots=3, max_points=2 * NP + 64) as fe:
        prev = None
        for k in range(F + 1):
            f = st.frame(k)
            slot = fe.acquire()
            fe.preprocess([slot], [f])
            im = Cv2Image(f, level_num=ML)
            im.WIN = WIN
            im.preprocess(6.0, 8, 8)
            if prev is not None:
                pslot, pim, pkp = prev
                pred = st.predict(k - 1, pkp)
                r_next, r_st = pim.track_keypoints(im, pkp, pred)
                g_next, g_st = fe.track([pslot], [slot], [pkp], [pred])
                g_next, g_st = g_next[0], g_st[0]
                acc["tracked"] += len(pkp)
                acc["status_diff"] += int(((g_st != 0) != (r_st != 0)).sum())
                ok = (g_st != 0) & (r_st != 0)
                if ok.any():
                    acc["max_pos_diff"] = max(acc["max_pos_diff"], float(np.abs(g_next[ok] - r_next[ok]).max()))
                fe.release(pslot)
            if k < F:
                ref_kp = im.detect_keypoints(np.zeros((0, 2)), NP, 20.0)
                kp0 = fe.detect([slot], [np.zeros((0, 2))], NP, 20.0)[0]
                kp1 = fe.detect([slot], [np.zeros((0, 2))], NP, 20.0, harris_fma=1)[0]
                d0, d1 = setdiff(kp0, ref_kp), setdiff(kp1, ref_kp)
                # the parity target itself: cv2 in its plain C++ float order (setUseOptimized(False))
                cv2.setUseOptimized(False)
                try:
                    plain_kp = im.detect_keypoints(np.zeros((0, 2)), NP, 20.0)
                    if k % 4 == 0:          # the response map, every 4th frame: cv2's sliding box sums leave rare 1-2 ulp residues
                        Rc = cv2.cornerHarris(im.image, 3, 3, 0.04)
                        Rg = fe.harris_response(slot)
                        dd = np.argwhere(Rc != Rg)
                        acc["harris_px"] += Rc.size
                        acc["harris_px_diff_vs_plain_cv2"] += len(dd)
                        for yy, xx in dd:
                            u = abs(ordered(Rc[yy, xx]) - ordered(Rg[yy, xx]))
                            acc["harris_max_ulp_vs_plain_cv2"] = max(acc["harris_max_ulp_vs_plain_cv2"], u)
                finally:
                    cv2.setUseOptimized(True)
                acc["detect_diff_vs_plain_cv2"] += setdiff(kp0, plain_kp)
                acc["frames"] += 1
                acc["keypoints"] += len(ref_kp)
                acc["detect_plain_diff"] += d0
                acc["detect_fma_diff"] += d1
                acc["frames_detect_plain_differ"] += int(d0 > 0)
                acc["frames_detect_fma_differ"] += int(d1 > 0)
                prev = (slot, im, ref_kp)
    acc["seconds"] = round(time.time() - t0, 1)
    acc["detect_plain_rate"] = acc["detect_plain_diff"] / max(2 * acc["keypoints"], 1)
    acc["detect_fma_rate"] = acc["detect_fma_diff"] / max(2 * acc["keypoints"], 1)
    acc["status_diff_rate"] = acc["status_diff"] / max(acc["tracked"], 1)
    acc["config"] = f"{W}x{H}, {NP} points, maxLevel {ML}, {WIN}x{WIN}"
    report["workloads"][name] = acc
    print(name, json.dumps(acc), flush=True)

e = report["workloads"].get("euroc", {})
report["summary"] = {
    "vs": f"cv2 {cv2.__version__} with setUseOptimized(True) (the mode a deployed rd_vio runs), synthetic streams, "
          "profiles/r2_disagreement.json",
    "euroc_detect_set_difference_rate_plain": e.get("detect_plain_rate"),
    "euroc_detect_set_difference_rate_harris_fma": e.get("detect_fma_rate"),
    "euroc_lk_status_difference_rate": e.get("status_diff_rate"),
    "euroc_lk_max_position_difference_px": e.get("max_pos_diff"),
    "frames": e.get("frames"),
    "euroc_detect_keypoints_differing_vs_plain_cv2": e.get("detect_diff_vs_plain_cv2"),
    "euroc_harris_map_pixels_differing_vs_plain_cv2": [e.get("harris_px_diff_vs_plain_cv2"), e.get("harris_px")],
}
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(report, open(a.out, "w"), indent=1)
md = a.out.replace(".json", ".md")
with open(md, "w") as fmd:
    fmd.write("# Disagreement vs the reference's default OpenCV mode (cv2 %s, setUseOptimized(True))\n\n" % cv2.__version__)
    fmd.write("Made by `scripts/disagreement.py` on a B200 box.  detect: keypoints in the symmetric difference of the GPU set and "
              "cv2's, as a fraction of all keypoints of both sets; track: same keypoints and predictions on both sides.\n\n")
    fmd.write("| workload | frames | keypoints | detect diff, plain GPU | detect diff, GPU `harris_fma=1` | frames with a detect diff (plain / fma) | tracked | status flags differing | max position diff (px) | vs cv2 PLAIN mode: keypoints differing / Harris-map pixels differing (max ulp) |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for name, r in report["workloads"].items():
        fmd.write("| %s (%s) | %d | %d | %d (%.3f %%) | %d (%.3f %%) | %d / %d | %d | %d (%.3f %%) | %.2e | %d / %d of %d (%d) |\n" % (
            name, r["config"], r["frames"], r["keypoints"], r["detect_plain_diff"], 100 * r["detect_plain_rate"],
            r["detect_fma_diff"], 100 * r["detect_fma_rate"], r["frames_detect_plain_differ"], r["frames_detect_fma_differ"],
            r["tracked"], r["status_diff"], 100 * r["status_diff_rate"], r["max_pos_diff"], r["detect_diff_vs_plain_cv2"],
            r["harris_px_diff_vs_plain_cv2"], r["harris_px"], r["harris_max_ulp_vs_plain_cv2"]))
print("wrote", a.out, md)
