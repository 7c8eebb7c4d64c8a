"""tests/golden/golden_undistort_v1.npz: cv2.undistort (the reference's per-frame call, examples/dataset.hpp:232-236)
on the golden frame with the EuRoC cam0 calibration (/root/reference/configs/euroc_sensor.yaml:43-45 scaled to 320x240).
Run in the authoring container:  python tests/golden/make_golden_undistort.py"""
import hashlib, os
import cv2
import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(here, "golden_v1.npz"))
f0 = G["f0"]
H, W = f0.shape
s = W / 752.0
K = np.array([[458.654 * s, 0, 367.215 * s], [0, 457.296 * s, 248.375 * H / 480.0], [0, 0, 1]], np.float32)
D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
out = cv2.undistort(f0, K, D)
np.savez_compressed(os.path.join(here, "golden_undistort_v1.npz"), K=K, D=D,
                    undistorted_sha=np.array(hashlib.sha256(out.tobytes()).hexdigest()), rows=out[::30].copy(),
                    cv2_version=np.array(cv2.__version__))
print("ok", out.shape, int((out == 0).sum()), "black border px")
