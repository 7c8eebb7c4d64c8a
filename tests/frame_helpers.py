"""Test helper: the reference's Image plugin interface on top of the CPU oracle (TEST INFRASTRUCTURE ONLY), so that
oracle.frame_host.FeatureTracker can be replayed once on the oracle and once on the GPU plugin."""
import numpy as np


class OracleImage:
    """OpenCvImage restated on oracle/fe_oracle (same members as rd_vio_b200.frontend.GpuImage)."""

    def __init__(self, image, t=0.0, win=21, max_level=3):
        self.image, self.t, self.win, self.max_level = np.ascontiguousarray(image, np.uint8), t, win, max_level
        self.pre = self.pyr = None

    def preprocess(self, clip, tx, ty):
        from oracle import fe_oracle as orc
        self.pre = orc.clahe(self.image, clip, tx, ty)
        self.pyr = orc.Pyramid(self.pre, self.win, self.max_level)

    def detect_keypoints(self, keypoints, max_points=1000, keypoint_distance=10.0):
        from oracle import fe_oracle as orc
        return orc.detect_keypoints(self.pre, np.asarray(keypoints, np.float64).reshape(-1, 2), max_points,
                                    keypoint_distance)[0]

    def track_keypoints(self, next_image, curr_keypoints, next_keypoints=None):
        from oracle import fe_oracle as orc
        curr = np.asarray(curr_keypoints, np.float64).reshape(-1, 2)
        if not isinstance(next_image, OracleImage) or next_image.pyr is None or self.pyr is None or len(curr) == 0:
            has = next_keypoints is not None and len(next_keypoints) > 0
            return (np.asarray(next_keypoints, np.float64).reshape(-1, 2).copy() if has else np.zeros_like(curr),
                    np.zeros(len(curr), np.int8))
        nxt, st, _ = orc.track_keypoints(self.pyr, next_image.pyr, curr, next_keypoints, self.win, self.max_level)
        return nxt, st

    def release_image_buffer(self):
        self.image = self.pre = self.pyr = None


def replay(stream, n_frames, make_image, start=0):
    """FeatureTracker::run over frames start..start+n_frames-1 of a synthetic stream; per frame (pixels, track ids)."""
    from oracle.frame_host import FeatureTracker, Frame, q_from_matrix
    ft, out = FeatureTracker(), []
    for i in range(n_frames):
        k = start + i
        dq = q_from_matrix(stream.gyro_delta(k - 1)) if i else None
        fr = Frame(make_image(stream.frame(k), 0.05 * k), stream.K(), frame_id=i, delta_q=dq)
        ft.track_frame(fr)
        out.append((fr.keypoints().copy(), [t.id if t is not None else -1 for t in fr.tracks]))
    return out
