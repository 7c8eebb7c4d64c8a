"""A/B of the LK template cache on the LK kernel alone: 64 streams, track(A -> B) leaves templates, then track(B -> C) on the
carried points is timed with the cache off and on (CUDA events; run under ncu for instruction counts).
  python scripts/lk_cache_ab.py [reps]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rd_vio_b200 import _native as N, workload as WL
from rd_vio_b200.frontend import FrontEnd


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    S, W, H, NP, T = 64, 752, 480, 150, 8
    WL.ensure_rings(list(range(S)), W, H, T)
    L = N.lib()
    ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
    fe = FrontEnd(W, H, 3, 21, num_slots=3 * S, max_points=512, stream=ts.cuda_stream)
    h = fe.handle
    sets = [np.array([fe.acquire() for _ in range(S)], np.int32) for _ in range(3)]
    frames = torch.from_numpy(np.stack([WL.load_ring(s, W, H, T)[:3] for s in range(S)])).cuda()     # [S][3][H][W]
    vp = lambda t: C.c_void_p(t.data_ptr())
    for k in range(3):
        ptrs = (C.c_void_p * S)(*[frames[i, k].data_ptr() for i in range(S)])
        N.check(L.rdfe_preprocess_batch_dev(h, sets[k].ctypes.data, S, ptrs, W, 6.0, 8, 8), "pre")
    stride = 512
    dp, tp = fe.detect_params(max_points=NP), fe.track_params(has_prediction=0)
    xy0 = torch.zeros((S, stride, 2), dtype=torch.float64, device="cuda"); c0 = torch.zeros(S, dtype=torch.int32, device="cuda")
    N.check(L.rdfe_detect_batch_dev(h, sets[0].ctypes.data, S, C.byref(dp), vp(xy0), vp(c0), stride, None, None, None), "detect")
    for mode in (0, 1):
        N.check(L.rdfe_set_template_cache(h, mode), "cache")
        xy1 = xy0.clone(); st1 = torch.zeros((S, stride), dtype=torch.int8, device="cuda")
        N.check(L.rdfe_track_batch_dev(h, sets[0].ctypes.data, sets[1].ctypes.data, S, C.byref(tp), vp(xy0), vp(xy1), vp(c0), stride, vp(st1)), "track01")
        fe.sync()
        # carried = tracked points compacted (host side, untimed)
        x1, s1, n0 = xy1.cpu().numpy(), st1.cpu().numpy(), c0.cpu().numpy()
        car = np.zeros((S, stride, 2)); cc = np.zeros(S, np.int32)
        for i in range(S):
            p = x1[i, :n0[i]][s1[i, :n0[i]] != 0]
            car[i, :len(p)] = p; cc[i] = len(p)
        car_d, cc_d = torch.from_numpy(car).cuda(), torch.from_numpy(cc).cuda()
        out = car_d.clone(); st2 = torch.zeros((S, stride), dtype=torch.int8, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for r in range(reps + 3):
            if r == 3:
                e0.record()
            N.check(L.rdfe_track_batch_dev(h, sets[1].ctypes.data, sets[2].ctypes.data, S, C.byref(tp), vp(car_d), vp(out), vp(cc_d), stride, vp(st2)), "track12")
        e1.record(); fe.sync()
        lk_, hit_ = C.c_ulonglong(0), C.c_ulonglong(0)
        N.check(L.rdfe_template_cache_stats(h, C.byref(lk_), C.byref(hit_), 1), "stats")
        print(f"cache {'on ' if mode else 'off'}: {1e3 * e0.elapsed_time(e1) / reps:8.1f} us per LK launch, {int(cc.sum())} points, "
              f"status ok {int(st2.sum().item())}, lookups {lk_.value} hits {hit_.value}", flush=True)
    fe.close()


if __name__ == "__main__":      # WL.ensure_rings spawns worker processes that re-import this file
    main()
