"""Diagnostic (run by hand on a GPU box): lock-step replay, GPU vs oracle, reporting the first frame whose
stage outputs differ.  python tests/diag_replay.py [stream] [frames]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fe_oracle as orc
from rd_vio_b200.frontend import FrontEnd
from rd_vio_b200.synthetic import SyntheticStream

sid = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 200
st = SyntheticStream(sid, 752, 480, period=200)
bad = 0
with FrontEnd(752, 480, 3, 21, num_slots=4, max_points=1024) as fe:
    last_pyr, last_slot, last_kp = None, None, None
    for i in range(nfr):
        f = st.frame(i)
        slot = fe.acquire()
        fe.preprocess([slot], [f])
        pre = orc.clahe(f)
        pyr = orc.Pyramid(pre, 21, 3)
        g0 = fe.download_level(slot, 0, 0)
        if not np.array_equal(g0, pre):
            print(f"frame {i}: CLAHE differs at {np.argwhere(g0 != pre)[:5]}"); bad += 1
        kp = np.zeros((0, 2))
        if last_pyr is not None:
            nxt, stt, _ = orc.track_keypoints(last_pyr, pyr, last_kp, None)
            gn, gs = fe.track([last_slot], [slot], [last_kp], None)
            gn, gs = gn[0], gs[0]
            if not np.array_equal(gs != 0, stt != 0):
                d = np.nonzero((gs != 0) != (stt != 0))[0]
                print(f"frame {i}: status differs at {d}: gpu {gs[d]} oracle {stt[d]} pts {last_kp[d]} gpu_next {gn[d]} orc_next {nxt[d]}"); bad += 1
            ok = (stt != 0) & (gs != 0)
            if ok.any() and np.abs(gn[ok] - nxt[ok]).max() > 0:
                print(f"frame {i}: tracked positions differ by {np.abs(gn[ok] - nxt[ok]).max()}"); bad += 1
            kp = nxt[stt != 0]
            fe.release(last_slot)
        okp, ogx, ogr = orc.detect_keypoints(pre, kp, 150, 20.0)
        gkp, ggx, ggr = fe.detect([slot], [kp], 150, 20.0, return_gftt=True)
        gkp, ggx, ggr = gkp[0], ggx[0], ggr[0]
        if gkp.shape != okp.shape or np.abs(gkp - okp).max() > 0:
            bad += 1
            print(f"frame {i}: detect differs: gpu {gkp.shape} oracle {okp.shape}; existing {len(kp)}")
            R = fe.harris_response(slot)
            Ro = orc.harris(pre, 0.04, 0)
            dr = np.argwhere(R != Ro)
            print(f"   harris response differs at {len(dr)} px", dr[:5], (R[R != Ro][:5], Ro[R != Ro][:5]))
            og = orc.gftt_select(Ro, 150, 1e-3, 20.0)
            print("   gftt gpu n", len(ggx), "oracle n", len(og[0]))
            m = min(len(ggx), len(og[0]))
            dd = np.nonzero(np.any(ggx[:m] != og[0][:m], axis=1))[0]
            print("   first gftt diff idx", dd[:5], "gpu", ggx[dd[:3]], ggr[dd[:3]], "orc", og[0][dd[:3]], og[1][dd[:3]])
            np.savez_compressed(os.path.join("gpurun_out", f"diag_frame{i}.npz"), frame=f, kp=kp, gkp=gkp, okp=okp, ggx=ggx,
                                ggr=ggr, ogx=og[0], ogr=og[1])
            if bad > 3:
                break
        last_pyr, last_slot, last_kp = pyr, slot, okp
print("done; mismatching checks:", bad)
