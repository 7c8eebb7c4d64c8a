"""Diagnostic soak (run by hand on a GPU box, not collected by pytest): lock-step replay, GPU vs oracle, every
stage of every frame compared bit for bit; reports the frames whose outputs differ.
  python tests/diag_replay.py [--workload euroc|advio|hd] [--stream S] [--frames N] [--pred none|noisy|bad]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fe_oracle as orc
from rd_vio_b200.frontend import FrontEnd
from rd_vio_b200.synthetic import SyntheticStream
from rd_vio_b200.workload import WORKLOADS

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="euroc")
ap.add_argument("--stream", type=int, default=3)
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--pred", default="none")
ap.add_argument("--skip", type=int, default=1, help="frame stride (larger = larger motion)")
ap.add_argument("--cache", action="store_true", help="LK template cache on (must not change a single result)")
a = ap.parse_args()
wl = WORKLOADS[a.workload]
W, H, NP, ML, WIN = wl["width"], wl["height"], wl["points"], wl["max_level"], wl["win"]
st = SyntheticStream(a.stream, W, H, period=200)
bad = 0
tag = f"[{a.workload} s{a.stream} {a.pred} skip{a.skip}{' cache' if a.cache else ''}]"
with FrontEnd(W, H, ML, WIN, num_slots=4, max_points=4 * NP + 64) as fe:
    fe.set_template_cache(a.cache)
    last_pyr, last_slot, last_kp = None, None, None
    for i in range(a.frames):
        f = st.frame(i * a.skip)
        slot = fe.acquire()
        fe.preprocess([slot], [f])
        pre = orc.clahe(f)
        pyr = orc.Pyramid(pre, WIN, ML)
        g0 = fe.download_level(slot, 0, 0)
        if not np.array_equal(g0, pre):
            print(tag, f"frame {i}: CLAHE differs at {np.argwhere(g0 != pre)[:5]}"); bad += 1
        kp = np.zeros((0, 2))
        if last_pyr is not None:
            pred = None
            if a.pred == "noisy":
                pred = st.predict((i - 1) * a.skip, last_kp)
            elif a.pred == "bad":
                pred = last_kp + np.random.default_rng(i).normal(0, 6.0, last_kp.shape)
            nxt, stt, _ = orc.track_keypoints(last_pyr, pyr, last_kp, pred, WIN, ML)
            gn, gs = fe.track([last_slot], [slot], [last_kp], [pred] if pred is not None else None)
            gn, gs = gn[0], gs[0]
            if not np.array_equal(gs != 0, stt != 0):
                d = np.nonzero((gs != 0) != (stt != 0))[0]
                print(tag, f"frame {i}: status differs at {d}: gpu {gs[d]} oracle {stt[d]} pts {last_kp[d]} gpu_next {gn[d]} orc_next {nxt[d]}"); bad += 1
            ok = (stt != 0) & (gs != 0)
            if ok.any() and np.abs(gn[ok] - nxt[ok]).max() > 0:
                print(tag, f"frame {i}: tracked positions differ by {np.abs(gn[ok] - nxt[ok]).max()}"); bad += 1
            kp = nxt[stt != 0]
            fe.release(last_slot)
        okp, ogx, ogr = orc.detect_keypoints(pre, kp, NP, 20.0)
        gkp, ggx, ggr = fe.detect([slot], [kp], NP, 20.0, return_gftt=True)
        gkp, ggx, ggr = gkp[0], ggx[0], ggr[0]
        if gkp.shape != okp.shape or np.abs(gkp - okp).max() > 0:
            bad += 1
            print(tag, f"frame {i}: detect differs: gpu {gkp.shape} oracle {okp.shape}; existing {len(kp)}")
            R = fe.harris_response(slot)
            Ro = orc.harris(pre, 0.04, 0)
            dr = np.argwhere(R != Ro)
            print(f"   harris response differs at {len(dr)} px", dr[:5], (R[R != Ro][:5], Ro[R != Ro][:5]))
            og = orc.gftt_select(Ro, NP, 1e-3, 20.0)
            print("   gftt gpu n", len(ggx), "oracle n", len(og[0]))
            m = min(len(ggx), len(og[0]))
            dd = np.nonzero(np.any(ggx[:m] != og[0][:m], axis=1))[0]
            print("   first gftt diff idx", dd[:5], "gpu", ggx[dd[:3]], ggr[dd[:3]], "orc", og[0][dd[:3]], og[1][dd[:3]])
            if os.path.isdir("gpurun_out"):
                np.savez_compressed(os.path.join("gpurun_out", f"diag_{a.workload}_s{a.stream}_f{i}.npz"), frame=f, kp=kp, gkp=gkp,
                                    okp=okp, ggx=ggx, ggr=ggr, ogx=og[0], ogr=og[1])
            if bad > 6:
                break
        last_pyr, last_slot, last_kp = pyr, slot, okp
print(tag, "done; frames", a.frames, "tracked at end", len(last_kp), "mismatching checks:", bad)
