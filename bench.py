#!/usr/bin/env python
"""bench.py -- frames/s of the rd_vio visual front-end step (BASELINE.json metric).

One step = one new frame for every camera stream of the batch:
    preprocess(new) + track_keypoints(prev -> new, forward+backward LK + gating) + detect_keypoints(new)
i.e. FeatureTracker::run's per-frame plugin calls (/root/reference/src/rdvio/src/feature_tracker.cpp:32-98).

Workload at N GPUs: BASELINE.json configs[1] per GPU -- 752x480, 64 independent streams, 150 carried
keypoints + 150 detect, maxLevel 3 (4 images), 21x21 window; streams are partitioned across ranks
(weak scaling, no data-path collective: streams share no state).

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun by the driver)
  python bench.py --impl reference ...                   the reference's OpenCV path on host cores

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`: same metric through the host-pointer C ABI (H2D of every frame + D2H of results in the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rd_vio_b200 import workload as WL  # noqa: E402

METRIC = "frames/s detect+KLT-track @752x480"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="euroc", choices=list(WL.WORKLOADS))
    ap.add_argument("--streams", type=int, default=64, help="streams per GPU")
    ap.add_argument("--ring", type=int, default=8, help="distinct frames per stream kept resident")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--undistort", action="store_true",
                    help="also run cv::undistort's remap in front of preprocess (SURVEY 8(f) rank 1; not the headline config)")
    ap.add_argument("--profile-steps", type=int, default=20, help="extra steps with per-kernel events")
    ap.add_argument("--timeline", default="", help="write a per-launch timeline of 8 pipelined steps to this JSON file")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short BASELINE configs[2] (ADVIO-shaped) and configs[3] (1080p / 31x31) measurements")
    ap.add_argument("--other-steps", type=int, default=30, help="timed steps of each other_configs measurement")
    ap.add_argument("--no-chained", action="store_true",
                    help="skip the chained measurement (keypoints carried on the device from step to step, LK template cache off/on)")
    ap.add_argument("--no-compaction", action="store_true",
                    help="round-1 step semantics (lost tracks stay in the keypoint list); default = the reference's (frame.cpp:160-170)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm
def run_cpu(args, wl, stream_ids, steps, warmup, budget_s=None, single_stream=False):
    """Times the reference's CPU path (oracle/) on all host cores. Returns dict."""
    from oracle.cpu_frontend import CpuFrontEndPool, pick_backend
    backend = pick_backend()
    cores = os.cpu_count() or 1
    pool = CpuFrontEndPool(stream_ids, wl, args.ring, backend, workers=cores)
    try:
        for t in range(warmup):
            pool.step(t)
        times, t = [], warmup
        t_begin = time.perf_counter()
        for _ in range(steps):
            times.append(pool.step(t)); t += 1
            if budget_s and time.perf_counter() - t_begin > budget_s and len(times) >= 2:
                break
        total = float(np.sum(times))
        single = None
        if single_stream:
            pool.close()      # free the cores before timing one stream alone
            from oracle.cpu_frontend import single_stream_latency
            single = single_stream_latency(wl, args.ring)
        return {"value": len(stream_ids) * len(times) / total, "unit": UNIT, "cores": pool.workers, "single_stream": single,
                "kind": "reference" if backend == "cv2" else "port",
                "sample": f"{len(times)} steps x {len(stream_ids)} streams = {len(times) * len(stream_ids)} frames of the same "
                          f"synthetic workload, {pool.workers} worker processes x 1 OpenCV thread "
                          f"({'cv2 ' + __import__('cv2').__version__ if backend == 'cv2' else 'C oracle port'}); "
                          "bias: through Python cv2 the LK calls take level-0 images, so OpenCV rebuilds 4 pyramids + 2 Scharr sets "
                          "per frame that the C++ reference would reuse (measured ~2.8 of 24 ms at 1 thread) and the CLAHE/GFTT "
                          "objects are constructed per call: the arm is <= ~12 % pessimistic",
                "ms_per_step": 1e3 * total / len(times), "steps": len(times)}
    finally:
        pool.close()


def main_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    stream_ids = list(range(args.streams))
    WL.ensure_rings(stream_ids, wl["width"], wl["height"], args.ring)
    res = run_cpu(args, wl, stream_ids, args.steps, max(args.warmup, 1), budget_s=150.0, single_stream=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": res["steps"], "warmup": max(args.warmup, 1), "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32/f32", "data": "synthetic",
        "config": workload_config(args, wl, 1),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "single_stream")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, wl, n_gpus):
    name = {"euroc": "BASELINE configs[1]", "advio": "BASELINE configs[2] (per-GPU share)", "hd": "BASELINE configs[3]"}.get(args.workload, args.workload)
    return {"workload": f"{name}: {wl['width']}x{wl['height']} u8, {args.streams} independent streams per GPU "
                        f"batched per launch, {wl['points']} carried keypoints + {wl['points']}-point detect, maxLevel "
                        f"{wl['max_level']} ({wl['max_level'] + 1} images), {wl['win']}x{wl['win']} LK window, CLAHE 6.0/8x8, "
                        f"Harris-GFTT q=1e-3 minDist 20, Poisson radius 20, LK (30, 0.01)",
            "streams_per_gpu": args.streams, "global_streams": args.streams * n_gpus, "ring_frames": args.ring,
            "parallelism": f"streams partitioned across {n_gpus} GPU(s), no collective",
            "undistort": bool(getattr(args, "undistort", False)),
            "step_semantics": "round-1 (lost tracks kept)" if getattr(args, "no_compaction", False) else
                              "reference: only status != 0 points are carried into detect (frame.cpp:160-170)",
            "l2": f"inputs larger than L2: {args.streams}x{args.ring} resident frames = "
                  f"{args.streams * args.ring * wl['width'] * wl['height'] / 1e6:.0f} MB cycled (L2 126 MB)"}


# --------------------------------------------------------------------------- GPU arm
CONFIG_NAMES = {"euroc": "BASELINE configs[1]", "advio": "BASELINE configs[2] (per-GPU share)", "hd": "BASELINE configs[3]"}


def ncu_counters(workload, S):
    """Per-launch counters of the committed ncu launch list of this workload (profiles/r2_traffic.json, made by
    scripts/make_traffic.py): DRAM bytes, warp instructions, shared-memory wavefronts.  None when not captured."""
    if workload != "euroc" or S != 64:
        return {}
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))["kernels"]
        except Exception:
            continue
    return {}


def measure_gpu(args, env, wl_name, S, T, steps, warmup, profile_steps, want_e2e, want_timeline="", want_chained=False):
    """One workload on this rank's GPU: device-resident value, per-kernel times, e2e + copy-only legs."""
    import torch
    from rd_vio_b200 import _native as N
    from rd_vio_b200 import parallel as PAR
    from rd_vio_b200.frontend import FrontEnd
    from rd_vio_b200.synthetic import SyntheticStream
    rank, world, local, dist = env["rank"], env["world"], env["local"], env["dist"]
    wl = WL.WORKLOADS[wl_name]
    stream_ids = PAR.partition_streams(rank, world, S)
    WL.ensure_rings(stream_ids, wl["width"], wl["height"], T, workers=max(1, (os.cpu_count() or 1) // max(world, 1)))
    L = N.lib()
    W, H, NP = wl["width"], wl["height"], wl["points"]
    stride = 2 * NP
    # a real (non-default) stream shared by torch (copies, events) and the library (kernels): the legacy
    # default stream has handle 0, which the C ABI reads as "create your own"
    # high priority: the small copies that hand each step its keypoints must not queue behind the large grids
    tstream = torch.cuda.Stream(priority=int(os.environ.get("BENCH_MAIN_PRIO", "-1")))
    torch.cuda.set_stream(tstream)
    # three slot sets in rotation (prev, new, next-new): lets the preprocess of step t+1 overlap step t
    NSETS = 3
    cstride = 1024 if want_chained else 0          # chained run: the list grows beyond 2 x points before it settles
    fe = FrontEnd(W, H, wl["max_level"], wl["win"], num_slots=NSETS * S, max_points=max(stride, 512, cstride), device=local,
                  stream=tstream.cuda_stream)
    h = fe.handle
    slots = [np.array([fe.acquire() for _ in range(S)], np.int32) for _ in range(NSETS)]
    slotsA = slots[0]
    N.check(L.rdfe_set_pipelining(h, 1), "set_pipelining")
    N.check(L.rdfe_set_step_compaction(h, 0 if args.no_compaction else 1), "set_step_compaction")
    if args.undistort:
        sc = W / 752.0
        fe.set_undistort(np.array([[458.654 * sc, 0, 367.215 * sc], [0, 457.296 * sc, 248.375 * H / 480.0], [0, 0, 1]], np.float32),
                         np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32))

    # ---- resident inputs: frames [S][T][H][W] in HBM (and pinned on the host for e2e)
    host_frames = torch.empty((S, T, H, W), dtype=torch.uint8, pin_memory=True)
    hf = host_frames.numpy()
    for i, sid in enumerate(stream_ids):
        hf[i] = WL.load_ring(sid, W, H, T)
    dev_frames = host_frames.cuda(non_blocking=True)
    torch.cuda.synchronize()

    def frame_ptrs(base_ptr, k):
        return (C.c_void_p * S)(*[base_ptr + ((i * T + k) * H * W) for i in range(S)])

    dptrs = [frame_ptrs(dev_frames.data_ptr(), k) for k in range(T)]
    hptrs = [frame_ptrs(host_frames.data_ptr(), k) for k in range(T)]

    dp = fe.detect_params(max_points=NP, keypoint_distance=20.0)
    tp = fe.track_params(has_prediction=1)

    def vp(t):
        return C.c_void_p(t.data_ptr())

    # ---- carried keypoints of every ring frame (our own detect, untimed) + IMU-style predictions
    curr_xy = torch.zeros((T, S, stride, 2), dtype=torch.float64, device="cuda")
    cnt = torch.zeros((T, S), dtype=torch.int32, device="cuda")
    for k in range(T):
        N.check(L.rdfe_preprocess_batch_dev(h, slotsA.ctypes.data, S, dptrs[k], W, 6.0, 8, 8), "preprocess")
        N.check(L.rdfe_detect_batch_dev(h, slotsA.ctypes.data, S, C.byref(dp), vp(curr_xy[k]), vp(cnt[k]), stride,
                                        None, None, None), "detect")
    fe.sync()
    curr_h, cnt_h = curr_xy.cpu().numpy(), cnt.cpu().numpy()
    pred_h = np.zeros_like(curr_h)
    for i, sid in enumerate(stream_ids):
        st = SyntheticStream(sid, W, H, period=T)
        for k in range(T):
            n = cnt_h[k, i]
            pred_h[k, i, :n] = st.predict(k, curr_h[k, i, :n])
    pred_xy = torch.from_numpy(pred_h).cuda()
    work_xy = torch.zeros((S, stride, 2), dtype=torch.float64, device="cuda")
    work_cnt = torch.zeros((S,), dtype=torch.int32, device="cuda")
    status = torch.zeros((S, stride), dtype=torch.int8, device="cuda")
    mean_pts = float(cnt_h.mean())

    def step_dev(t):
        k = t % T
        prev, new = slots[t % NSETS], slots[(t + 1) % NSETS]
        work_xy.copy_(pred_xy[k], non_blocking=True)
        work_cnt.copy_(cnt[k], non_blocking=True)
        N.check(L.rdfe_frontend_step_dev(h, prev.ctypes.data, new.ctypes.data, S, dptrs[(k + 1) % T], W, 6.0, 8, 8,
                                         C.byref(tp), vp(curr_xy[k]), vp(work_xy), vp(cnt[k]), vp(status),
                                         C.byref(dp), vp(work_cnt), stride), "frontend_step")

    # prime: frame 0 preprocessed into the "prev" slots of step 0
    N.check(L.rdfe_preprocess_batch_dev(h, slots[0].ctypes.data, S, dptrs[0], W, 6.0, 8, 8), "preprocess")
    fe.sync()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    t = 0
    sampler = ClockSampler(local)
    sampler.start()                      # clocks are sampled under load: warm-up + timed region
    w0 = time.perf_counter()
    for _ in range(max(warmup, 3)):
        step_dev(t); t += 1
    torch.cuda.synchronize()
    while time.perf_counter() - w0 < 0.8:   # extra untimed warm-up so nvidia-smi (100 ms period) sees the load
        for _ in range(20):
            step_dev(t); t += 1
        torch.cuda.synchronize()
    barrier()
    launches0 = fe.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step_dev(t); t += 1
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = fe.kernel_launches() - launches0
    # last step's LK flags, masked by that step's counts (entries beyond a stream's count are stale)
    k_last = (t - 1) % T
    live = torch.arange(stride, device="cuda")[None, :] < cnt[k_last][:, None]
    n_tracked = float((status.bool() & live).sum().item())
    tracked_ok = n_tracked / max(float(cnt_h[k_last].sum()), 1.0)
    new_per_frame = float(work_cnt.float().mean().item()) - n_tracked / S if not args.no_compaction else None
    fe.sync()   # also surfaces a candidate-buffer overflow
    # whole-job frames/s = frames of all ranks / MAX over ranks of the device time (rd_vio_b200/parallel.py)
    value, ms_max = PAR.aggregate_throughput(S * steps, ms, dist, torch.device("cuda", local))

    # ---- per-kernel device time (separate pass, events around every launch) -> roofline of the dominant kernel
    prof = None
    if rank == 0 and profile_steps > 0:
        N.check(L.rdfe_profile_enable(h, 1), "profile_enable")
        for _ in range(profile_steps):
            step_dev(t); t += 1
        nk = L.rdfe_profile_num_kernels()
        pms = (C.c_double * nk)(); pn = (C.c_int64 * nk)()
        N.check(L.rdfe_profile_collect(h, pms, pn), "profile_collect")
        N.check(L.rdfe_profile_enable(h, 0), "profile_enable")
        prof = {L.rdfe_profile_kernel_name(i).decode(): {"ms_per_step": pms[i] / profile_steps,
                                                        "launches_per_step": pn[i] / profile_steps,
                                                        "us_per_launch": 1e3 * pms[i] / max(pn[i], 1)} for i in range(nk)}
        prof = {k: v for k, v in prof.items() if v["launches_per_step"] > 0}
    elif profile_steps > 0:
        for _ in range(profile_steps):
            step_dev(t); t += 1
        fe.sync()

    if rank == 0 and want_timeline:
        N.check(L.rdfe_profile_enable(h, 2), "profile_enable")
        for _ in range(10):
            step_dev(t); t += 1
        cap = 1024
        kid = (C.c_int * cap)(); t0s = (C.c_float * cap)(); t1s = (C.c_float * cap)(); ncap = C.c_int(0)
        N.check(L.rdfe_profile_timeline(h, kid, t0s, t1s, cap, C.byref(ncap)), "profile_timeline")
        N.check(L.rdfe_profile_enable(h, 0), "profile_enable")
        with open(want_timeline, "w") as f:
            json.dump([{"kernel": L.rdfe_profile_kernel_name(kid[i]).decode(), "start_us": round(1e3 * t0s[i], 2),
                        "end_us": round(1e3 * t1s[i], 2)} for i in range(ncap.value)], f)

    # ---- chained: what FeatureTracker::run does frame after frame -- the tracked + newly detected keypoints of step t
    # ARE the carried keypoints of step t+1 (they never leave the device), predictions from the gyro rotation
    # (rdfe_predict_rotation_dev = frame.cpp:82-93).  Only here can the LK template cache act (the headline run above
    # restarts every step from the stored detections of its ring frame, as BASELINE's "150 carried" asks).
    chained = None
    if want_chained:
        Hs = np.zeros((T, S, 9))
        for i, sid in enumerate(stream_ids):
            st = SyntheticStream(sid, W, H, period=T)
            K = st.K(); Ki = np.linalg.inv(K)
            for k in range(T):
                Hs[k, i] = (K @ st.gyro_delta(k).T @ Ki).reshape(9)
        H_dev = torch.from_numpy(Hs).cuda()
        tpc = fe.track_params(has_prediction=1)
        chained = {"stride": cstride}
        for label, cache_on in (("cache_off", 0), ("cache_on", 1)):
            N.check(L.rdfe_set_template_cache(h, cache_on), "set_template_cache")
            bufs = [torch.zeros((S, cstride, 2), dtype=torch.float64, device="cuda") for _ in range(2)]
            cnts = [torch.zeros((S,), dtype=torch.int32, device="cuda") for _ in range(2)]
            cstatus = torch.zeros((S, cstride), dtype=torch.int8, device="cuda")
            tt = t - (t % T) + T                      # start on ring frame 0
            bufs[0][:, :stride] = curr_xy[0]
            cnts[0].copy_(cnt[0])
            N.check(L.rdfe_preprocess_batch_dev(h, slots[tt % NSETS].ctypes.data, S, dptrs[0], W, 6.0, 8, 8), "preprocess")
            fe.sync()
            hist = []

            def cstep(tt, a):
                k = tt % T
                prev, new = slots[tt % NSETS], slots[(tt + 1) % NSETS]
                cur, nxt, ccur, cnxt = bufs[a], bufs[a ^ 1], cnts[a], cnts[a ^ 1]
                N.check(L.rdfe_predict_rotation_dev(h, S, vp(H_dev[k]), vp(cur), vp(ccur), cstride, vp(nxt)), "predict")
                cnxt.copy_(ccur, non_blocking=True)
                N.check(L.rdfe_frontend_step_dev(h, prev.ctypes.data, new.ctypes.data, S, dptrs[(k + 1) % T], W, 6.0, 8, 8,
                                                 C.byref(tpc), vp(cur), vp(nxt), vp(ccur), vp(cstatus), C.byref(dp), vp(cnxt),
                                                 cstride), "frontend_step(chained)")

            a = 0
            for _ in range(3 * T):                     # settle: the keypoint population reaches its steady state
                cstep(tt, a); tt += 1; a ^= 1
            fe.sync()
            if cache_on:
                L.rdfe_template_cache_stats(h, None, None, 1)
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(steps):
                cstep(tt, a); tt += 1; a ^= 1
            c1.record()
            barrier()
            fe.sync()
            cval, cms = PAR.aggregate_throughput(S * steps, c0.elapsed_time(c1), dist, torch.device("cuda", local))
            res = {"value": cval, "unit": UNIT, "ms_per_step": cms / steps, "steps": steps,
                   "mean_keypoints_after_step": float(cnts[a].float().mean().item())}
            if cache_on:
                lk_, hit_ = C.c_ulonglong(0), C.c_ulonglong(0)
                N.check(L.rdfe_template_cache_stats(h, C.byref(lk_), C.byref(hit_), 1), "cache_stats")
                res["tracked_points_per_frame"] = lk_.value / max(S * steps, 1)
                res["cache_hit_frac"] = hit_.value / max(lk_.value, 1)
            if rank == 0:            # serialised per-kernel times of this mode (LK is the kernel the cache changes)
                N.check(L.rdfe_profile_enable(h, 1), "profile_enable")
                for _ in range(10):
                    cstep(tt, a); tt += 1; a ^= 1
                nk = L.rdfe_profile_num_kernels()
                pms = (C.c_double * nk)(); pn = (C.c_int64 * nk)()
                N.check(L.rdfe_profile_collect(h, pms, pn), "profile_collect")
                N.check(L.rdfe_profile_enable(h, 0), "profile_enable")
                res["kernels_us_per_launch"] = {L.rdfe_profile_kernel_name(i).decode(): round(1e3 * pms[i] / pn[i], 1)
                                                for i in range(nk) if pn[i] > 0}
            else:
                for _ in range(10):
                    cstep(tt, a); tt += 1; a ^= 1
                fe.sync()
            chained[label] = res
            t = tt
        N.check(L.rdfe_set_template_cache(h, 0), "set_template_cache")
        chained["how"] = ("device-resident, keypoints carried from step to step on the device (tracked survivors + new detections, "
                          "frame.cpp:160-170), predictions by rdfe_predict_rotation_dev from the synthetic gyro increment; "
                          "3 ring periods of settling before the timed steps; same kernels, streams and ring as `value`")
        # leave the slot rotation where the e2e leg expects it
        N.check(L.rdfe_preprocess_batch_dev(h, slots[t % NSETS].ctypes.data, S, dptrs[t % T], W, 6.0, 8, 8), "preprocess")
        fe.sync()

    # ---- end to end through the host-pointer C ABI: pinned host frames + keypoints in, results out, every step
    e2e = None
    if want_e2e:
        h_curr = torch.from_numpy(curr_h).pin_memory().numpy()
        h_pred = torch.from_numpy(pred_h).pin_memory().numpy()
        h_cnt = cnt_h.copy()
        h_next = torch.empty((S, stride, 2), dtype=torch.float64).pin_memory().numpy()
        h_status = torch.empty((S, stride), dtype=torch.int8).pin_memory().numpy()
        h_wcnt = np.zeros(S, np.int32)

        def submit(tt):
            k = tt % T
            prev, new = slots[tt % NSETS], slots[(tt + 1) % NSETS]
            tk = C.c_int()
            N.check(L.rdfe_frontend_step_submit(h, prev.ctypes.data, new.ctypes.data, S, hptrs[(k + 1) % T], W, 6.0, 8, 8,
                                                C.byref(tp), h_curr[k].ctypes.data, h_pred[k].ctypes.data,
                                                h_cnt[k].ctypes.data, C.byref(dp), stride, C.byref(tk)), "step_submit")
            return tk.value

        def wait(tk):
            N.check(L.rdfe_frontend_step_wait(h, tk, h_next.ctypes.data, h_wcnt.ctypes.data, h_status.ctypes.data), "step_wait")

        h2d = S * H * W + 2 * S * stride * 16 + 2 * S * 4
        d2h = S * stride * 16 + S * 4 + S * stride + 4
        # the end-to-end leg always times at least 100 steps: with the driver's small K the fill and drain of the three-deep
        # pipeline (about two steps) would otherwise weigh 10 % of the timed region
        e_steps = min(max(steps, 100), 400)
        def run_pipelined(nsteps, tt):       # up to three steps in flight (rdfe_frontend_step_submit's pipeline depth)
            pending = []
            for _ in range(nsteps):
                pending.append(submit(tt)); tt += 1
                if len(pending) == 3:
                    wait(pending.pop(0))
            while pending:
                wait(pending.pop(0))
            return tt

        t = run_pipelined(6, t)                  # warm the pipeline
        barrier()
        w0 = time.perf_counter()
        t = run_pipelined(e_steps, t)
        barrier()
        sec = time.perf_counter() - w0
        e_value, sec_max_ms = PAR.aggregate_throughput(S * e_steps, sec * 1e3, dist, torch.device("cuda", local))
        # copy-only leg: the same pinned frames through the same upload path (rdfe_upload_only: same staging, copy
        # stream and single 2-D copy as submit), no kernels -- what this box's host side delivers to N GPUs at once
        fe.sync()
        for tt in range(3):
            N.check(L.rdfe_upload_only(h, slots[tt % NSETS].ctypes.data, S, hptrs[tt % T], W, 1), "upload_only")
        barrier()
        w0 = time.perf_counter()
        for tt in range(e_steps):
            N.check(L.rdfe_upload_only(h, slots[tt % NSETS].ctypes.data, S, hptrs[tt % T], W, 1 if tt == e_steps - 1 else 0),
                    "upload_only")
        barrier()
        csec = time.perf_counter() - w0
        c_value, _ = PAR.aggregate_throughput(S * e_steps, csec * 1e3, dist, torch.device("cuda", local))
        e2e = {"value": e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "steps": e_steps,
               "h2d_gbs_in_e2e": e_value * H * W / 1e9,
               "h2d_only": {"frames_per_s": c_value, "gbs_all_gpus": c_value * H * W / 1e9, "gbs_per_gpu": c_value * H * W / 1e9 / world,
                            "e2e_over_copy_only": e_value / c_value if c_value > 0 else None,
                            "how": "rdfe_upload_only: the frame uploads of the same steps alone (same pinned buffers, same "
                                   "staging, copy stream and one strided 2-D copy per step), no kernels, all ranks at once"},
               "how": "rdfe_frontend_step_submit/_wait (C ABI, pinned HOST buffers, three steps in flight): every step "
                      f"copies its {S} frames + carried keypoints + predictions H2D and its tracked/detected keypoints, "
                      "counts and status D2H; wall clock bracketed by barrier+synchronize, max over ranks"}

    out = {"workload": wl_name, "wl": wl, "S": S, "T": T, "steps": steps, "warmup": max(warmup, 3), "value": value, "ms_max": ms_max,
           "launches": int(launches), "clocks": clocks, "prof": prof, "e2e": e2e, "mean_pts": mean_pts, "tracked_ok": tracked_ok,
           "new_per_frame": new_per_frame, "stream_ids": stream_ids, "chained": chained}
    fe.close()
    del dev_frames, host_frames, curr_xy, pred_xy
    torch.cuda.empty_cache()
    return out


def annotate_kernels(m, hbm_peak, peak_src, sm_mhz):
    """roofline of the dominant kernel + every kernel against the same HBM peak; LK is charged for the points the
    run actually carried (mean_pts), not the nominal count."""
    wl, S = m["wl"], m["S"]
    W, H = wl["width"], wl["height"]
    total_nominal, _ = WL.algorithmic_bytes(W, H, wl["points"], wl["max_level"], wl["win"])
    total_bytes, stage_bytes = WL.algorithmic_bytes(W, H, m["mean_pts"], wl["max_level"], wl["win"])
    prof = m["prof"]
    counters = ncu_counters(m["workload"], S)
    roofline = None
    if prof:
        f_hz = (sm_mhz or 1965.0) * 1e6
        for kname, v in prof.items():
            sb = stage_bytes.get(kname, 0) * S
            v["algorithmic_bytes_per_step"] = sb
            v["achieved_gbs"] = (sb / (v["ms_per_step"] * 1e-3) / 1e9) if v["ms_per_step"] > 0 else 0.0
            v["hbm_frac"] = v["achieved_gbs"] / hbm_peak
            for cname, c in counters.items():
                if (cname.startswith(kname) or (kname == "poisson_append" and cname.startswith("poisson_"))) and v["us_per_launch"] > 0:
                    t_s = v["us_per_launch"] * 1e-6
                    if c.get("warp_inst_per_launch"):
                        # warp instructions / (4 schedulers x 148 SMs x clock x time): 1.0 = every issue slot used
                        v["issue_frac"] = c["warp_inst_per_launch"] / (4 * 148 * f_hz * t_s)
                    if c.get("smem_wavefronts_per_launch"):
                        # shared-memory wavefronts / (1 per cycle per SM)
                        v["smem_frac"] = c["smem_wavefronts_per_launch"] / (148 * f_hz * t_s)
                    v["dram_bytes_per_launch_ncu"] = c.get("dram_read_bytes_per_launch", 0) + c.get("dram_write_bytes_per_launch", 0)
        dom = max(prof, key=lambda kname: prof[kname]["ms_per_step"])
        d = prof[dom]
        per_launch_bytes = stage_bytes.get(dom, 0) * S / max(d["launches_per_step"], 1)
        ach = per_launch_bytes / (d["us_per_launch"] * 1e-6) / 1e9 if d["us_per_launch"] > 0 else 0.0
        bound = "hbm"
        roofline = {"kernel": dom, "bound": bound, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": d.get("dram_bytes_per_launch_ncu"), "peak_source": peak_src,
                    "issue_frac": d.get("issue_frac"), "smem_frac": d.get("smem_frac"),
                    "note": "the contract's bound is HBM (byte model of SURVEY.md 8(d)); the kernel itself is limited by "
                            "instruction issue / dependent-latency chains, not by DRAM: issue_frac = warp instructions / "
                            "(4 x 148 x SM clock x time) and smem_frac = shared-memory wavefronts / (148 x clock x time), "
                            "counters from the committed ncu launch list (profiles/r2_traffic.json)",
                    "algorithmic_bytes_per_launch": per_launch_bytes, "us_per_launch": d["us_per_launch"],
                    "carried_points_charged": m["mean_pts"],
                    "share_of_step": d["ms_per_step"] / max(sum(v["ms_per_step"] for v in prof.values()), 1e-12)}
    world = m.get("world", 1)
    step = {"algorithmic_bytes_per_frame": total_bytes, "algorithmic_bytes_per_frame_nominal_points": total_nominal,
            "frac_per_gpu": (m["value"] / world) * total_bytes / 1e9 / hbm_peak, "peak_gbs": hbm_peak, "peak_source": peak_src}
    return roofline, step


def main_b200(args, wl):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        # NCCL only reduces timing scalars here (streams share no state).  Its start-up banner goes to stdout,
        # which must carry exactly one JSON line: point fd 1 at stderr while the communicator comes up.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    env = {"rank": rank, "world": world, "local": local, "dist": dist}
    m = measure_gpu(args, env, args.workload, args.streams, args.ring, args.steps, args.warmup, args.profile_steps,
                    not args.no_e2e, args.timeline, want_chained=not args.no_chained and args.workload == "euroc")
    m["world"] = world

    # ---- BASELINE configs[2] / configs[3]: short measurements in the same process, every N (all ranks take part)
    others = {}
    if not args.no_other_configs and args.workload == "euroc":
        for name, S2, T2 in (("advio", 32, 5), ("hd", 32, 4)):
            try:
                mo = measure_gpu(args, env, name, S2, T2, args.other_steps, 3, 8, False)
                mo["world"] = world
                others[name] = mo
            except Exception as e:      # report, never hide; the headline line must still be printed
                others[name] = {"error": repr(e)}
                if dist is not None:
                    raise

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    hbm_peak, peak_src = peaks()
    roofline, step = annotate_kernels(m, hbm_peak, peak_src, m["clocks"].get("sm_mhz"))
    other_cfgs = {}
    for name, mo in others.items():
        if "error" in mo:
            other_cfgs[name] = mo
            continue
        r2, st2 = annotate_kernels(mo, hbm_peak, peak_src, mo["clocks"].get("sm_mhz"))
        w2 = mo["wl"]
        other_cfgs[name] = {
            "config": CONFIG_NAMES[name] + f": {w2['width']}x{w2['height']}, {mo['S']} streams per GPU, {w2['points']} keypoints, "
                      f"maxLevel {w2['max_level']}, {w2['win']}x{w2['win']} window, ring {mo['T']} frames (> L2)",
            "value": mo["value"], "unit": UNIT, "n_gpus": world, "steps": mo["steps"], "warmup": mo["warmup"],
            "ms_per_step": mo["ms_max"] / mo["steps"], "hbm_roofline_step": st2, "roofline": r2,
            "kernels_us_per_launch": {k: round(v["us_per_launch"], 1) for k, v in (mo["prof"] or {}).items()},
            "mean_carried_keypoints": mo["mean_pts"], "tracked_ok_frac": mo["tracked_ok"], "clocks": mo["clocks"]}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            cpu_streams = m["stream_ids"][:min(args.streams, 64)]
            r = run_cpu(args, wl, cpu_streams, steps=40, warmup=1, budget_s=20.0)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:   # report, never hide
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}

    parity_report = None
    try:    # committed disagreement table vs OpenCV's default (dispatched / FMA) mode, north_star's reporting clause
        parity_report = json.load(open(os.path.join(ROOT, "profiles", "r2_disagreement.json")))["summary"]
    except Exception:
        parity_report = None

    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": m["ms_max"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32/f32", "data": "synthetic",
        "config": workload_config(args, wl, world),
        "clocks": m["clocks"], "e2e": m["e2e"], "gpu_launches": m["launches"],
        "roofline": roofline, "cpu_baseline": cpu,
        "hbm_roofline_step": step,
        "kernels": m["prof"], "mean_carried_keypoints": m["mean_pts"], "tracked_ok_frac": m["tracked_ok"],
        "new_keypoints_per_frame": m["new_per_frame"],
        "chained": m["chained"],
        "other_configs": other_cfgs,
        "scope_note": "the step is the plugin's three calls; Frame::track_keypoints' two host RANSAC masks (frame.cpp:99-132, "
                      "SURVEY 8(f) rank 2) between track and detect stay on the host and are NOT in the timed region of either arm",
        "disagreement_vs_default_opencv": parity_report,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    wl = WL.WORKLOADS[args.workload]
    if args.impl == "reference":
        return main_reference(args, wl)
    return main_b200(args, wl)


if __name__ == "__main__":
    sys.exit(main())
