"""Measure the BASELINE.json configs that are not bench.py's headline workload:
C1 (CPU restatement of ObserveImage with ORB on synthetic stereo images), C2 (2000x2000
single pair), C3 (stereo + triangulation, 2000 features, KITTI rig), C5 sample (20000
features x 32-frame window).  GPU numbers next to the OpenCV CPU path on the same arrays.
Writes gpurun_out/configs.json."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import synth
import vision_slam_frontend_b200 as vsf
from oracle import cv2_ref, restate

RATIO = restate.NN_MATCH_RATIO
out = {"host": {"cpu_count": os.cpu_count(), "cv2": cv2_ref.version() if cv2_ref.available() else None}}


def best_of(fn, n=5, warm=1):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), float(np.median(ts))


def cpu_knn(Q, T, threads):
    cv2_ref.set_threads(threads)
    return best_of(lambda: cv2_ref.knn_match_raw(Q, T), n=5)


stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)

# ------------------------------------------------------------------ C2
Q, T = synth.descriptor_pair(2000, 2000, seed=0)
with vsf.Context(max_features=2048, window=2) as ctx:
    ctx.set_stream(stream.cuda_stream)
    lat = best_of(lambda: ctx.get_matches(Q, T, RATIO), n=50, warm=5)
    dq = torch.from_numpy(Q).cuda(); dt = torch.from_numpy(T).cuda()
    qp = (C.c_void_p * 1)(dq.data_ptr()); nn = (C.c_int * 1)(2000)
    def launch():
        assert ctx._L.vsf_window_match_device(ctx._h, qp, nn, 1, C.c_void_p(dt.data_ptr()), 2000, RATIO) == 0
    for _ in range(20):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(500):
        launch()
    e1.record(stream); torch.cuda.synchronize()
    b2b = e0.elapsed_time(e1) / 500 * 1e-3
    # CUDA graph of one launch, replayed
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        launch()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(500):
        g.replay()
    e1.record(stream); torch.cuda.synchronize()
    gr = e0.elapsed_time(e1) / 500 * 1e-3
c1t, _ = cpu_knn(Q, T, 1)
cat, _ = cpu_knn(Q, T, os.cpu_count())
out["C2"] = {"shape": "2000x2000x256b single pair", "comparisons": 4e6,
             "gpu_host_api_latency_s": lat[0], "gpu_host_api_latency_median_s": lat[1],
             "gpu_back_to_back_s": b2b, "gpu_graph_replay_s": gr,
             "gpu_cmp_per_s_back_to_back": 4e6 / b2b, "gpu_cmp_per_s_graph": 4e6 / gr,
             "gpu_cmp_per_s_host_api": 4e6 / lat[0],
             "cpu_1thread_s": c1t, "cpu_all_threads_s": cat, "cpu_threads": os.cpu_count(),
             "cpu_cmp_per_s_all_threads": 4e6 / cat}
print(json.dumps(out["C2"]), flush=True)

# ------------------------------------------------------------------ C3
P1, P2 = synth.kitti_projections()
F = synth.kitti_fundamental()
kl, dl, kr, dr, X, perm = synth.stereo_frame(2000, seed=1)
with vsf.Context(max_features=2048, window=2) as ctx:
    def frame():
        ctx.window_clear()
        return ctx.observe_features(0, kl, dl, kr, dr, F, P1, P2, RATIO)
    tf = best_of(frame, n=30, warm=3)
    got = frame()
    ok = perm >= 0
    x1 = np.stack([kl["x"][ok], kl["y"][ok]], 1); x2 = np.stack([kr["x"][perm[ok]], kr["y"][perm[ok]]], 1)
    tt = best_of(lambda: ctx.triangulate(P1, P2, x1, x2), n=30, warm=3)
cv2_ref.set_threads(os.cpu_count())
def cpu_frame():
    sm = cv2_ref.knn_match_raw(dl, dr)
    r = restate.FrontendOracle(P1, P2, F, frame_life=1)
    return sm
fo = restate.FrontendOracle(P1, P2, F, frame_life=1)
res = fo.observe_features(kl, dl, kr, dr)
Lc, Rc = res.left, res.right
def cpu_c3():
    cv2_ref.knn_match_raw(dl, dr)                                   # stereo L->R
    cv2_ref.knn_match_raw(Rc.descriptors, Lc.descriptors)           # R'->L'
    cv2_ref.triangulate_points(P1, P2, x1[:len(res.points)], x2[:len(res.points)])
tc = best_of(cpu_c3, n=5)
ttc = best_of(lambda: cv2_ref.triangulate_points(P1, P2, x1, x2), n=5)
out["C3"] = {"shape": "stereo L/R match + filter + R'->L' match + triangulation, 2000 features, KITTI 1241x376 rig",
             "gpu_observe_features_s": tf[0], "gpu_observe_features_median_s": tf[1],
             "kept_pairs": int(len(got["kept_left"])), "triangulated": int(len(got["tri_matches"])),
             "gpu_triangulate_only_s": tt[0], "points": int(len(x1)),
             "gpu_triangulated_points_per_s_host_api": len(x1) / tt[0],
             "cpu_knn_x2_plus_triangulate_s": tc[0], "cpu_triangulate_only_s": ttc[0],
             "cpu_triangulated_points_per_s": len(x1) / ttc[0], "cpu_threads": os.cpu_count()}
print(json.dumps(out["C3"]), flush=True)

# ------------------------------------------------------------------ C5 sample
n, W = 20000, 32
with vsf.Context(max_features=n, window=W) as ctx:
    ctx.set_stream(stream.cuda_stream)
    poses = 4
    seq = torch.empty((W + poses, n, 32), dtype=torch.uint8, device="cuda")
    ctx.synth_sequence_device(seq.data_ptr(), n, 0, W + poses, n // 10, 7)
    base, fb = seq.data_ptr(), n * 32
    def step(t):
        qp = (C.c_void_p * W)(*[base + (t + j) * fb for j in range(W)])
        nn = (C.c_int * W)(*([n] * W))
        assert ctx._L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * fb), n, RATIO) == 0
    step(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(poses):
        step(t)
    e1.record(stream); torch.cuda.synchronize()
    per_pose = e0.elapsed_time(e1) / poses * 1e-3
    host = seq[:2].cpu().numpy()
cv2_ref.set_threads(os.cpu_count())
tcp = best_of(lambda: cv2_ref.knn_match_raw(host[0], host[1]), n=2, warm=0)
out["C5_sample"] = {"shape": "20000 features x 32-frame window, %d poses timed" % poses,
                    "comparisons_per_pose": W * n * n, "gpu_s_per_pose": per_pose,
                    "gpu_cmp_per_s": W * n * n / per_pose,
                    "cpu_one_pair_s": tcp[0], "cpu_cmp_per_s": n * n / tcp[0],
                    "cpu_s_per_pose_extrapolated": tcp[0] * W, "cpu_threads": os.cpu_count(),
                    "poses_100k_on_8_gpus_hours": 100000 / 8 * per_pose / 3600}
print(json.dumps(out["C5_sample"]), flush=True)

# ------------------------------------------------------------------ C1 (CPU restatement, ORB)
try:
    import cv2
    rng = np.random.default_rng(0)
    big = cv2.GaussianBlur((rng.random((600, 2400)) * 255).astype(np.uint8), (0, 0), 1.2)
    big = cv2.resize(big, None, fx=2.0, fy=2.0, interpolation=cv2.INTER_CUBIC)
    orb = cv2.ORB_create(1000, 1.04, 50, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)
    t_ext = t_match = t_tri = 0.0
    prev = None
    poses = 100
    cv2.setNumThreads(os.cpu_count())
    for p in range(poses):
        x0 = 40 + 20 * p
        left = np.ascontiguousarray(big[200:680, x0:x0 + 640]); right = np.ascontiguousarray(big[200:680, x0 + 12:x0 + 652])
        t0 = time.perf_counter()
        kpl, dL = orb.detectAndCompute(left, None); kpr, dR = orb.detectAndCompute(right, None)
        t1 = time.perf_counter()
        if dL is None or dR is None or len(dL) < 2 or len(dR) < 2:
            continue
        cv2_ref.knn_match_raw(dL, dR)
        if prev is not None:
            cv2_ref.knn_match_raw(prev, dL)
        cv2_ref.knn_match_raw(dR, dL)
        t2 = time.perf_counter()
        m = min(len(kpl), len(kpr))
        xl = np.array([k.pt for k in kpl[:m]], np.float32); xr = np.array([k.pt for k in kpr[:m]], np.float32)
        cv2_ref.triangulate_points(P1, P2, xl, xr)
        t3 = time.perf_counter()
        t_ext += t1 - t0; t_match += t2 - t1; t_tri += t3 - t2
        prev = dL
    out["C1_cpu_restatement"] = {
        "what": "Python-cv2 restatement of ObserveImage (src/slam_frontend.cc:400-472) on synthetic 640x480 stereo "
                "pairs, ORB 1000 features, window 1, %d poses; NOT the reference binary (ROS/OpenCV C++ absent)" % poses,
        "s_per_pose_extract": t_ext / poses, "s_per_pose_match": t_match / poses,
        "s_per_pose_triangulate": t_tri / poses, "features_last": int(len(dL)), "cpu_threads": os.cpu_count()}
    print(json.dumps(out["C1_cpu_restatement"]), flush=True)
except Exception as e:      # noqa
    out["C1_cpu_restatement"] = {"error": repr(e)}

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
