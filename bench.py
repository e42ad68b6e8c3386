#!/usr/bin/env python
"""bench.py — Hamming comparisons/s of the frame-to-window matching hot path.

Workload (BASELINE.json configs[3], "C4"): a synthetic 10k-pose sequence, 5000 packed 256-bit
descriptors per pose, every pose matched (k=2 Hamming kNN + Lowe ratio test + ordered
compaction) against each of the 10 prior frames of the sliding window: 10 x 5000 x 5000 =
2.5e8 comparisons and 10 matched frame pairs per pose.  A STEP is one batch of 768 consecutive
poses (1.92e11 comparisons, 7680 frame pairs; 192 groups of four poses, three kernel launches
each), so the driver's `--steps 20` walks the 10k-pose sequence one and a half times and the
timed region lasts ~0.65 s (512 poses per step gave 0.45 s once a pose took 44 us).  With N GPUs
every rank walks its own pose range (weak scaling, no data-path collective); match lists are
gathered with NCCL after the timed region only.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--desc-bytes 32|61|64]
    python bench.py --impl reference ...      # OpenCV CPU path on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import synth  # noqa: E402  (tests/synth.py: seeded generators + numpy twin of the CUDA generator)

METRIC = "hamming_comparisons_per_sec"
UNIT = "cmp/s"
RATIO = float(np.float32(0.6))          # FrontendConfig::nn_match_ratio_ (slam_frontend.cc:555)
BEST_PERCENT = 0.3                      # FrontendConfig::best_percent_   (slam_frontend.cc:554)
SEED = 20240917
L2_BYTES = 126 * 1024 * 1024


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--features", type=int, default=5000)
    ap.add_argument("--window", type=int, default=10)
    ap.add_argument("--desc-bytes", type=int, default=32,
                    help="descriptor width: 32 = ORB (BASELINE's 256-bit shape), 61 = AKAZE (the reference's "
                         "default extractor, src/slam_frontend.cc:553), 64 = BRISK/FREAK")
    ap.add_argument("--poses-per-step", type=int, default=768, help="poses in one step (batch)")
    ap.add_argument("--stride", type=int, default=0, help="landmark stride per pose (default N/10)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads legs (C2, C3, C5, other width)")
    ap.add_argument("--e2e-lag", type=int, default=12,
                    help="frames submitted ahead of the one being collected in the pipelined e2e leg "
                         "(1 .. VSF_PIPELINE_DEPTH - 1)")
    ap.add_argument("--popc-mode", type=int, default=-1)
    ap.add_argument("--split", type=int, default=0)
    ap.add_argument("--qpt", type=int, default=0)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--engine", type=int, default=0,
                    help="0 auto (tensor cores on this workload), 1 POPC pipe, 2 tensor int8, 3 tensor e4m3")
    return ap.parse_args()


def workload_config(a):
    bits = 8 * a.desc_bytes
    B = a.poses_per_step
    return {
        "workload": "C4 frame-to-window matching: %d features x %d prior frames per pose, %d-bit descriptors, "
                    "k=2 Hamming kNN + ratio %.1f + ordered compaction; one step = a batch of %d consecutive "
                    "poses of the synthetic sequence" % (a.features, a.window, bits, 0.6, B),
        "features_per_frame": a.features, "window": a.window, "descriptor_bits": bits,
        "poses_per_step": B,
        "comparisons_per_step": B * a.window * a.features * a.features,
        "frame_pairs_per_step": B * a.window,
        "parallelism": "pose ranges sharded over %d rank(s), no data-path collective" % a.gpus,
    }


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.015):
        self.index = index
        self.period = period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        names = {}
        if nv is not None:
            for k in dir(nv):
                if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason"):
                    v = getattr(nv, k)
                    if isinstance(v, int) and v:
                        names.setdefault(v, k.replace("nvmlClocksEventReason", "")
                                         .replace("nvmlClocksThrottleReason", ""))
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, name in names.items():
                        if mask & bit and bit != 1:      # bit 1 = GpuIdle
                            self.reasons.add(name)
                else:
                    out = subprocess.run(
                        ["nvidia-smi", "-i", str(self.index),
                         "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                         "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"],
                        capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append(int(out[0]))
                    self.max_mhz = int(out[1])
                    for nm, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                      "sw_power_cap"], out[2:]):
                        if v.strip().lower().startswith("active"):
                            self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_mhz_min": (min(self.samples) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------- CPU legs
class CpuPair:
    """The reference's per-frame-pair CPU work: BFMatcher.knnMatch(k=2) + ratio
    (Frontend::GetMatches, slam_frontend.cc:521-538) + std::sort + best_percent cut
    (GetFeatureMatches, :289-291).  The C++ reference does the ratio/sort/cut glue
    in microseconds; Python attribute access on 10^4 cv2.DMatch objects would add
    milliseconds that the reference never pays, so the timed call is the bare
    knnMatch plus libstdc++'s std::sort on the (precomputed) survivor distances —
    i.e. the reference's real cost, nothing more."""

    def __init__(self, past, cur, cv2_ref, restate, use_cv2):
        from oracle import native
        self.past, self.cur = np.ascontiguousarray(past), np.ascontiguousarray(cur)
        self.cv2_ref, self.restate, self.use_cv2, self.native = cv2_ref, restate, use_cv2, native
        self.survivors = native.get_matches(past, cur, RATIO)      # untimed, for the sort leg

    def run(self):
        if self.use_cv2:
            raw = self.cv2_ref.knn_match_raw(self.past, self.cur)
        else:
            raw = self.native.knn2_hamming(self.past, self.cur)
        order = self.restate.sort_order_stdsort(self.survivors)
        keep = self.restate.num_good_matches(len(self.survivors), np.float32(BEST_PERCENT))
        return raw, order[:keep]


def cpu_setup():
    from oracle import cv2_ref, restate
    use_cv2 = cv2_ref.available()
    if use_cv2:
        cores = cv2_ref.set_threads(os.cpu_count() or 1)
        label = "cv2 %s BFMatcher(NORM_HAMMING).knnMatch(k=2) [the OpenCV function the reference calls] " \
                "+ restated ratio/sort/cut glue" % cv2_ref.version()
    else:
        from oracle import native
        cores = native.num_threads()
        label = "oracle/oracle_knn.c (OpenMP) + restated ratio/sort/cut glue"
    return cv2_ref, restate, use_cv2, cores, label


def cpu_pairs(a, n_frames):
    cv2_ref, restate, use_cv2, cores, label = cpu_setup()
    n, stride = a.features, (a.stride or max(1, a.features // 10))
    frames = [synth.synth_pose(n, p, stride, SEED, a.desc_bytes) for p in range(n_frames)]
    work = [CpuPair(frames[j], frames[j + 1], cv2_ref, restate, use_cv2) for j in range(n_frames - 1)]
    return work, cores, label


def cpu_baseline(a, budget_s):
    """Bounded sample of the same workload on the host cores (rank 0, N=1)."""
    work, cores, label = cpu_pairs(a, 3)
    n = a.features
    work[0].run()                                                         # warm-up
    t0 = time.perf_counter()
    pairs = 0
    while True:
        work[pairs % 2].run()
        pairs += 1
        el = time.perf_counter() - t0
        if el >= budget_s or pairs >= 5000:
            break
    return {"value": pairs * n * n / el, "unit": UNIT, "cores": int(cores), "kind": "port",
            "frame_pairs_per_s": pairs / el,
            "sample": "%d frame pairs of %dx%d (%d-byte rows) in %.1f s; %s; the reference C++ cannot be built here "
                      "(no OpenCV C++/Eigen/ROS), so this is the restated path on the OpenCV wheel"
                      % (pairs, n, n, a.desc_bytes, el, label)}


def run_reference(a):
    """--impl reference: the reference's CPU path with all host threads, same metric/config.
    A step of the GPU arm is a batch of `poses_per_step` poses; the CPU needs ~0.1 s per pose, so
    every step here times a bounded SAMPLE of that batch - one pose, i.e. `window` frame pairs -
    and `ms_per_step` is that sample's time scaled to the whole batch (value is unaffected:
    comparisons of the sample / time of the sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n, W, B = a.features, a.window, a.poses_per_step
    work, cores, label = cpu_pairs(a, W + 1)
    for _ in range(max(1, min(a.warmup, 3))):
        work[0].run()
    t0 = time.perf_counter()
    work[0].run()
    per_pair = time.perf_counter() - t0
    # size the run to a few minutes at most
    max_steps = max(1, int(150.0 / max(per_pair * W, 1e-6)))
    timed = min(a.steps, max_steps)
    t0 = time.perf_counter()
    for s in range(timed):
        for j in range(W):
            work[j].run()
    el = time.perf_counter() - t0
    value = timed * W * n * n / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * el / timed * B,
        "sample_ms_per_step": 1e3 * el / timed,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": workload_config(a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(cores), "kind": "port",
                         "sample": "each step = ONE pose of the %d-pose batch (%d frame pairs of %dx%d, 1/%d of a "
                                   "step); ms_per_step is the sample scaled to the batch; %d of %d steps timed; %s"
                                   % (B, W, n, n, B, timed, a.steps, label)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frame_pairs_per_s": timed * W / el,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
_REAL_STDOUT = None   # set when file descriptor 1 has been redirected (multi-rank runs)


def run_b200(a):
    import torch
    import torch.distributed as dist

    import vision_slam_frontend_b200 as vsf
    from vision_slam_frontend_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout = one JSON line: NCCL writes its version banner (and any NCCL_DEBUG output)
        # to file descriptor 1 from native code, so the descriptor itself is pointed at stderr and
        # the JSON line goes to a saved copy of the real stdout
        global _REAL_STDOUT
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        if "VSF_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["VSF_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    a.gpus = world

    n, W, B = a.features, a.window, a.poses_per_step
    stride = a.stride or max(1, n // 10)
    K, WU = a.steps, max(a.warmup, 3)
    cmp_per_pose = W * n * n
    cmp_per_step = B * cmp_per_pose

    ctx = vsf.Context(device=local, max_features=n, desc_bytes=a.desc_bytes, window=W)
    row_bytes = ctx.row_bytes
    # a dedicated (non-default) stream: kernels, events and the clock are all on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_tuning(a.popc_mode, a.split, a.qpt, a.variant)
    ctx.set_engine(a.engine, int(os.environ.get("VSF_ENGINE_FLAGS", "0")))   # flags: A/B timing knobs of include/vsf.h
    # host threads for the reference's std::sort (e2e leg): the ranks of one box share its cores
    host_threads = max(2, min(16, (os.cpu_count() or 16) // world))
    if os.environ.get("VSF_HOST_THREADS"):
        host_threads = int(os.environ["VSF_HOST_THREADS"])
    ctx.set_host_threads(host_threads)

    # ---- device-resident sequence: a window of this rank's pose range, larger than L2; the
    # steps walk it cyclically (every frame is re-read from HBM: 189 MB never fit the 126 MB L2)
    frame_bytes = n * row_bytes
    n_poses = max(2 * W + 2, int(1.5 * L2_BYTES / frame_bytes) + 1)
    first_pose, _ = sharding.pose_range(rank, world, (K + WU) * B * world)
    seq = torch.empty((n_poses, n, row_bytes), dtype=torch.uint8, device="cuda")
    ctx.synth_sequence_device(seq.data_ptr(), n, first_pose, n_poses, stride, SEED)
    flush = torch.empty(int(1.5 * L2_BYTES), dtype=torch.uint8, device="cuda")
    base = seq.data_ptr()

    def launch_step(s):
        # poses [s*B, (s+1)*B) of the walk, each against the W poses before it
        ctx.window_match_block_device(base, n, n_poses, s * B, B, RATIO)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush.fill_(1)                      # evict the freshly generated sequence from L2 ...
    torch.cuda.synchronize()
    for s in range(WU):                 # ... then warm up (clocks, instruction caches, allocator)
        launch_step(s)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count()
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for s in range(K):
            launch_step(WU + s)
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    two_kernels = launches < 3 * K * B
    counts = ctx.fetch_window(W, with_matches=False)
    engine = ctx.last_engine
    # ---- per-kernel durations (CUDA events on the ctx stream around every kernel of a launch;
    # a separate pass because the events serialise kernels that otherwise overlap via PDL)
    # (the kernels are timed ALONE here, which is what the burst peak they are held against was
    # measured as; half a second of idle first, so that this pass does not inherit the clock the
    # power cap left the GPU at after the back-to-back region - the pass reports its own clock)
    # A launch of the timed region is a group of GP poses as one batch: that is the
    # launch timed here, GP poses per call.
    time.sleep(0.5)
    ctx.set_profile(True)
    kt = np.zeros(4)
    GP = max(1, min(4, 40 // W)) if engine >= 2 else 1
    KP = 200 // GP
    with ClockSampler(local, period_s=0.002) as kclocks:
        for t in range(KP):
            ctx.window_match_block_device(base, n, n_poses, t * GP, GP, RATIO)
            kt += np.array(ctx.last_kernel_times())
    ctx.set_profile(False)
    kt /= KP
    rank_ms = [ms]
    rank_mhz = [clocks.summary()["sm_mhz"] or 0.0]
    if world > 1:
        g = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(g, torch.tensor([ms, rank_mhz[0]], dtype=torch.float64, device="cuda"))
        rank_ms = [float(t[0].item()) for t in g]
        rank_mhz = [float(t[1].item()) for t in g]
    ms_max = max(rank_ms)
    value = world * K * cmp_per_step / (ms_max * 1e-3)

    # ---- e2e: the host-buffer C-ABI call, H2D + kernel + D2H (+ sort/cut) every step ---------
    e2e = None
    if not a.no_e2e:
        e2e = measure_e2e(a, ctx, seq, n, W, K, WU, B, world, dist, torch, host_threads)

    # ---- gather one pose's match lists from every rank (outside the timed region): libvsf_nccl's
    # vsf_gather_matches, two ncclAllGathers straight from the device regions on the ctx stream
    gathered = None
    if world > 1:
        from vision_slam_frontend_b200 import nccl
        comm = nccl.Comm.from_torch_distributed(local)
        ctx.window_match_block_device(base, n, n_poses, 0, 1, RATIO)
        S = nccl.match_list_stride(ctx)
        d_counts = torch.zeros((world, W), dtype=torch.int32, device="cuda")
        d_lists = torch.zeros((world, W, S, 16), dtype=torch.uint8, device="cuda")
        comm.gather_matches(ctx, W, d_counts.data_ptr(), d_lists.data_ptr())     # warm-up (NCCL channel setup)
        ctx.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        comm.gather_matches(ctx, W, d_counts.data_ptr(), d_lists.data_ptr())
        g1.record(stream)
        ctx.synchronize()
        gathered = {"matches": int(d_counts.sum().item()), "ranks": world, "ms": g0.elapsed_time(g1),
                    "bytes_per_rank": W * S * 16 + 4 * W,
                    "api": "vsf_gather_matches (libvsf_nccl.so): ncclAllGather of counts + list regions, device to device"}
        comm.close()

    if rank == 0:
        # integer-pipe denominators, measured here
        popc_rate = ctx.probe_pipe(0, 4096)
        lop3_rate = ctx.probe_pipe(1, 4096)
        mixed_rate = ctx.probe_pipe(2, 4096)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
        per_pose_s = ms * 1e-3 / (K * B)
        alg_bytes = row_bytes * (W * n + n) + 16 * (W * n)   # descriptors read + {idx0,idx1,d0,d1} written
        achieved_gbs = alg_bytes / per_pose_s / 1e9
        cmp_rate_1gpu = cmp_per_pose / per_pose_s
        popc_peak_cmp = popc_rate / (row_bytes / 4.0)      # one POPC per 32-bit word of a comparison
        main_s = float(kt[1]) * 1e-3                       # dominant kernel, average launch duration
        tensor = engine >= 2
        wide = row_bytes == 64
        # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        traffic, traffic_src = None, None
        if n == 5000 and W == 10:
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                key = ("tensor_int8_wide" if wide else "tensor_int8") if engine == 2 else ("popc" if engine == 1 else "none")
                te = tj[key]
                traffic, traffic_src = te["traffic_bytes"], te["source"]
            except Exception:
                pass
        if tensor:
            # one comparison = one dot product of +-1 bytes over every descriptor bit: 2 ops per bit
            ops_per_cmp = 2.0 * 8 * row_bytes
            ops = ops_per_cmp * cmp_per_pose * GP              # per launch: a group of GP poses
            bf16 = float(peaks.get("bf16_tflops", 1590.0))
            kname = ("vsf::w64::pair::knn2_tc64_pair_kernel (cta_group::2)" if wide else
                     ("vsf::knn2_tc_kernel<int8>" if engine == 2 else "vsf::knn2_tc_kernel<e4m3>"))
            roofline = {
                "bound": "tensor", "achieved": ops / main_s / 1e12, "peak": 2.0 * bf16, "unit": "TFLOP/s",
                "frac": ops / main_s / 1e12 / (2.0 * bf16), "traffic": traffic, "traffic_unit": "bytes/launch",
                "traffic_source": traffic_src,
                "kernel": kname,
                "kernel_ms": float(kt[1]),
                "poses_per_launch": GP,
                "kernel_timing": "CUDA events around the kernel, %d launches of %d poses each (the launch shape of the "
                                 "timed region) one at a time after 0.5 s of idle (SM %s MHz in that pass; the timed "
                                 "region ran at %s MHz under the power cap)"
                                 % (KP, GP, kclocks.summary()["sm_mhz"], clocks.summary()["sm_mhz"]),
                "peak_source": "2 x bf16_tflops (burst) of %s: 8-bit operands run the tensor pipe at twice the "
                               "bf16 rate; ops are int8 multiply-accumulates counted as 2 (TOP/s)" % peak_src,
                "algorithmic_ops_per_launch": ops,
                "peak_sustained": 2.0 * float(peaks.get("bf16_tflops_sustained", 0.0)) or None,
                "frac_of_sustained": (ops / main_s / 1e12 / (2.0 * float(peaks["bf16_tflops_sustained"])))
                                     if peaks.get("bf16_tflops_sustained") else None,
                "nominal_peak": 4500.0,
                "frac_of_nominal": ops / main_s / 1e12 / 4500.0,
                "whole_step_frac": ops / GP / per_pose_s / 1e12 / (2.0 * bf16),
                "other_kernels_ms": {"expand_train": float(kt[0]), "refine": float(kt[2]), "compact": float(kt[3])},
                "note": "%d ops per %d-bit comparison x comparisons per launch (a group of poses as one batch) / "
                        "CUDA-event duration of the tensor-core kernel; the finish kernel (exact POPC re-scan of 16 "
                        "train rows per candidate query, ratio test, ordered compaction) is the other kernel of a "
                        "launch sequence, listed in other_kernels_ms; whole_step_frac divides a pose's ops by the "
                        "whole pose time of the timed region.  frac may exceed 1: its denominator, twice the "
                        "cuBLAS bf16 rate the driver measured, is a proxy (cuBLAS reaches ~73 %% of the nominal 2.25 "
                        "PFLOP/s); frac_of_nominal holds the kernel to the nominal dense int8 peak.  frac is against "
                        "the burst figure (kernel timed alone); the kernel runs back to back for the whole timed "
                        "region, for which the driver's sustained figure (frac_of_sustained) is the like-for-like "
                        "denominator: under tensor load the SM clock the kernel itself sees is ~1.72 GHz, not "
                        "the 1.965 GHz nvidia-smi reports (profiles/r01b_tc_timeline_c4.json)"
                        % (int(ops_per_cmp), 8 * row_bytes),
            }
        else:
            roofline = {
                "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak, "traffic": traffic, "traffic_unit": "bytes/launch",
                "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "POPC engine: integer-pipe bound by design (arithmetic intensity N/32 comparisons "
                        "per byte), HBM is idle at roofline; see roofline_int",
            }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": WU, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "i8" if engine == 2 else ("f8e4m3" if engine == 3 else "u32"),
            "data": "synthetic",
            "config": workload_config(a),
            "l2": "sequence buffer %.0f MB > %d MB L2, walked cyclically, each frame first read from HBM; L2 flushed, "
                  "then %d warm-up steps, then the timed region" % (n_poses * frame_bytes / 2 ** 20, L2_BYTES // 2 ** 20, WU),
            "timed_region_s": ms_max * 1e-3,
            "us_per_pose": 1e3 * ms_max / (K * B),
            "per_rank_ms": rank_ms,
            "per_rank_sm_mhz": rank_mhz,
            "matched_frame_pairs_per_s": world * K * B * W / (ms_max * 1e-3),
            "engine": {1: "popc", 2: "tensor_int8", 3: "tensor_e4m3"}.get(engine, str(engine)),
            "roofline": roofline,
            "roofline_hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved_gbs / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes,
                             "note": "whole pose; the path is compute-bound, HBM is idle by design"},
            "roofline_int": {
                "bound": "int_pipe_popc", "achieved": cmp_rate_1gpu, "unit": "cmp/s per GPU",
                "peak": popc_peak_cmp, "frac": cmp_rate_1gpu / popc_peak_cmp,
                "peak_source": "POPC lane-ops/s measured by vsf_probe_pipe on this GPU / %d POPC per comparison"
                               % (row_bytes // 4),
                "popc_ops_per_s": popc_rate, "lop3_ops_per_s": lop3_rate,
                "popc_lop3_mixed_ops_per_s": mixed_rate,
                "nominal_peak": 16 * ctx.sm_count * 1.965e9 / (row_bytes / 4.0),
                "note": "north-star denominator: the integer-pipe roofline of the naive one-POPC-per-word "
                        "comparison; the tensor-core engine is not bound by it (frac > 1)",
            },
            "gpu_launches": int(launches),
            "kernel": (("knn2_tc (distance) + knn2_tc_finish (refine + ordered compaction) kernels; the poses are "
                        "launched in groups of four, a group as ONE batch of its 40 frame pairs (distance + finish "
                        "launch, the next group's frames expanded beside them on a side stream), programmatic "
                        "dependent launch")
                       if two_kernels else
                       ("expand_train + knn2_tc + refine + compact kernels "
                        "(4 launches per pose, programmatic dependent launch)")) if tensor
                      else "vsf::knn2_kernel<WORDS,R,MODE> (one launch per pose)",
            "kernel_ms": ({"main": float(kt[1]), "finish": float(kt[2]), "poses_per_launch": GP,
                           "note": "CUDA events around the two kernels of a launch sequence (a group of poses as one "
                                   "batch), one group at a time (the events serialise what overlaps in the timed "
                                   "region); the expansion of the next group's frames runs on a side stream"}
                          if two_kernels else
                          {"expand_train": float(kt[0]), "main": float(kt[1]), "refine": float(kt[2]),
                           "compact": float(kt[3])}),
            "last_step_survivors": [int(c) for c in counts],
            "clocks": clocks.summary(),
        }
        if e2e:
            line["e2e"] = e2e
        if gathered is not None:
            line["nccl_gather"] = gathered
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a, a.cpu_seconds)
        if world == 1 and not a.no_extra:
            ctx.close()
            ctx = None
            del seq, flush
            torch.cuda.empty_cache()
            line["extra_workloads"] = extra_workloads(a, torch, stream)
        print(json.dumps(line), file=_REAL_STDOUT or sys.stdout, flush=True)
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(a, ctx, seq, n, W, K, WU, B, world, dist, torch, host_threads):
    """Same metric through the reference-facing C-ABI calls with HOST buffers.  Every pose moves
    the new frame's descriptors host->device from pinned memory, runs the kernels, brings the
    ratio survivors back, sorts + cuts them (best_percent) into FeatureMatch lists in host
    memory and pushes the frame into the window.  Variants:
      pipelined  vsf_window_run_sequence = the loop vsf_window_submit(t) ... vsf_window_collect(t - lag)
                 written in C++ like a caller of the C ABI would: the host sorts frame t-lag while the
                 device matches frame t (headline; host std::sort = reference order).  All K steps
                 (K x B frames) in one timed region.  sort_mode 2 = the same order, replayed on
                 the device (no host cores); sort_mode 0 = stable device sort
      sync       vsf_window_feature_matches + vsf_window_commit, one blocking call per frame
                 (bounded: 300 frames)"""
    from vision_slam_frontend_b200 import PIPELINE_DEPTH, capi
    L = ctx._L
    rb = ctx.row_bytes
    n_host = min(seq.shape[0], 1200)
    host = torch.empty((n_host, n, rb), dtype=torch.uint8).pin_memory()
    host.copy_(seq[:n_host])
    torch.cuda.synchronize()
    hp = host.numpy()
    lag = max(1, min(a.e2e_lag, PIPELINE_DEPTH - 1))
    ring = 4
    out = np.zeros((ring, W, n), capi.FEATURE_MATCH_DTYPE)
    cnts = np.zeros((ring, W), np.int32)
    res = {}

    def maxrank(x):
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for sort_mode, key in ((1, "pipelined_exact_host_sort"), (2, "pipelined_exact_device_sort"),
                           (0, "pipelined_stable_device_sort")):
        ctx.window_clear()
        for p in range(W):
            ctx.window_push(p, hp[p][:, :a.desc_bytes] if a.desc_bytes != rb else hp[p])
        ctx.window_run_sequence(hp, W, WU * B // 4 + 8, RATIO, BEST_PERCENT, sort_mode, lag, out, cnts)   # warm-up
        first = W + WU * B // 4 + 8
        sync_all()
        t0 = time.perf_counter()
        h2d, d2h = ctx.window_run_sequence(hp, first, K * B, RATIO, BEST_PERCENT, sort_mode, lag, out, cnts)
        torch.cuda.synchronize()
        el = maxrank(time.perf_counter() - t0)
        res[key] = {"value": world * K * B * W * n * n / el, "ms_per_step": 1e3 * el / K,
                    "us_per_pose": 1e6 * el / (K * B),
                    "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K, "frames_in_flight": lag + 1}

    fids = np.zeros(W, np.uint64)
    counts = np.zeros(W, np.int32)
    out1 = np.zeros((W, n), capi.FEATURE_MATCH_DTYPE)
    nf = C.c_int(0)
    Ks = 300
    for sort_mode, key in ((1, "sync_exact_stdsort"), (0, "sync_device_sort")):
        ctx.window_clear()
        for p in range(W):
            ctx.window_push(p, hp[p][:, :a.desc_bytes] if a.desc_bytes != rb else hp[p])

        def step(t):
            D = hp[(W + t) % n_host]
            rc = L.vsf_window_feature_matches(ctx._h, D.ctypes.data, n, rb, RATIO, BEST_PERCENT, sort_mode,
                                              fids.ctypes.data, counts.ctypes.data, out1.ctypes.data, n, C.byref(nf))
            if rc:
                raise RuntimeError(L.vsf_last_error(ctx._h).decode())
            L.vsf_window_commit(ctx._h, W + t, n)

        for t in range(10):
            step(t)
        sync_all()
        t0 = time.perf_counter()
        for t in range(Ks):
            step(10 + t)
        torch.cuda.synchronize()
        el = maxrank(time.perf_counter() - t0)
        res[key] = {"value": world * Ks * W * n * n / el, "us_per_pose": 1e6 * el / Ks, "frames": Ks}

    # headline e2e: pipelined, the reference's bit-identical order, computed where
    # VSF_SORT_EXACT_AUTO puts it: host std::sort with >= 12 host threads per rank, else the device
    # replay of std::sort (e.g. 4 or 8 ranks sharing a 32-core node)
    auto_mode = 1 if host_threads >= 12 else 2
    head = res["pipelined_exact_host_sort" if auto_mode == 1 else "pipelined_exact_device_sort"]
    return {"value": head["value"], "unit": UNIT, "h2d_bytes_per_step": head["h2d_bytes_per_step"],
            "d2h_bytes_per_step": head["d2h_bytes_per_step"], "steps": K,
            "ms_per_step": head["ms_per_step"], "us_per_pose": head["us_per_pose"],
            "api": "vsf_window_run_sequence = vsf_window_submit / vsf_window_collect per frame (%d frames in flight; "
                   "sort_mode=VSF_SORT_EXACT_AUTO -> %s: the reference's std::sort order, bit-identical), pinned host "
                   "buffers; every pose's H2D, kernels, D2H, sort + cut inside the timed region; all %d x %d frames of "
                   "the K steps in one region"
                   % (lag + 1, "host std::sort (mode 1)" if auto_mode == 1 else "device replay of std::sort (mode 2)", K, B),
            "sort_mode": auto_mode,
            "matched_frame_pairs_per_s": head["value"] / (n * n),
            "host_threads": host_threads,
            "variants": res}


# ------------------------------------------------------------------- the other BASELINE configs
def extra_workloads(a, torch, stream):
    """Short, bounded legs for the BASELINE configs that are not the headline workload, each
    with the CPU figure (OpenCV on all host threads) from the same run:
      C2  2000 x 2000 single pair: host-API latency, back-to-back launches, CUDA-graph replay
      C3  stereo L/R match + filter + R'->L' match + triangulation, 2000 features, KITTI rig
      C5  20000 features x 32-frame window, a sample of 32 poses
      C4w the headline shape at the other descriptor width (61-byte AKAZE rows when the main
          run is 32-byte, and vice versa), device-resident and end to end"""
    import vision_slam_frontend_b200 as vsf
    from oracle import cv2_ref, restate
    out = {}
    have_cv2 = cv2_ref.available()
    ncpu = os.cpu_count() or 1
    if have_cv2:
        cv2_ref.set_threads(ncpu)

    def best_of(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts), float(np.median(ts))

    def dev_time(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / reps

    # ---- C2
    try:
        Q, T = synth.descriptor_pair(2000, 2000, seed=0)
        with vsf.Context(max_features=2048, window=2) as ctx:
            ctx.set_stream(stream.cuda_stream)
            lat = best_of(lambda: ctx.get_matches(Q, T, RATIO), 100, 5)
            dq, dt = torch.from_numpy(Q).cuda(), torch.from_numpy(T).cuda()
            qp = (C.c_void_p * 1)(dq.data_ptr())
            nn = (C.c_int * 1)(2000)

            def launch():
                assert ctx._L.vsf_window_match_device(ctx._h, qp, nn, 1, C.c_void_p(dt.data_ptr()), 2000, RATIO) == 0
            for _ in range(20):
                launch()
            b2b = dev_time(launch, 500)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                launch()
            for _ in range(5):
                g.replay()
            gr = dev_time(g.replay, 500)
            eng = ctx.last_engine
        cpu = best_of(lambda: cv2_ref.knn_match_raw(Q, T), 5, 1)[0] if have_cv2 else None
        out["C2"] = {"shape": "2000 x 2000 x 256-bit single pair (4e6 comparisons)", "engine": eng,
                     "host_api_latency_us": 1e6 * lat[0], "host_api_latency_median_us": 1e6 * lat[1],
                     "back_to_back_us": 1e6 * b2b, "graph_replay_us": 1e6 * gr,
                     "cmp_per_s_host_api": 4e6 / lat[0], "cmp_per_s_graph": 4e6 / gr,
                     "cpu_us": None if cpu is None else 1e6 * cpu, "cpu_cmp_per_s": None if cpu is None else 4e6 / cpu,
                     "cpu_threads": ncpu}
    except Exception as e:      # noqa: BLE001
        out["C2"] = {"error": repr(e)}

    # ---- C3
    try:
        P1, P2 = synth.kitti_projections()
        F = synth.kitti_fundamental()
        kl, dl, kr, dr, X, perm = synth.stereo_frame(2000, seed=1)
        with vsf.Context(max_features=2048, window=2) as ctx:
            def frame():
                ctx.window_clear()
                return ctx.observe_features(0, kl, dl, kr, dr, F, P1, P2, RATIO)
            tf = best_of(frame, 50, 5)
            got = frame()
            ok = perm >= 0
            x1 = np.stack([kl["x"][ok], kl["y"][ok]], 1)
            x2 = np.stack([kr["x"][perm[ok]], kr["y"][perm[ok]]], 1)
            tt = best_of(lambda: ctx.triangulate(P1, P2, x1, x2), 30, 3)
            pipelined = None
            if hasattr(ctx, "observe_submit"):
                pipelined = c3_pipelined(ctx, (kl, dl, kr, dr), F, P1, P2, torch)
        cpu_f = cpu_t = None
        if have_cv2:
            fo = restate.FrontendOracle(P1, P2, F, frame_life=1)
            r = fo.observe_features(kl, dl, kr, dr)

            def cpu_c3():
                cv2_ref.knn_match_raw(dl, dr)                                   # stereo L->R
                cv2_ref.knn_match_raw(r.right.descriptors, r.left.descriptors)  # R'->L'
                cv2_ref.triangulate_points(P1, P2, x1[:len(r.points)], x2[:len(r.points)])
            cpu_f = best_of(cpu_c3, 5, 1)[0]
            cpu_t = best_of(lambda: cv2_ref.triangulate_points(P1, P2, x1, x2), 5, 1)[0]
        out["C3"] = {"shape": "stereo L->R match + epipolar filter + R'->L' match + triangulation, 2000 features, "
                              "KITTI 1241x376 rig, one blocking vsf_observe_features call per frame",
                     "observe_features_us": 1e6 * tf[0], "observe_features_median_us": 1e6 * tf[1],
                     "kept_pairs": int(len(got["kept_left"])), "triangulated": int(len(got["tri_matches"])),
                     "triangulate_only_us": 1e6 * tt[0], "points": int(len(x1)),
                     "triangulated_points_per_s_host_api": len(x1) / tt[0],
                     "pipelined": pipelined,
                     "cpu_frame_us": None if cpu_f is None else 1e6 * cpu_f,
                     "cpu_triangulate_only_us": None if cpu_t is None else 1e6 * cpu_t,
                     "cpu_triangulated_points_per_s": None if cpu_t is None else len(x1) / cpu_t, "cpu_threads": ncpu}
    except Exception as e:      # noqa: BLE001
        out["C3"] = {"error": repr(e)}

    # ---- C2 / C3 as a C++ caller sees them (vsf_latency_probe: no Python between the calls)
    try:
        exe = os.path.join(ROOT, "vision_slam_frontend_b200", "vsf_latency_probe")
        pr = subprocess.run([exe, "0", "2000", "32", "1"], capture_output=True, text=True, timeout=120)
        out["cxx_latency"] = json.loads(pr.stdout.strip().splitlines()[-1]) if pr.returncode == 0 else \
            {"error": "rc %d: %s" % (pr.returncode, pr.stderr[-300:])}
    except Exception as e:      # noqa: BLE001
        out["cxx_latency"] = {"error": repr(e)}

    # ---- C5 sample
    try:
        n5, W5, poses = 20000, 32, 32
        with vsf.Context(max_features=n5, window=W5) as ctx:
            ctx.set_stream(stream.cuda_stream)
            npz = 2 * W5 + 8
            seq = torch.empty((npz, n5, 32), dtype=torch.uint8, device="cuda")
            ctx.synth_sequence_device(seq.data_ptr(), n5, 0, npz, n5 // W5, 7)
            ctx.window_match_block_device(seq.data_ptr(), n5, npz, 0, 2, RATIO)
            t5 = dev_time(lambda: ctx.window_match_block_device(seq.data_ptr(), n5, npz, 2, poses, RATIO), 1) / poses
            surv = int(ctx.fetch_window(W5, with_matches=False).sum())
            hostp = seq[:2].cpu().numpy()
            del seq
        cpu5 = best_of(lambda: cv2_ref.knn_match_raw(hostp[0], hostp[1]), 2, 0)[0] if have_cv2 else None
        out["C5"] = {"shape": "20000 features x 32-frame window; %d consecutive poses timed, device-resident" % poses,
                     "comparisons_per_pose": W5 * n5 * n5, "ms_per_pose": 1e3 * t5, "cmp_per_s": W5 * n5 * n5 / t5,
                     "survivors_last_pose": surv,
                     "cpu_one_pair_s": cpu5, "cpu_cmp_per_s": None if cpu5 is None else n5 * n5 / cpu5,
                     "cpu_s_per_pose_extrapolated": None if cpu5 is None else cpu5 * W5, "cpu_threads": ncpu,
                     "hours_100k_poses_on_8_gpus": 100000 / 8 * t5 / 3600}
    except Exception as e:      # noqa: BLE001
        out["C5"] = {"error": repr(e)}

    # ---- C4 at the other descriptor width
    try:
        width = 61 if a.desc_bytes <= 32 else 32
        n, W = a.features, a.window
        stride = a.stride or max(1, n // 10)
        with vsf.Context(max_features=n, desc_bytes=width, window=W) as ctx:
            ctx.set_stream(stream.cuda_stream)
            rb = ctx.row_bytes
            npz = 300
            seq = torch.empty((npz, n, rb), dtype=torch.uint8, device="cuda")
            ctx.synth_sequence_device(seq.data_ptr(), n, 0, npz, stride, SEED)
            ctx.window_match_block_device(seq.data_ptr(), n, npz, 0, 64, RATIO)
            tw = dev_time(lambda: ctx.window_match_block_device(seq.data_ptr(), n, npz, 64, 512, RATIO), 1) / 512
            ctx.set_profile(True)
            kt = np.zeros(4)
            for t in range(50):
                ctx.window_match_block_device(seq.data_ptr(), n, npz, t, 1, RATIO)
                kt += np.array(ctx.last_kernel_times())
            ctx.set_profile(False)
            kt /= 50
            host = torch.empty((npz, n, rb), dtype=torch.uint8).pin_memory()
            host.copy_(seq)
            torch.cuda.synchronize()
            hp = host.numpy()
            from vision_slam_frontend_b200 import capi
            lag = max(1, min(a.e2e_lag, capi.PIPELINE_DEPTH - 1))
            ob = np.zeros((4, W, n), capi.FEATURE_MATCH_DTYPE)
            oc = np.zeros((4, W), np.int32)
            for p in range(W):
                ctx.window_push(p, hp[p][:, :width])
            ctx.window_run_sequence(hp, W, 64, RATIO, BEST_PERCENT, 1, lag, ob, oc)
            t0 = time.perf_counter()
            ctx.window_run_sequence(hp, W + 64, 512, RATIO, BEST_PERCENT, 1, lag, ob, oc)
            te = (time.perf_counter() - t0) / 512
            eng = ctx.last_engine
            h2 = seq[:2].cpu().numpy()[:, :, :width].copy()
            del seq
        cpuw = best_of(lambda: cv2_ref.knn_match_raw(h2[0], h2[1]), 3, 1)[0] if have_cv2 else None
        ops = 2.0 * 8 * rb * W * n * n
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16 = float(peaks.get("bf16_tflops", 1590.0))
        out["C4_other_width"] = {
            "shape": "%d features x %d prior frames, %d-byte descriptors (%d-byte device rows)" % (n, W, width, rb),
            "engine": eng, "us_per_pose": 1e6 * tw, "cmp_per_s": W * n * n / tw,
            "e2e_us_per_pose": 1e6 * te, "e2e_cmp_per_s": W * n * n / te,
            "kernel_ms": {"expand_train": float(kt[0]), "main": float(kt[1]), "finish": float(kt[2])},
            "roofline_frac_main_kernel": ops / (float(kt[1]) * 1e-3) / 1e12 / (2.0 * bf16) if kt[1] > 0 else None,
            "cpu_one_pair_s": cpuw, "cpu_cmp_per_s": None if cpuw is None else n * n / cpuw, "cpu_threads": ncpu}
    except Exception as e:      # noqa: BLE001
        out["C4_other_width"] = {"error": repr(e)}
    return out


def c3_pipelined(ctx, frame, F, P1, P2, torch):
    """C3 through the pipelined full-frame calls (vsf_observe_submit / vsf_observe_collect)."""
    kl, dl, kr, dr = frame
    K = synth.KITTI_K.astype(np.float32)
    dist = np.zeros(5, np.float32)
    reps, lag = 200, 3
    ctx.window_clear()
    for t in range(8):
        ctx.observe_submit(t, kl, dl, kr, dr, F, P1, P2, K, dist, RATIO)
        if ctx.observe_in_flight() > lag:
            ctx.observe_collect()
    while ctx.observe_in_flight():
        ctx.observe_collect()
    t0 = time.perf_counter()
    for t in range(reps):
        ctx.observe_submit(8 + t, kl, dl, kr, dr, F, P1, P2, K, dist, RATIO)
        if ctx.observe_in_flight() > lag:
            ctx.observe_collect()
    while ctx.observe_in_flight():
        ctx.observe_collect()
    el = time.perf_counter() - t0
    # one frame at a time: submit + collect, host to host
    ts = []
    for t in range(50):
        t0 = time.perf_counter()
        ctx.observe_submit(1000 + t, kl, dl, kr, dr, F, P1, P2, K, dist, RATIO)
        ctx.observe_collect()
        ts.append(time.perf_counter() - t0)
    return {"us_per_frame_pipelined": 1e6 * el / reps, "frames_in_flight": lag + 1,
            "us_per_frame_submit_then_collect": 1e6 * min(ts), "window": ctx.window}


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
