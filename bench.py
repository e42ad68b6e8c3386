#!/usr/bin/env python
"""bench.py — Hamming comparisons/s of the frame-to-window matching hot path.

Workload (BASELINE.json configs[3], "C4"): every step is one pose of a synthetic
sequence — 5000 packed 256-bit descriptors matched (k=2 Hamming kNN + Lowe ratio
test + ordered compaction) against each of the 10 prior frames of the sliding
window: 10 x 5000 x 5000 = 2.5e8 comparisons and 10 matched frame pairs per step,
ONE kernel launch.  With N GPUs every rank walks its own pose range (weak
scaling, no data-path collective); match lists are gathered with NCCL after the
timed region only.

    python bench.py [--gpus N] [--steps K] [--warmup W]
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
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--features", type=int, default=5000)
    ap.add_argument("--window", type=int, default=10)
    ap.add_argument("--stride", type=int, default=0, help="landmark stride per pose (default N/10)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-lag", type=int, default=6,
                    help="frames submitted ahead of the one being collected in the pipelined e2e leg "
                         "(1 .. VSF_PIPELINE_DEPTH - 1)")
    ap.add_argument("--popc-mode", type=int, default=-1)
    ap.add_argument("--split", type=int, default=0)
    ap.add_argument("--qpt", type=int, default=0)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--engine", type=int, default=0,
                    help="0 auto (tensor cores on this workload), 1 POPC pipe, 2 tensor int8, 3 tensor e4m3")
    return ap.parse_args()


def workload_config(a, world):
    return {
        "workload": "C4 frame-to-window matching: %d features x %d prior frames per pose, "
                    "256-bit descriptors, k=2 Hamming kNN + ratio %.1f + ordered compaction"
                    % (a.features, a.window, 0.6),
        "features_per_frame": a.features, "window": a.window, "descriptor_bits": 256,
        "comparisons_per_step": a.window * a.features * a.features,
        "frame_pairs_per_step": a.window,
        "parallelism": "pose ranges sharded over %d rank(s), no data-path collective" % world,
    }


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
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
            self._stop.wait(0.02)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
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


def cpu_baseline(a, budget_s):
    """Bounded sample of the same workload on the host cores (rank 0, N=1)."""
    cv2_ref, restate, use_cv2, cores, label = cpu_setup()
    n, stride = a.features, (a.stride or max(1, a.features // 10))
    frames = [synth.synth_pose(n, p, stride, SEED) for p in range(3)]
    work = [CpuPair(frames[j], frames[j + 1], cv2_ref, restate, use_cv2) for j in range(2)]
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
            "sample": "%d frame pairs of %dx%d in %.1f s; %s; the reference C++ cannot be built here "
                      "(no OpenCV C++/Eigen/ROS), so this is the restated path on the OpenCV wheel"
                      % (pairs, n, n, el, label)}


def run_reference(a):
    """--impl reference: the reference's CPU path with all host threads, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cv2_ref, restate, use_cv2, cores, label = cpu_setup()
    n, stride = a.features, (a.stride or max(1, a.features // 10))
    pool = [synth.synth_pose(n, p, stride, SEED) for p in range(a.window + 1)]
    # a step = a bounded sample of one pose: ONE (past, current) frame pair of the
    # window (N x N comparisons).  Size the run to a few minutes at most.
    work = [CpuPair(pool[j], pool[j + 1], cv2_ref, restate, use_cv2) for j in range(a.window)]
    for _ in range(max(1, min(a.warmup, 3))):
        work[0].run()
    t0 = time.perf_counter()
    work[0].run()
    per_pair = time.perf_counter() - t0
    steps = a.steps
    max_steps = max(1, int(150.0 / max(per_pair, 1e-6)))
    timed = min(steps, max_steps)
    t0 = time.perf_counter()
    for s in range(timed):
        work[s % a.window].run()
    el = time.perf_counter() - t0
    value = timed * n * n / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * el / timed,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": workload_config(a, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(cores), "kind": "port",
                         "sample": "each step = one %dx%d frame pair of the window (1/%d of a pose); "
                                   "%d of %d steps timed; %s" % (n, n, a.window, timed, a.steps, label)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frame_pairs_per_s": timed / el,
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

    n, W = a.features, a.window
    stride = a.stride or max(1, n // 10)
    K, WU = a.steps, max(a.warmup, 3)
    cmp_per_step = W * n * n

    ctx = vsf.Context(device=local, max_features=n, desc_bytes=32, window=W)
    L = ctx._L
    # a dedicated (non-default) stream: kernels, events and the clock are all on it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_tuning(a.popc_mode, a.split, a.qpt, a.variant)
    ctx.set_engine(a.engine, 0)
    # host threads for the reference's std::sort (e2e leg): the ranks of one box share its cores
    host_threads = max(2, min(16, (os.cpu_count() or 16) // world))
    if os.environ.get("VSF_HOST_THREADS"):
        host_threads = int(os.environ["VSF_HOST_THREADS"])
    ctx.set_host_threads(host_threads)

    # ---- device-resident sequence: this rank's pose range, larger than L2 ----------------
    frame_bytes = n * 32
    n_poses = max(K + WU + W + 1, int(1.5 * L2_BYTES / frame_bytes) + 1)
    first_pose, _ = sharding.pose_range(rank, world, n_poses * world)
    seq = torch.empty((n_poses, n, 32), dtype=torch.uint8, device="cuda")
    ctx.synth_sequence_device(seq.data_ptr(), n, first_pose, n_poses, stride, SEED)
    flush = torch.empty(int(1.5 * L2_BYTES), dtype=torch.uint8, device="cuda")
    base = seq.data_ptr()

    def launch_step(t):
        # pose index (t + W) against the W poses before it — one kernel launch
        t = t % (n_poses - W)
        qp = (C.c_void_p * W)(*[base + (t + j) * frame_bytes for j in range(W)])
        nn = (C.c_int * W)(*([n] * W))
        rc = L.vsf_window_match_device(ctx._h, qp, nn, W, C.c_void_p(base + (t + W) * frame_bytes),
                                       n, RATIO)
        if rc:
            raise RuntimeError(L.vsf_last_error(ctx._h).decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(WU):
        launch_step(t)
    flush.fill_(1)                      # evict the freshly generated sequence from L2
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        e0.record(stream)
        for t in range(K):
            launch_step(WU + t)
        e1.record(stream)
        barrier()
    ms = e0.elapsed_time(e1)
    counts = ctx.fetch_window(W, with_matches=False)
    engine = ctx.last_engine
    # ---- per-kernel durations (CUDA events on the ctx stream around every kernel of a launch;
    # a separate pass because the events serialise kernels that otherwise overlap via PDL)
    ctx.set_profile(True)
    kt = np.zeros(4)
    KP = min(K, 200)
    for t in range(KP):
        launch_step(WU + t)
        kt += np.array(ctx.last_kernel_times())
    ctx.set_profile(False)
    kt /= KP
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * K * cmp_per_step / (ms_max * 1e-3)

    # ---- e2e: the host-buffer C-ABI call, H2D + kernel + D2H (+ sort/cut) every step ---------
    e2e = None
    if not a.no_e2e:
        e2e = measure_e2e(a, ctx, seq, n, W, K, WU, world, dist, torch)

    # ---- gather the last step's match lists over NCCL (outside the timed region) ------------
    gathered = None
    if world > 1:
        last = ctx.fetch_window(W)
        gathered = sharding.gather_match_lists(last, device=torch.device("cuda", local))
        gathered = sum(len(m) for r in gathered for m in r)

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
        per_launch_s = ms * 1e-3 / K
        alg_bytes = 32 * (W * n + n) + 16 * (W * n)        # descriptors read + {idx0,idx1,d0,d1} written
        achieved_gbs = alg_bytes / per_launch_s / 1e9
        cmp_rate_1gpu = cmp_per_step / per_launch_s
        popc_peak_cmp = popc_rate / 8.0                    # 8 POPC per 256-bit comparison
        main_s = float(kt[1]) * 1e-3                       # dominant kernel, average launch duration
        tensor = engine >= 2
        # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        traffic, traffic_src = None, None
        if n == 5000 and W == 10:
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
                te = tj["tensor_int8" if engine == 2 else ("popc" if engine == 1 else "none")]
                traffic, traffic_src = te["traffic_bytes"], te["source"]
            except Exception:
                pass
        if tensor:
            # one comparison = one 256-term dot product of +-1 bytes = 256 MACs = 512 ops
            ops = 512.0 * cmp_per_step
            bf16 = float(peaks.get("bf16_tflops", 1590.0))
            roofline = {
                "bound": "tensor", "achieved": ops / main_s / 1e12, "peak": 2.0 * bf16, "unit": "TFLOP/s",
                "frac": ops / main_s / 1e12 / (2.0 * bf16), "traffic": traffic, "traffic_unit": "bytes/launch",
                "traffic_source": traffic_src,
                "kernel": "vsf::knn2_tc_kernel<int8>" if engine == 2 else "vsf::knn2_tc_kernel<e4m3>",
                "kernel_ms": float(kt[1]),
                "peak_source": "2 x bf16_tflops (burst) of %s: 8-bit operands run the tensor pipe at twice the "
                               "bf16 rate; ops are int8 multiply-accumulates counted as 2 (TOP/s)" % peak_src,
                "algorithmic_ops_per_launch": ops,
                "peak_sustained": 2.0 * float(peaks.get("bf16_tflops_sustained", 0.0)) or None,
                "frac_of_sustained": (ops / main_s / 1e12 / (2.0 * float(peaks["bf16_tflops_sustained"])))
                                     if peaks.get("bf16_tflops_sustained") else None,
                "nominal_peak": 4500.0,
                "frac_of_nominal": ops / main_s / 1e12 / 4500.0,
                "other_kernels_ms": {"expand_train": float(kt[0]), "refine": float(kt[2]), "compact": float(kt[3])},
                "note": "512 ops per 256-bit comparison x comparisons per launch / CUDA-event duration of the "
                        "tensor-core kernel; the refine (exact POPC re-scan of <= 32 train rows per query) and "
                        "the compaction are separate, small kernels listed in other_kernels_ms.  frac is against "
                        "the burst figure (kernel timed alone); the kernel runs back to back for the whole timed "
                        "region, for which the driver's sustained figure (frac_of_sustained) is the like-for-like "
                        "denominator: under tensor load the SM clock the kernel itself sees is ~1.72 GHz, not "
                        "the 1.965 GHz nvidia-smi reports (profiles/r01_tc_timeline_c4.json)",
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
            "config": dict(workload_config(a, world),
                           l2="sequence buffer %.0f MB > %d MB L2, each frame first read from HBM; "
                              "L2 flushed before the timed region"
                              % (n_poses * frame_bytes / 2 ** 20, L2_BYTES // 2 ** 20)),
            "matched_frame_pairs_per_s": world * K * W / (ms_max * 1e-3),
            "engine": {1: "popc", 2: "tensor_int8", 3: "tensor_e4m3"}.get(engine, str(engine)),
            "roofline": roofline,
            "roofline_hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved_gbs / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes,
                             "note": "whole step; the path is compute-bound, HBM is idle by design"},
            "roofline_int": {
                "bound": "int_pipe_popc", "achieved": cmp_rate_1gpu, "unit": "cmp/s per GPU",
                "peak": popc_peak_cmp, "frac": cmp_rate_1gpu / popc_peak_cmp,
                "peak_source": "POPC lane-ops/s measured by vsf_probe_pipe on this GPU / 8 POPC per comparison",
                "popc_ops_per_s": popc_rate, "lop3_ops_per_s": lop3_rate,
                "popc_lop3_mixed_ops_per_s": mixed_rate,
                "nominal_peak": 2 * ctx.sm_count * 1.965e9,
                "note": "north-star denominator: the integer-pipe roofline of the naive 8-POPC comparison; the "
                        "tensor-core engine is not bound by it (frac > 1)",
            },
            "gpu_launches": K * (4 if tensor else 1),
            "kernel": ("expand_train_kernel + knn2_tc_kernel + knn2_tc_refine_kernel + knn2_compact_kernel "
                       "(4 launches per step, programmatic dependent launch)") if tensor
                      else "vsf::knn2_kernel<8,R,MODE> (one launch per step)",
            "kernel_ms": {"expand_train": float(kt[0]), "main": float(kt[1]), "refine": float(kt[2]),
                          "compact": float(kt[3])},
            "last_step_survivors": [int(c) for c in counts],
            "clocks": clocks.summary(),
        }
        if e2e:
            line["e2e"] = e2e
        if gathered is not None:
            line["nccl_gathered_matches"] = gathered
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(a, a.cpu_seconds)
        print(json.dumps(line), file=_REAL_STDOUT or sys.stdout, flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(a, ctx, seq, n, W, K, WU, world, dist, torch):
    """Same metric through the reference-facing C-ABI calls with HOST buffers.  Every step moves
    the new frame's descriptors host->device from pinned memory, runs the kernels, brings the
    ratio survivors back, sorts + cuts them (best_percent) into FeatureMatch lists in host
    memory and pushes the frame into the window.  Three variants:
      pipelined  vsf_window_submit(t) ... vsf_window_collect(t - lag): the host sorts frame t-lag
                 while the device matches frame t (headline; host std::sort = reference order)
      sync       vsf_window_feature_matches + vsf_window_commit, one blocking call per frame
      device_sort  the blocking call with the stable device sort"""
    L = ctx._L
    Ke = min(K, 300)
    n_host = Ke + WU + W + 1
    host = torch.empty((n_host, n, 32), dtype=torch.uint8).pin_memory()
    host.copy_(seq[:n_host])
    torch.cuda.synchronize()
    hp = host.numpy()
    fids = np.zeros(W, np.uint64)
    counts = np.zeros(W, np.int32)
    out = np.zeros((W, n), dtype=[("a", "<u8"), ("b", "<u8")])
    nf = C.c_int(0)
    fid = C.c_uint64(0)
    h2d_b, d2h_b = C.c_size_t(0), C.c_size_t(0)
    from vision_slam_frontend_b200 import PIPELINE_DEPTH
    lag = max(1, min(a.e2e_lag, PIPELINE_DEPTH - 1))
    res = {}

    def check(rc):
        if rc:
            raise RuntimeError(L.vsf_last_error(ctx._h).decode())

    def timed(step, drain=None, after_warmup=None):
        for t in range(WU):
            step(t)
        if drain:
            drain()
        if after_warmup:
            after_warmup()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(Ke):
            step(WU + t)
        if drain:
            drain()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        tt = torch.tensor([el], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    for sort_mode, key in ((1, "sync_exact_stdsort"), (0, "sync_device_sort")):
        ctx.window_clear()
        for p in range(W):
            ctx.window_push(p, hp[p])
        acc = {"d2h": 0}

        def step(t):
            D = hp[W + t]
            check(L.vsf_window_feature_matches(ctx._h, D.ctypes.data, n, 32, RATIO, BEST_PERCENT,
                                               sort_mode, fids.ctypes.data, counts.ctypes.data,
                                               out.ctypes.data, n, C.byref(nf)))
            L.vsf_window_commit(ctx._h, W + t, n)
            L.vsf_window_last_transfer(ctx._h, C.byref(h2d_b), C.byref(d2h_b))
            acc["d2h"] += d2h_b.value

        el = timed(step)
        res[key] = {"value": world * Ke * W * n * n / el, "ms_per_step": 1e3 * el / Ke,
                    "d2h_bytes_per_step": acc["d2h"] / (Ke + WU)}

    for sort_mode, key in ((1, "pipelined_exact_stdsort"), (0, "pipelined_device_sort")):
        ctx.window_clear()
        for p in range(W):
            ctx.window_push(p, hp[p])
        acc = {"d2h": 0, "collected": 0}

        def collect():
            check(L.vsf_window_collect(ctx._h, C.byref(fid), fids.ctypes.data, counts.ctypes.data,
                                       out.ctypes.data, n, C.byref(nf)))
            L.vsf_window_last_transfer(ctx._h, C.byref(h2d_b), C.byref(d2h_b))
            acc["d2h"] += d2h_b.value
            acc["collected"] += 1

        def step(t):
            D = hp[W + t]
            t0 = time.perf_counter()
            check(L.vsf_window_submit(ctx._h, W + t, D.ctypes.data, n, 32, RATIO, BEST_PERCENT, sort_mode,
                                      1))   # VSF_SUBMIT_PINNED_DESC: hp is page-locked
            t1 = time.perf_counter()
            if L.vsf_window_in_flight(ctx._h) > lag:
                collect()
            acc["submit_s"] = acc.get("submit_s", 0.0) + (t1 - t0)
            acc["collect_s"] = acc.get("collect_s", 0.0) + (time.perf_counter() - t1)

        def drain():
            while L.vsf_window_in_flight(ctx._h) > 0:
                collect()

        def reset():
            acc.update(d2h=0, collected=0, submit_s=0.0, collect_s=0.0)

        el = timed(step, drain, reset)
        assert acc["collected"] == Ke
        res[key] = {"value": world * Ke * W * n * n / el, "ms_per_step": 1e3 * el / Ke,
                    "d2h_bytes_per_step": acc["d2h"] / Ke, "frames_in_flight": lag + 1,
                    "host_submit_us_per_step": 1e6 * acc["submit_s"] / Ke,
                    "host_collect_us_per_step": 1e6 * acc["collect_s"] / Ke}

    # headline e2e: pipelined, bit-identical order (host std::sort, like the reference)
    head = res["pipelined_exact_stdsort"]
    return {"value": head["value"], "unit": UNIT, "h2d_bytes_per_step": n * 32,
            "d2h_bytes_per_step": head["d2h_bytes_per_step"], "steps": Ke,
            "ms_per_step": head["ms_per_step"],
            "api": "vsf_window_submit / vsf_window_collect (%d frames in flight; sort_mode=1: host std::sort, "
                   "bit-identical order), pinned host buffers; every step's H2D, kernels, D2H, sort + cut "
                   "inside the timed region" % (lag + 1),
            "matched_frame_pairs_per_s": head["value"] / (n * n),
            "host_threads": max(2, min(16, (os.cpu_count() or 16) // world)),
            "variants": res}


def rank_is_zero():
    return int(os.environ.get("RANK", "0")) == 0


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
