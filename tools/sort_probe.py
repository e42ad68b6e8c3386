"""Times the device sort + cut (sort_mode 0 / 2) on the match lists of a real C4 pose: run under
`ncu -k regex:sort_cut --metrics gpu__time_duration.sum` for the kernel's own duration, or alone
for the host-API time of vsf_window_feature_matches in each mode."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import synth
import vision_slam_frontend_b200 as vsf

RATIO = float(np.float32(0.6))
n, W = int(os.environ.get("N", 5000)), int(os.environ.get("W", 10))
frames = [synth.synth_pose(n, p, n // 10, 20240917) for p in range(W + 1)]
with vsf.Context(max_features=n, window=W) as ctx:
    for p in range(W):
        ctx.window_push(p, frames[p])
    for mode in (0, 2, 1):
        for _ in range(3):
            got = ctx.window_feature_matches(frames[W], RATIO, 0.3, mode)
        t0 = time.perf_counter()
        for _ in range(20):
            got = ctx.window_feature_matches(frames[W], RATIO, 0.3, mode)
        print("sort_mode %d: %.1f us per blocking call, kept %s" % (mode, (time.perf_counter() - t0) / 20 * 1e6,
                                                                    [len(p) for _, p in got]), flush=True)
