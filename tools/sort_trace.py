"""Per-phase cycle counts of the exact device sort (VSF_SORT_TRACE=1) on tie-heavy lists."""
import os, sys
os.environ["VSF_SORT_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import vision_slam_frontend_b200 as vsf
rng = np.random.default_rng(0)
with vsf.Context(max_features=8192, window=2) as ctx:
    for n in (4500, 4500, 2000, 500):
        m = np.zeros(n, vsf.DMATCH_DTYPE)
        m["queryIdx"] = np.arange(n); m["trainIdx"] = np.arange(n)
        m["distance"] = np.clip(rng.normal(15, 4, n), 0, 60).astype(np.int32).astype(np.float32)
        for exact in (True, False):
            ctx.debug_sort_device(m, 0.3, exact)
