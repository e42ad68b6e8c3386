"""racecheck / synccheck target for the round-2 kernels with intricate synchronisation: the exact
device sort (shared-memory phases separated by barriers) and the CTA-pair kernel (cluster
barriers, remote mbarrier arrivals)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import synth
import vision_slam_frontend_b200 as vsf
from oracle import native, restate
rng = np.random.default_rng(0)
with vsf.Context(device=0, max_features=4096, window=2) as ctx:
    for n in (40, 700, 3000):
        m = np.zeros(n, vsf.DMATCH_DTYPE)
        m["queryIdx"] = np.arange(n); m["trainIdx"] = np.arange(n); m["distance"] = rng.integers(0, 25, n).astype(np.float32)
        got = ctx.debug_sort_device(m, 0.3, True)
        keep = restate.num_good_matches(n, np.float32(0.3))
        exp = m[restate.sort_order_stdsort(m)][:keep]
        assert (got[:, 0] == exp["queryIdx"].astype(np.uint64)).all()
if os.environ.get("PAIR", "1") == "1":
    with vsf.Context(device=0, max_features=3600, desc_bytes=61, window=2) as ctx:
        ctx.set_engine(2, 0)
        Q, T = synth.descriptor_pair(3600, 3500, width=61, seed=12)
        assert (ctx.get_matches(Q, T, restate.NN_MATCH_RATIO) == native.get_matches(Q, T, restate.NN_MATCH_RATIO)).all()
print("ok")
