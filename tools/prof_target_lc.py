"""Small fixed loop-closure workload for ncu captures: N pairs of BASELINE configs[3] (seed 777), one registration batch."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from lidar_slam_arvc_b200 import engine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
wl = bench.LoopClosureWorkload(464, 10000, os.cpu_count() or 1)
tg, sr, init = wl.tg[:N], wl.sr[:N], wl.init[:N]
scans = np.unique(np.concatenate([tg, sr]))
wl.materialise(scans)
eng = engine.Engine(0)
for k in scans:
    eng.upload(int(k), wl.scans[int(k)])
eng.preprocess(scans, eng.make_preprocess_params())
res = eng.icp_batch(tg, sr, init, eng.make_icp_params(engine.P2PLANE))
print("pairs", N, "scans", len(scans), "mean updates", float(res["updates"].mean()))
eng.close()
