"""Developer probe: per-pair ICP passes of the full config-4 list (for offline study of sharding strategies)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from lidar_slam_arvc_b200 import engine  # noqa: E402

wl = bench.LoopClosureWorkload(464, 10000, 16)
scans = np.unique(np.concatenate([wl.tg, wl.sr]))
wl.materialise(scans)
eng = engine.Engine(0)
for k in scans:
    eng.upload(int(k), wl.scans[int(k)])
eng.preprocess(scans, eng.make_preprocess_params())
ip = eng.make_icp_params(engine.P2PLANE)
out = []
t0 = time.perf_counter()
for lo in range(0, len(wl.tg), 2000):
    r = eng.icp_batch(wl.tg[lo:lo + 2000], wl.sr[lo:lo + 2000], wl.init[lo:lo + 2000], ip)
    out.append(r)
    print(lo, time.perf_counter() - t0, flush=True)
r = np.concatenate(out)
np.savez("gpurun_out/lc_costs.npz", tg=wl.tg, sr=wl.sr, passes=r["passes"], updates=r["updates"], fitness=r["fitness"], rmse=r["rmse"])
print("mean updates", r["updates"].mean(), "hist", np.bincount(r["updates"]))
