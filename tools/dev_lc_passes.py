"""Developer probe: per-pass kernel times of one 2048-pair loop-closure batch (unrolled loop, per-launch events)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from lidar_slam_arvc_b200 import engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
wl = bench.LoopClosureWorkload(464, 10000, 16)
scans = np.unique(np.concatenate([wl.tg[:n], wl.sr[:n]]))
wl.materialise(scans)
eng = engine.Engine(0)
for k in scans:
    eng.upload(int(k), wl.scans[int(k)])
eng.preprocess(scans, eng.make_preprocess_params())
ip = eng.make_icp_params(engine.P2PLANE)
for graph in (1, 0):
    eng.set_option("icp_loop_graph", graph)
    eng.icp_batch(wl.tg[:n], wl.sr[:n], wl.init[:n], ip)
    eng.sync()
    eng.profile_enable(True)
    r = eng.icp_batch(wl.tg[:n], wl.sr[:n], wl.init[:n], ip)
    prof = eng.profile_report()
    eng.profile_enable(False)
    if graph:
        print("graph: %.1f ms for %d pairs" % (prof["icp_graph"][1], n))
        continue
    tot = sum(v[1] for v in prof.values())
    print("unrolled: %.1f ms;" % tot, {k: round(v[1], 1) for k, v in prof.items() if not k.startswith("icp_pass_") and not k.startswith("icp_far_")})
    print("far by pass:", [round(prof["icp_far_%02d" % p][1], 2) for p in range(31)])
    print("search by pass:", [round(prof["icp_pass_%02d" % p][1], 2) for p in range(31)])
    print("active pairs by pass:", [int((r["passes"] > p).sum()) for p in range(31)])
eng.close()
