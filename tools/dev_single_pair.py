"""Developer probe: where a one-pair registration call spends its time (per-kernel CUDA events, unrolled loop)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

seq = synth.Sequence(3, synth.OS1_64, start=30.0)
eng = engine.Engine(0)
pp = eng.make_preprocess_params()
ip = eng.make_icp_params(engine.P2PLANE)
for k in range(3):
    eng.upload(k, seq.scans[k])
eng.preprocess([0, 1, 2], pp)
init = np.array([seq.relative_odo(0, 1)])
for graph in (0, 1):
    eng.set_option("icp_loop_graph", graph)
    for rep in range(3):
        eng.icp_batch([0], [1], init, ip)
    eng.profile_enable(True)
    t0 = time.perf_counter()
    r = eng.icp_batch([0], [1], init, ip)
    t1 = time.perf_counter()
    prof = eng.profile_report()
    eng.profile_enable(False)
    print("graph=%d call %.3f ms passes %d" % (graph, (t1 - t0) * 1e3, r["passes"][0]))
    for k, v in sorted(prof.items()):
        print("   %-14s n=%3d total %.3f ms avg %.1f us" % (k, v[0], v[1], v[1] / v[0] * 1e3))
# one scan's preprocessing, kernel by kernel
eng.invalidate([2])
eng.sync()
eng.profile_enable(True)
t0 = time.perf_counter()
eng.preprocess([2], pp)
eng.sync()
t1 = time.perf_counter()
prof = eng.profile_report()
eng.profile_enable(False)
print("preprocess of one scan: %.3f ms host; kernels %.3f ms" % ((t1 - t0) * 1e3, sum(v[1] for v in prof.values())))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print("   %-16s n=%3d total %.3f ms" % (k, v[0], v[1]))
ts = []
for rep in range(20):
    t0 = time.perf_counter()
    eng.upload(2, seq.scans[2])
    t1 = time.perf_counter()
    eng.preprocess([2], pp)
    t2 = time.perf_counter()
    eng.sync()
    t3 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1, t3 - t2))
ts = np.median(np.array(ts), axis=0) * 1e3
print("host: upload call %.3f ms, preprocess call %.3f ms, sync %.3f ms" % tuple(ts))
eng.close()
