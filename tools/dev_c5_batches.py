"""Developer probe: the loop-closing registration batches of configs[4] one by one (pairs, passes, time)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_slam_arvc_b200 import engine, pipeline, synth  # noqa: E402

n5 = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
seq = synth.Sequence(n5, synth.OS1_64, start=0.0, workers=os.cpu_count())
odo = [seq.relative_odo(k, k + 1) for k in range(n5 - 1)]
eng = engine.Engine(0)
eng.reserve(10 << 30)
rel, recs = pipeline.scan_matcher(eng, seq.scans, odo, batch=100)
eng.sync()
log = []
orig_async, orig_finish, orig_pre = eng.icp_batch_async, eng.icp_batch_finish, eng.preprocess
t_start = {}


def a(tg, sr, init, ip):
    eng.sync()
    t0 = time.perf_counter()
    tk = orig_async(tg, sr, init, ip)
    t_start[tk[0]] = (t0, np.asarray(init))
    return tk


def f(tk):
    r = orig_finish(tk)
    t0, init = t_start.pop(tk[0])
    dt = time.perf_counter() - t0
    d = np.linalg.norm(r["T"][:, :3, 3] - init[:, :3, 3], axis=1)
    log.append((len(r), dt * 1e3, r["passes"].mean(), r["passes"].max(), np.median(d), d.max(), r["fitness"].min()))
    return r


pre_ms = [0.0, 0]


def p(ids, pp):
    eng.sync()
    t0 = time.perf_counter()
    orig_pre(ids, pp)
    eng.sync()
    pre_ms[0] += (time.perf_counter() - t0) * 1e3
    pre_ms[1] += len(np.atleast_1d(ids))


eng.icp_batch_async, eng.icp_batch_finish, eng.preprocess = a, f, p
rep = pipeline.run_backend(eng, seq.scans, rel, odo, skip_loop_closing=50, skip_optimization=50, number_of_triplets_loop_closing=20,
                           distance_backwards=7.0, radius_threshold=5.0, seed=0)
print("preprocess: %.1f ms for %d scan requests" % tuple(pre_ms))
print("pairs  ms   ms/pair  passes mean/max   |t - t_init| median/max   min fitness")
for row in log:
    print("%4d %7.2f %7.3f   %5.1f / %2d      %.3f / %.3f   %.3f" % (row[0], row[1], row[1] / row[0], row[2], row[3], row[4], row[5], row[6]))
print("total icp ms %.1f" % sum(r[1] for r in log))
eng.close()
