"""Developer probe of round 2 (not the bench contract): normals kernel A/B + counters, ICP loop graph vs unrolled,
single-pair latency.  Usage: python tools/dev_r2.py [n_scans]   (ARVC_NORMALS_IMPL=point for the per-point kernel)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seq = synth.Sequence(n_scans, synth.OS1_64, start=30.0, workers=8)
eng = engine.Engine(0)
pp = eng.make_preprocess_params()
ip = eng.make_icp_params(engine.P2PLANE)
ids = list(range(n_scans))
tg, sr = ids[:-1], ids[1:]
init = np.array([seq.relative_odo(a, b) for a, b in zip(tg, sr)])
for k in ids:
    eng.upload(k, seq.scans[k])
print("impl", os.environ.get("ARVC_NORMALS_IMPL", "block"))
for rep in range(3):
    eng.invalidate(ids)
    eng.sync()
    eng.profile_enable(True)
    t0 = time.perf_counter()
    eng.preprocess(ids, pp)
    eng.sync()
    t1 = time.perf_counter()
    prof = eng.profile_report()
    eng.profile_enable(False)
    print("rep %d preprocess %.2f ms (%.3f ms/scan): %s" % (rep, (t1 - t0) * 1e3, (t1 - t0) * 1e3 / n_scans,
          {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:6]}))
cs = [eng.get_counters(k) for k in ids[:4]]
for c in cs:
    print("counters", c)
nn = eng.get_nn_counts(0)
pts, nrm = eng.get_points(0, normals=True)
np.savez(os.path.join("gpurun_out", "normals_%s.npz" % os.environ.get("ARVC_NORMALS_IMPL", "block")), nn=nn, nrm=nrm)
for graph in (1, 0, 1):
    eng.set_option("icp_loop_graph", graph)
    for rep in range(3):
        eng.sync()
        l0 = eng.kernel_launches()
        t0 = time.perf_counter()
        res = eng.icp_batch(tg, sr, init, ip)
        t1 = time.perf_counter()
    print("icp graph=%d: %.2f ms (%.3f ms/pair) launches %d updates %s" % (graph, (t1 - t0) * 1e3, (t1 - t0) * 1e3 / len(tg), eng.kernel_launches() - l0,
          res["updates"].tolist()))
    if graph:
        rg = res
    else:
        assert np.array_equal(rg["T"], res["T"]) and np.array_equal(rg["updates"], res["updates"]), "graph and unrolled loops differ"
# small batches and the reference's sequential pattern (run_scanmatcher.py:196-213): upload, preprocess, 1 pair, free
for graph in (1, 0):
    eng.set_option("icp_loop_graph", graph)
    for nb in (1, 4, 16):
        ts = []
        for rep in range(5):
            t0 = time.perf_counter()
            eng.icp_batch(tg[:nb], sr[:nb], init[:nb], ip)
            ts.append(time.perf_counter() - t0)
        print("graph=%d batch of %2d pairs: %.3f ms (%.3f ms/pair)" % (graph, nb, min(ts) * 1e3, min(ts) * 1e3 / nb))
    e2 = engine.Engine(0)
    e2.set_option("icp_loop_graph", graph)
    e2.upload(0, seq.scans[0]); e2.preprocess([0], pp)
    ts = []
    for k in range(1, n_scans):
        t0 = time.perf_counter()
        e2.upload(k, seq.scans[k])
        e2.preprocess([k], pp)
        r = e2.icp_batch([k - 1], [k], init[k - 1:k], ip)
        e2.free(k - 1)
        ts.append(time.perf_counter() - t0)
    print("graph=%d sequential pattern: median %.3f ms/pair, min %.3f" % (graph, np.median(ts) * 1e3, min(ts) * 1e3))
    e2.close()
eng.close()
