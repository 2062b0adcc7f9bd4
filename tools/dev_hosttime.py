"""Developer probe: host-side enqueue time vs device time of one hot-path step (33 scans / 32 pairs)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

P = 32
seq = synth.Sequence(P + 1, synth.OS1_64, start=30.0)
eng = engine.Engine(0)
pp = eng.make_preprocess_params()
ip = eng.make_icp_params(engine.P2PLANE)
ids = np.arange(P + 1)
init = np.array([seq.relative_odo(a, a + 1) for a in range(P)])
for k in ids:
    eng.upload(k, seq.scans[k])
for rep in range(8):
    eng.sync()
    t0 = time.perf_counter()
    eng.invalidate(ids)
    t1 = time.perf_counter()
    eng.preprocess(ids, pp)
    t2 = time.perf_counter()
    tk = eng.icp_batch_async(ids[:-1], ids[1:], init, ip)
    t3 = time.perf_counter()
    res = eng.icp_batch_finish(tk)
    t4 = time.perf_counter()
    print("rep %d: invalidate %.2f ms | preprocess enqueue %.2f ms | icp enqueue %.2f ms | wait %.2f ms | total %.2f ms"
          % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3))
