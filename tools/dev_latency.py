"""Developer probe: sequential scan-matcher latency (one new scan preprocessed + one pair registered per iteration),
the call pattern of run_scanmatcher.py:196-213."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

n = 12
seq = synth.Sequence(n, synth.OS1_64, start=30.0, workers=8)
eng = engine.Engine(0)
pp = eng.make_preprocess_params()
ip = eng.make_icp_params(engine.P2PLANE)
for rep in range(2):
    eng.upload(0, seq.scans[0]); eng.preprocess([0], pp); eng.sync()
    ts = []
    for i in range(n - 1):
        t0 = time.perf_counter()
        eng.upload(i + 1, seq.scans[i + 1])
        eng.preprocess([i + 1], pp)
        r = eng.icp_batch([i], [i + 1], seq.relative_odo(i, i + 1)[None], ip)
        eng.free(i)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("rep %d: ms per pair (upload + preprocess + icp + free): median %.2f  min %.2f  max %.2f" % (rep, np.median(ts), min(ts), max(ts)), [round(t, 2) for t in ts])
