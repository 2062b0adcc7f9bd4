"""Developer probe: does arvc_ctx_reserve help or hurt when the pool already holds memory (the bench's situation)?"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_slam_arvc_b200 import engine, pipeline, synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "none"
n5 = 1500
seq = synth.Sequence(n5, synth.OS1_64, start=0.0, workers=os.cpu_count())
odo = [seq.relative_odo(k, k + 1) for k in range(n5 - 1)]
eng = engine.Engine(0)
pp = eng.make_preprocess_params()
# the bench's situation: ~460 scans were resident (configs[3]) and have been freed again
ids = list(range(10000, 10460))
for k in ids:
    eng.upload(k, seq.scans[k % n5])
eng.preprocess(ids, pp)
eng.sync()
for k in ids:
    eng.free(k)
eng.sync()
if mode == "reserve":
    eng.reserve(10 << 30)
for rep in range(2):
    eng.sync()
    t0 = time.perf_counter()
    rel, recs = pipeline.scan_matcher(eng, seq.scans, odo, batch=100)
    eng.sync()
    dt = time.perf_counter() - t0
    print("%s rep %d: front end %.3f s = %.0f pairs/s" % (mode, rep, dt, (n5 - 1) / dt))
eng.close()
