"""Developer probe: the back end of configs[4] (loop closing every 50 keyframes + host pose graph) on a shorter sequence."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_slam_arvc_b200 import engine, pipeline, synth  # noqa: E402

n5 = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
seq = synth.Sequence(n5, synth.OS1_64, start=0.0, workers=os.cpu_count())
odo = [seq.relative_odo(k, k + 1) for k in range(n5 - 1)]
eng = engine.Engine(0)
if os.environ.get('RESERVE_GB'):
    eng.reserve(int(float(os.environ['RESERVE_GB']) * 2**30))
t0 = time.perf_counter()
rel, recs = pipeline.scan_matcher(eng, seq.scans, odo, batch=100)
eng.sync()
print("front end %.3f s" % (time.perf_counter() - t0))
for rep_i in range(1):
    rep = pipeline.run_backend(eng, seq.scans, rel, odo, skip_loop_closing=50, skip_optimization=50, number_of_triplets_loop_closing=20,
                               distance_backwards=7.0, radius_threshold=5.0, seed=0)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in rep.items()})
eng.close()
