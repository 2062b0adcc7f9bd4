import cProfile, pstats, os, sys, io
sys.path.insert(0, "/root/repo")
import numpy as np
from lidar_slam_arvc_b200 import engine, pipeline, synth
n5 = 5000
seq = synth.Sequence(n5, synth.OS1_64, start=0.0, workers=os.cpu_count())
odo = [seq.relative_odo(k, k + 1) for k in range(n5 - 1)]
eng = engine.Engine(0)
eng.reserve(10 << 30)
rel, recs = pipeline.scan_matcher(eng, seq.scans, odo, batch=100)
eng.sync()
pr = cProfile.Profile()
pr.enable()
rep = pipeline.run_backend(eng, seq.scans, rel, odo, skip_loop_closing=50, skip_optimization=50, number_of_triplets_loop_closing=20, distance_backwards=7.0, radius_threshold=5.0, seed=0)
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in rep.items()})
