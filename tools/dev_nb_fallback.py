"""Developer probe: how many points of the block-cooperative normals kernel end up on the fallback list, and why."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
seq = synth.Sequence(n, synth.OS1_64, start=0.0, step=4.0, workers=os.cpu_count())
eng = engine.Engine(0)
pp = eng.make_preprocess_params()
for k in range(n):
    eng.upload(k, seq.scans[k])
eng.preprocess(list(range(n)), pp)
tot = np.zeros(16, dtype=np.int64)
for k in range(n):
    tot += np.array(list(eng.get_counters(k).values()), dtype=np.int64)
npts = sum(eng.info(k)["n_points"] for k in range(n))
print("points %d | fallback list %d (%.2f %%) | whole blocks %d (~%d pts) | single points %d | trial blocks %d"
      % (npts, tot[5], 100.0 * tot[5] / npts, tot[6], tot[5] - tot[7], tot[7], tot[8]))
eng.invalidate(list(range(n)))
eng.sync()
eng.profile_enable(True)
eng.preprocess(list(range(n)), pp)
prof = eng.profile_report()
eng.profile_enable(False)
for k in ("normals", "normals_fallback", "normals_redo", "normals_eigen"):
    if k in prof:
        print("   %-18s %.3f ms" % (k, prof[k][1]))
eng.close()
