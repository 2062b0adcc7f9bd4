"""Small fixed workload for ncu captures: 9 OS1-64 scans, 8 consecutive point-to-plane pairs, one hot-path pass."""
import sys

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 9
seq = synth.Sequence(n_scans, synth.OS1_64, start=30.0)
eng = engine.Engine(0)
ids = list(range(n_scans))
for k in ids:
    eng.upload(k, seq.scans[k])
eng.preprocess(ids, eng.make_preprocess_params())
init = np.array([seq.relative_odo(a, a + 1) for a in ids[:-1]])
res = eng.icp_batch(ids[:-1], ids[1:], init, eng.make_icp_params(engine.P2PLANE))
print("updates", res["updates"].tolist(), "rmse", float(res["rmse"].mean()))
