"""Small end-to-end run for compute-sanitizer: tiny scans, both methods, voxel and plain, trace API."""
import sys

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

seq = synth.Sequence(3, synth.TINY_16, start=30.0)
eng = engine.Engine(0)
for voxel in (None, 0.3):
    for k in range(3):
        eng.upload(k, seq.scans[k])
    eng.preprocess([0, 1, 2], eng.make_preprocess_params(voxel_size=voxel))
    init = np.array([seq.relative_odo(0, 1), seq.relative_odo(1, 2)])
    for method in (engine.P2PLANE, engine.P2P):
        r = eng.icp_batch([0, 1], [1, 2], init, eng.make_icp_params(method))
        print(voxel, method, r["updates"].tolist(), r["rmse"].tolist())
    tr = eng.icp_trace(0, 1, init[0], eng.make_icp_params(engine.P2PLANE))
    print("trace passes", tr["passes"])
eng.upload(9, np.zeros((0, 3), dtype=np.float32))
eng.preprocess([9], eng.make_preprocess_params())
print(eng.icp_batch([0], [9], np.eye(4)[None], eng.make_icp_params(engine.P2PLANE))["fitness"])
eng.close()
print("done")
