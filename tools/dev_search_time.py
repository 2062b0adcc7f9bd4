"""Developer probe: ICP kernel times of a 40-scan consecutive batch (ARVC_LIB_VARIANT selects a build)."""
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seq = synth.Sequence(n, synth.OS1_64, start=30.0, workers=8)
eng = engine.Engine(0)
eng.set_option("icp_loop_graph", 0)
ids = list(range(n))
for k in ids:
    eng.upload(k, seq.scans[k])
eng.preprocess(ids, eng.make_preprocess_params())
init = np.array([seq.relative_odo(a, a + 1) for a in ids[:-1]])
ip = eng.make_icp_params(engine.P2PLANE)
best = None
for rep in range(4):
    eng.sync()
    eng.profile_enable(True)
    r = eng.icp_batch(ids[:-1], ids[1:], init, ip)
    prof = eng.profile_report()
    eng.profile_enable(False)
    t = {"search": sum(v[1] for k, v in prof.items() if k.startswith("icp_pass")), "far": sum(v[1] for k, v in prof.items() if k.startswith("icp_far")), "select": prof["icp_select"][1],
         "accum": prof["icp_accum"][1], "finish": prof["icp_finish"][1]}
    if best is None or sum(t.values()) < sum(best.values()):
        best = t
print("%-10s per pair (us): %s  total %.1f" % (os.path.basename(os.environ.get("ARVC_LIB_VARIANT", "default")),
                                              {k: round(v * 1e3 / (n - 1), 1) for k, v in best.items()}, sum(best.values()) * 1e3 / (n - 1)))
eng.close()
