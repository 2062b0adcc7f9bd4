"""Developer probe: configs[2] (128-beam point-to-point, 12 consecutive pairs) device time per step, kernel by kernel."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

n3 = 13
seq = synth.Sequence(n3, synth.OS_128, start=30.0, workers=os.cpu_count())
eng = engine.Engine(0)
ids = np.arange(n3)
init = np.array([seq.relative_odo(k, k + 1) for k in range(n3 - 1)])
pp = eng.make_preprocess_params(want_normals=False)
ip = eng.make_icp_params(engine.P2P)
for k in ids:
    eng.upload(int(k), seq.scans[k])


def step():
    eng.invalidate(ids)
    eng.preprocess(ids, pp)
    return eng.icp_batch(ids[:-1], ids[1:], init, ip)


for _ in range(3):
    step()
eng.sync()
t0 = time.perf_counter()
for _ in range(5):
    r = step()
eng.sync()
ms = (time.perf_counter() - t0) / 5 * 1e3
print("%-10s config 3: %.2f ms/step  %.1f pairs/s  mean updates %.2f" % (os.path.basename(os.environ.get("ARVC_LIB_VARIANT", "default")), ms, (n3 - 1) / ms * 1e3, r["updates"].mean()))
eng.set_option("icp_loop_graph", 0)
step()
eng.sync()
eng.profile_enable(True)
step()
prof = eng.profile_report()
agg = {}
for k, v in prof.items():
    base = k.rstrip("0123456789_") if k.startswith("icp_") else k
    agg[base] = agg.get(base, 0.0) + v[1]
print("   " + "  ".join("%s %.2f" % kv for kv in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
eng.close()
