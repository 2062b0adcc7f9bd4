"""Developer timing probe (not the bench contract): stage times of a small 64-beam batch on cuda:0."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 9
sensor = {"64": synth.OS1_64, "128": synth.OS_128, "32": synth.SMALL_32}[sys.argv[2] if len(sys.argv) > 2 else "64"]
method = engine.P2P if (len(sys.argv) > 3 and sys.argv[3] == "p2p") else engine.P2PLANE
voxel = float(sys.argv[4]) if len(sys.argv) > 4 else None
t0 = time.time()
seq = synth.Sequence(n_scans, sensor, start=30.0)
print("generated %d scans in %.1fs, %d pts each" % (n_scans, time.time() - t0, len(seq.scans[0])))
eng = engine.Engine(0)
pp = eng.make_preprocess_params(voxel_size=voxel, want_normals=method == engine.P2PLANE)
ip = eng.make_icp_params(method)
ids = list(range(n_scans))
tg, sr = ids[:-1], ids[1:]
init = np.array([seq.relative_odo(a, b) for a, b in zip(tg, sr)])
for rep in range(3):
    eng.sync()
    t0 = time.perf_counter()
    for k in ids:
        eng.upload(k, seq.scans[k])
    eng.sync()
    t1 = time.perf_counter()
    eng.preprocess(ids, pp)
    eng.sync()
    t2 = time.perf_counter()
    res = eng.icp_batch(tg, sr, init, ip)
    t3 = time.perf_counter()
    print("rep %d: upload %.2f ms  preprocess %.2f ms (%.3f ms/scan)  icp %.2f ms (%.3f ms/pair)  -> %.1f pairs/s; updates %s"
          % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t1) * 1e3 / n_scans, (t3 - t2) * 1e3, (t3 - t2) * 1e3 / len(tg),
             len(tg) / (t3 - t0), res["updates"].tolist()))
print("info", eng.info(0), "launches", eng.kernel_launches())
print("rmse", res["rmse"][:4], "fitness", res["fitness"][:4])
