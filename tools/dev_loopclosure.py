"""Developer probe: BASELINE config 4 regime at full scan size - loop-closure-like pairs (2.5-5 m apart, initial guess
perturbed by N(0, 0.2 m) / N(0, 2 deg)) on resident, preprocessed 64-beam scans: ICP-only throughput and iteration counts."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, sharding, synth  # noqa: E402

n_scans, n_pairs = 60, int(sys.argv[1]) if len(sys.argv) > 1 else 300
seq = synth.Sequence(n_scans, synth.OS1_64, start=30.0, workers=8)
pairs = synth.loop_closure_pairs(seq.poses, n_pairs, radius=5.0, min_gap=5, seed=777)
eng = engine.Engine(0)
for k, s in enumerate(seq.scans):
    eng.upload(k, s)
eng.preprocess(list(range(n_scans)), eng.make_preprocess_params())
order = sharding.sort_pairs_for_cache([p[0] for p in pairs], [p[1] for p in pairs])
tg = np.array([pairs[k][0] for k in order]); sr = np.array([pairs[k][1] for k in order]); init = np.array([pairs[k][2] for k in order])
ip = eng.make_icp_params(engine.P2PLANE)
eng.profile_enable(False)
for rep in range(3):
    eng.sync()
    t0 = time.perf_counter()
    res = eng.icp_batch(tg, sr, init, ip)
    dt = time.perf_counter() - t0
    print("rep %d: %d pairs in %.1f ms -> %.0f pairs/s (ICP only); updates mean %.1f max %d; fitness min %.3f; rmse mean %.3f"
          % (rep, len(tg), dt * 1e3, len(tg) / dt, res["updates"].mean(), res["updates"].max(), res["fitness"].min(), res["rmse"].mean()))
gt_err = [np.linalg.norm(res["T"][k][:3, 3] - (np.linalg.inv(seq.poses[tg[k]]) @ seq.poses[sr[k]])[:3, 3]) for k in range(len(tg))]
print("translation error vs ground truth: median %.3f m, 90%% %.3f m, max %.3f m" % (np.median(gt_err), np.percentile(gt_err, 90), max(gt_err)))
eng.profile_enable(True)
res = eng.icp_batch(tg, sr, init, ip)
rep = eng.profile_report()
tot = sum(v[1] for v in rep.values())
print("kernels:", {k: round(v[1], 2) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:8]}, "total %.1f ms" % tot)
