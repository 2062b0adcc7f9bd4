import sys
import numpy as np
sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth
from oracle import oracle as orc
seq = synth.Sequence(3, synth.SMALL_32, start=12.0)
eng = engine.Engine(0)
for max_nn, radius in [(300, 0.3), (20, 0.3)]:
    eng.upload(4, seq.scans[0])
    eng.preprocess([4], eng.make_preprocess_params(normal_radius=radius, max_nn=max_nn))
    pts, nrm = eng.get_points(4, normals=True)
    on, cov, cnt = orc.estimate_normals(pts, radius, max_nn, return_cov=True)
    gc = eng.get_nn_counts(4)
    print("count mismatch", (gc != cnt).sum())
    w = np.linalg.eigvalsh(cov)
    gap = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2], 1e-300)
    err = np.minimum(np.linalg.norm(nrm - on, axis=1), np.linalg.norm(nrm + on, axis=1))
    err_signed = np.linalg.norm(nrm - on, axis=1)
    print("max_nn", max_nn, "n", len(pts), "err>1e-6:", (err > 1e-6).sum(), "signflips:", (err_signed > 1.0).sum())
    bad = np.where(err > 1e-6)[0]
    for i in bad[:15]:
        print(i, "cnt", cnt[i], "gap %.3e" % gap[i], "w", w[i], "err %.3e" % err[i], nrm[i], on[i])
    print("pcts of err", np.percentile(err, [50, 90, 99, 99.9, 100]))
    print("normnorm", np.abs(np.linalg.norm(nrm, axis=1) - 1).max())
    # relation of err with gap
    for lo, hi in [(0, 1e-8), (1e-8, 1e-6), (1e-6, 1e-4), (1e-4, 1e-2), (1e-2, 10)]:
        m = (gap >= lo) & (gap < hi)
        if m.any():
            print("gap [%g,%g): n=%d maxerr=%.3e" % (lo, hi, m.sum(), err[m].max()))
