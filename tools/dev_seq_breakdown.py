"""Developer probe: where one iteration of the reference's one-pair-per-call loop spends its host time."""
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lidar_slam_arvc_b200 import engine, euroc_synth, runtime, synth  # noqa: E402
from lidar_slam_arvc_b200.homogeneousmatrix import HomogeneousMatrix  # noqa: E402

n_scans = 40
seq = synth.Sequence(n_scans, synth.OS1_64, start=30.0, workers=os.cpu_count())
sys.path.insert(0, os.path.join(ROOT, "lidar_slam_arvc_b200", "dropin"))
eng = engine.Engine(0)
runtime.set_engine(eng)
import keyframemanager.keyframemanager as kfm  # noqa: E402

with tempfile.TemporaryDirectory() as d:
    scan_times = euroc_synth.write_euroc_tree(d, seq)
    odo = [HomogeneousMatrix(seq.relative_odo(i, i + 1)) for i in range(n_scans - 1)]
    rows = []
    # finer split of compute_transformation: enqueue / stage the next scan / wait
    acc = {"enqueue": [], "stage": [], "finish": []}
    ld = runtime.get_loader()

    def timed(name, fn):
        def w(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            acc[name].append(time.perf_counter() - t0)
            return r
        return w
    eng.icp_batch_async = timed("enqueue", eng.icp_batch_async)
    eng.icp_batch_finish = timed("finish", eng.icp_batch_finish)
    ld.stage_ahead = timed("stage", ld.stage_ahead)
    with contextlib.redirect_stdout(io.StringIO()):
        km = kfm.KeyFrameManager(directory=d, scan_times=scan_times, voxel_size=None, method="icppointplane")
        km.add_keyframe(0)
        km.load_pointcloud(0)
        km.pre_process(0)
        for i in range(n_scans - 1):
            t = [time.perf_counter()]
            km.add_keyframe(i + 1); t.append(time.perf_counter())
            km.load_pointcloud(i + 1); t.append(time.perf_counter())
            km.pre_process(i + 1); t.append(time.perf_counter())
            km.compute_transformation(i, i + 1, Tij=odo[i]); t.append(time.perf_counter())
            km.unload_pointcloud(i); t.append(time.perf_counter())
            rows.append(np.diff(t))
    rows = np.array(rows[3:]) * 1e3
    med = np.median(rows, axis=0)
    print("median ms: add %.3f | load %.3f | pre_process %.3f | compute_transformation %.3f | unload %.3f | total %.3f"
          % (*med, np.median(rows.sum(axis=1))))
    print("   compute_transformation split (median ms): " + " | ".join("%s %.3f" % (k, np.median(v[3:]) * 1e3) for k, v in acc.items()))
    print("   loader stats:", ld.stats)
    # device time of the same iteration, kernel by kernel
    eng.set_option("icp_loop_graph", 0)
    eng.profile_enable(True)
    with contextlib.redirect_stdout(io.StringIO()):
        km2 = kfm.KeyFrameManager(directory=d, scan_times=scan_times, voxel_size=None, method="icppointplane")
        km2.add_keyframe(0); km2.load_pointcloud(0); km2.pre_process(0)
        for i in range(10):
            km2.add_keyframe(i + 1); km2.load_pointcloud(i + 1); km2.pre_process(i + 1)
            km2.compute_transformation(i, i + 1, Tij=odo[i]); km2.unload_pointcloud(i)
    prof = eng.profile_report()
    eng.profile_enable(False)
    agg = {}
    for k, v in prof.items():
        base = k.rstrip("0123456789_") if k.startswith("icp_") else k
        a = agg.setdefault(base, [0, 0.0]); a[0] += v[0]; a[1] += v[1]
    tot = sum(v[1] for v in agg.values())
    print("device kernels per pair (unrolled loop, profiled): %.3f ms" % (tot / 10))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("   %-18s n/pair %5.1f  %.3f ms/pair" % (k, v[0] / 10, v[1] / 10))
runtime.set_engine(None)
eng.close()
