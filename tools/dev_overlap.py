"""Developer probe: do the normals (issue bound) and ICP (latency bound) kernels overlap usefully on two streams?"""
import sys
import threading
import time

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

P = 48
seq = synth.Sequence(2 * (P + 1), synth.OS1_64, start=30.0, workers=8)
A, B = engine.Engine(0), engine.Engine(0)
pp = A.make_preprocess_params()
ip = A.make_icp_params(engine.P2PLANE)
ids = np.arange(P + 1)
for k in ids:
    A.upload(k, seq.scans[k])
    B.upload(k, seq.scans[P + 1 + k])
initB = np.array([seq.relative_odo(P + 1 + a, P + 2 + a) for a in range(P)])
B.preprocess(ids, pp)
B.sync()


def prep():
    A.invalidate(ids)
    A.preprocess(ids, pp)
    A.sync()


def icp():
    B.icp_batch(ids[:-1], ids[1:], initB, ip)


for rep in range(3):
    t0 = time.perf_counter(); prep(); t1 = time.perf_counter(); icp(); t2 = time.perf_counter()
    th = [threading.Thread(target=prep), threading.Thread(target=icp)]
    t3 = time.perf_counter()
    for t in th:
        t.start()
    for t in th:
        t.join()
    t4 = time.perf_counter()
    print("rep %d: preprocess %.1f ms, icp %.1f ms, sequential %.1f ms, concurrent %.1f ms" %
          (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t2 - t0) * 1e3, (t4 - t3) * 1e3))
