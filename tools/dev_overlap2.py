"""Developer probe: which call of the load path blocks behind a running registration batch."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

seq = synth.Sequence(8, synth.OS1_64, start=30.0)
eng = engine.Engine(0)
pp = eng.make_preprocess_params()
for k in (0, 1):
    eng.upload(k, seq.scans[k])
eng.preprocess([0, 1], pp)
eng.sync()
ip = eng.make_icp_params()
n_rep = 200
init = np.repeat(seq.relative_odo(0, 1)[None], n_rep, axis=0)
for mode in ("pageable", "pinned-prealloc", "pinned-alloc-inside", "pinned-fresh-cudaMallocHost-inside"):
    bufs = []
    if mode == "pinned-prealloc":
        for k in range(2, 6):
            a, h = eng.pinned.empty(len(seq.scans[k]), np.float32)
            a[:] = seq.scans[k]
            bufs.append((a, h))
    eng.sync()
    t0 = time.perf_counter()
    ticket = eng.icp_batch_async([0] * n_rep, [1] * n_rep, init, ip)
    t1 = time.perf_counter()
    marks = []
    for j, k in enumerate(range(2, 6)):
        if mode == "pageable":
            eng.upload(10 + k, seq.scans[k])
        elif mode == "pinned-prealloc":
            eng.upload(10 + k, bufs[j][0])
        else:
            a, h = eng.pinned.empty(len(seq.scans[k]) * (1 if mode == "pinned-alloc-inside" else 2 + k), np.float32)
            a = a[:len(seq.scans[k])]
            marks.append(("alloc", time.perf_counter() - t0))
            a[:] = seq.scans[k]
            bufs.append((a, h))
            eng.upload(10 + k, a)
        marks.append(("upload%d" % k, time.perf_counter() - t0))
    for k in range(2, 6):
        eng.wait_upload(10 + k)
        marks.append(("wait%d" % k, time.perf_counter() - t0))
    eng.icp_batch_finish(ticket)
    t_icp = time.perf_counter() - t0
    print(mode, "enqueue %.2f ms, icp done %.2f ms:" % ((t1 - t0) * 1e3, t_icp * 1e3), " ".join("%s=%.2f" % (n, t * 1e3) for n, t in marks))
    for k in range(2, 6):
        eng.free(10 + k)
    for a, h in bufs:
        eng.pinned.release(h)
eng.close()
