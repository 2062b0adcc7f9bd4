"""Developer probe: step-time stability with / without a background `nvidia-smi -lms` poller and with torch loaded."""
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 32
seq = synth.Sequence(P + 1, synth.OS1_64, start=30.0, workers=8)
eng = engine.Engine(0)
pp = eng.make_preprocess_params()
ip = eng.make_icp_params(engine.P2PLANE)
ids = np.arange(P + 1)
init = np.array([seq.relative_odo(a, a + 1) for a in range(P)])
for k in ids:
    eng.upload(k, seq.scans[k])


def run(tag, n=12):
    ts = []
    for rep in range(n):
        eng.sync()
        t0 = time.perf_counter()
        eng.invalidate(ids)
        eng.preprocess(ids, pp)
        eng.icp_batch(ids[:-1], ids[1:], init, ip)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(tag, "ms/step: min %.1f median %.1f max %.1f" % (min(ts), np.median(ts), max(ts)), [round(t, 1) for t in ts])


run("warmup", 3)
run("quiet")
Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
for ms in ("1000", "250"):
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=" + Q, "--format=csv,noheader,nounits", "-lms", ms], stdout=subprocess.DEVNULL)
    time.sleep(1.0)
    run("nvidia-smi -lms " + ms)
    p.terminate()
    p.wait()
import torch  # noqa: E402
torch.cuda.init()
x = torch.zeros(10, device="cuda")
run("torch loaded")
eng.profile_enable(True)
run("event profiling on")
print(len(eng.profile_report()))
run("event profiling on, 2nd")
eng.profile_enable(False)
