"""Per-point search statistics of k_normals (ARVC_DEBUG_NORMALS=1 prints every 97th point): how many points take which path."""
import os, re, subprocess, sys
if os.environ.get("ARVC_DEBUG_NORMALS") != "1":
    env = dict(os.environ, ARVC_DEBUG_NORMALS="1")
    out = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True).stdout
    rows = [dict((k, float(v)) for k, v in re.findall(r"(\w+)=([-\d.]+)", l)) for l in out.splitlines() if l.startswith("NRM")]
    n = len(rows)
    single = [r for r in rows if r["total"] <= 300]
    trial_ok = [r for r in rows if r["used_trial"] == 1]
    full2 = [r for r in rows if r["used_trial"] == 0 and r["total"] > 300]
    spec = [r for r in full2 if 0 <= r["n_in"] <= 300]
    print("samples", n)
    print("single pass (total<=k): %.3f  mean total %.0f" % (len(single) / n, sum(r["total"] for r in single) / max(len(single), 1)))
    print("trial radius used:      %.3f  mean total_try %.0f" % (len(trial_ok) / n, sum(r["total_try"] for r in trial_ok) / max(len(trial_ok), 1)))
    print("full radius, 2 passes:  %.3f  mean total %.0f" % (len(full2) / n, sum(r["total"] for r in full2) / max(len(full2), 1)))
    print("  of which n_in<=k:     %.3f  mean total %.0f" % (len(spec) / n, sum(r["total"] for r in spec) / max(len(spec), 1)))
    tf = [r for r in rows if r["ntry"] > 0 and r["used_trial"] == 0]
    print("trial failed -> full:   %.3f" % (len(tf) / n))
    sys.exit(0)
sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth
seq = synth.Sequence(4, synth.OS1_64, start=0.0, step=7.0, workers=4)
eng = engine.Engine(0)
for k, s in enumerate(seq.scans):
    eng.upload(k, s)
eng.preprocess(list(range(4)), eng.make_preprocess_params())
eng.sync()
