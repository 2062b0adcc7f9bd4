"""Developer probe: normals kernel times + hand-back statistics for n OS1-64 scans (ARVC_LIB_VARIANT selects a build)."""
import os
import sys

sys.path.insert(0, ".")
from lidar_slam_arvc_b200 import engine, synth  # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sensor = {"64": synth.OS1_64, "128": synth.OS_128}[sys.argv[2] if len(sys.argv) > 2 else "64"]
seq = synth.Sequence(n_scans, sensor, start=30.0, workers=8)
eng = engine.Engine(0)
pp = eng.make_preprocess_params()
ids = list(range(n_scans))
for k in ids:
    eng.upload(k, seq.scans[k])
best = None
for rep in range(4):
    eng.invalidate(ids)
    eng.sync()
    eng.profile_enable(True)
    eng.preprocess(ids, pp)
    prof = eng.profile_report()
    eng.profile_enable(False)
    t = {k: v[1] for k, v in prof.items() if k.startswith("normals")}
    if best is None or sum(t.values()) < sum(best.values()):
        best = t
tot = {}
for k in ids:
    for a, b in eng.get_counters(k).items():
        tot[a] = tot.get(a, 0) + b
print("%-10s normals total %.3f ms/scan  %s  per-point %.1f%% (blocks %d, points %d) trial blocks %d of %d" % (
    os.path.basename(os.environ.get("ARVC_LIB_VARIANT", "default")), sum(best.values()) / n_scans,
    {k: round(v / n_scans, 4) for k, v in best.items()}, 100.0 * tot["normals_per_point"] / tot["n_points"],
    tot["normals_blocks_handed_back"], tot["normals_points_handed_back"], tot["normals_trial_blocks"], tot["n_points"] // 32))
if os.environ.get("ARVC_DEBUG_NORMALS"):
    nb = tot["n_points"] / 32.0
    print("   per block: records streamed %.0f, tile %.0f; per point: neighbours %.1f; points: one sweep %d, sweep+select %d, trial %d" % (
        tot["dbg_records_streamed"] / nb, tot["dbg_tile_records"] / nb, tot["dbg_neighbours"] / tot["n_points"],
        tot["dbg_points_one_sweep"], tot["dbg_points_sweep_then_select"], tot["dbg_points_trial"]))
    print("   blocks handed back: %d too wide (cell box > 256 cells), %d tile overflow" % (tot["dbg_blocks_too_wide"], tot["normals_blocks_handed_back"] - tot["dbg_blocks_too_wide"]))
eng.close()
