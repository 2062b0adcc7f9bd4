"""Host logic of the configs[4] pipeline (lidar_slam_arvc_b200/pipeline.py) on CPU: batched scan-matcher front end with the
one-scan halo, the run_graphSLAM.py loop around the drop-in LoopClosing / KeyFrameManager, the stand-in pose graph.
Arithmetic supplied by the oracle test double (no GPU in this container)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from lidar_slam_arvc_b200 import pipeline  # noqa: E402
from lidar_slam_arvc_b200.homogeneousmatrix import HomogeneousMatrix  # noqa: E402


def test_scan_matcher_batches_with_halo_equal_pair_by_pair():
    from fake_engine import OracleEngine
    from test_loopclosing_cpu import small_loop_sequence
    from oracle import oracle as orc
    seq = small_loop_sequence(n_scans=9)
    odo = [seq.relative_odo(k, k + 1) for k in range(8)]
    eng = OracleEngine()
    rel, recs = pipeline.scan_matcher(eng, seq.scans, odo, batch=4)          # batches [0..3], [3..6], [6..8]: one-scan halo
    assert rel.shape == (8, 4, 4) and not eng.raw                            # every scan was freed again
    pre = [orc.preprocess(s) for s in seq.scans]
    for k in range(8):
        ref = orc.icp(pre[k + 1][0], pre[k][0], pre[k][1], odo[k], orc.P2PLANE)
        np.testing.assert_array_equal(rel[k], ref.transformation)
        assert recs["updates"][k] == ref.updates


def test_pose_graph_stand_in_pulls_a_drifting_chain_onto_its_loop_closure():
    g = pipeline.PoseGraphStandIn()
    g.init_graph()
    step = np.eye(4)
    step[0, 3] = 1.0
    drift = step.copy()
    drift[1, 3] = 0.05                                                       # every odometry step drifts 5 cm sideways
    for k in range(10):
        g.add_initial_estimate(HomogeneousMatrix(drift), k + 1)
        g.add_edge(HomogeneousMatrix(drift), k, k + 1, 'SM')
    assert abs(g.current_estimate.atPose3(10).matrix()[1, 3] - 0.5) < 1e-12
    closure = np.eye(4)
    closure[0, 3] = 10.0                                                     # a strong observation: pose 10 is straight ahead of pose 0
    for _ in range(200):
        g.add_edge(HomogeneousMatrix(closure), 0, 10, 'SM')
    g.optimize()
    assert g.current_estimate.exists(10) and not g.current_estimate.exists(11)
    assert abs(g.current_estimate.atPose3(10).matrix()[1, 3]) < 0.05
    assert abs(g.current_estimate.atPose3(10).matrix()[0, 3] - 10.0) < 0.05


def test_backend_loop_runs_loop_closing_on_the_dropin():
    from fake_engine import OracleEngine
    from test_loopclosing_cpu import small_loop_sequence
    seq = small_loop_sequence()
    n = len(seq.scans)
    sm = [seq.relative_gt(k, k + 1) for k in range(n - 1)]                   # a perfect front end: the loop closes by itself
    odo = [seq.relative_odo(k, k + 1) for k in range(n - 1)]
    eng = OracleEngine()
    rep = pipeline.run_backend(eng, seq.scans, sm, odo, skip_loop_closing=33, skip_optimization=33, number_of_triplets_loop_closing=2, seed=5)
    assert rep["loop_closing_calls"] == 3 and rep["optimizations"] >= 3       # i = 0, 33 and 66 (also the last step)
    assert rep["loop_closing_pairs"] > 0 and rep["loop_closing_pairs"] % 2 == 0
    assert rep["loop_closure_edges"] > 0                                      # consistent triplets were found and added as edges
    assert rep["edges"] == 2 * (n - 1) + rep["loop_closure_edges"]
    assert rep["keyframes_resident_at_end"] > 0
    assert [c for c in eng.calls if c[0] == "icp_batch"]                      # one device batch per loop-closing invocation
    assert len([c for c in eng.calls if c[0] == "icp_batch"]) <= rep["loop_closing_calls"]
    assert abs(rep["total_s"] - (rep["loop_closing_s"] + rep["optimize_s"] + rep["other_host_s"])) < 1e-9
