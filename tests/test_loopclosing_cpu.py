"""SURVEY.md §8 f-1 on CPU: the batched drop-in graphslam.loopclosing.LoopClosing against the reference's UNMODIFIED
class (imported from /root/reference when present), both on top of the drop-in KeyFrameManager whose arithmetic is
the oracle test double.  Same seed -> same random draws -> the same edges with the same transforms, but one
device batch instead of 2 x triplets sequential registrations."""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from lidar_slam_arvc_b200 import euroc_synth, runtime, synth
from lidar_slam_arvc_b200.homogeneousmatrix import HomogeneousMatrix, rot2euler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "lidar_slam_arvc_b200", "dropin")
REF = "/root/reference"


def test_euler_pair_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "se3_helpers.npz"))
    for T, e1, e2 in zip(g["mats"], g["eulers"], g["eulers2"]):
        m1, m2 = rot2euler(T)
        np.testing.assert_allclose(m1, e1, atol=1e-12)
        np.testing.assert_allclose(m2, e2, atol=1e-12)
        a, b = HomogeneousMatrix(T).euler()
        np.testing.assert_array_equal(a.abg, m1)
        np.testing.assert_array_equal(b.abg, m2)
    for R, e1, e2 in zip(g["gimbal_mats"], g["gimbal_e1"], g["gimbal_e2"]):     # beta = +-pi/2 branch
        m1, m2 = rot2euler(R)
        np.testing.assert_allclose(m1, e1, atol=1e-12)
        np.testing.assert_allclose(m2, e2, atol=1e-12)


class _Pose:
    def __init__(self, M):
        self._M = M

    def matrix(self):
        return self._M.copy()


class _Values:
    def __init__(self, mats):
        self._mats = mats

    def exists(self, i):
        return 0 <= i < len(self._mats)

    def atPose3(self, i):
        return _Pose(self._mats[i])


class FakeGraphSLAM:
    """Duck type of graphslam.graphSLAM.GraphSLAM as LoopClosing uses it (gtsam is not installed here)."""

    def __init__(self, lidar_poses, T0_gps):
        self.T0_gps = T0_gps
        self.current_estimate = _Values([P @ T0_gps.array for P in lidar_poses])
        self.edges = []

    def add_edge(self, atb, i, j, sigmas):
        self.edges.append((int(i), int(j), np.array(atb.array), sigmas))


def small_loop_sequence(n_scans=68, step=0.8):
    """One and a bit laps of a 45 m loop: the last poses revisit the first ones, consecutive poses 0.8 m apart so that
    the triplet gate (1 m < d(j1, j2) < 2 m, 1 < |j1 - j2|) has solutions."""
    world = synth.World(seed=7, n_boxes=24, n_cylinders=12, outer=(20.0, 14.0), inner=(10.0, 4.0))
    return synth.Sequence(n_scans, synth.TINY_16, world=world, step=step)


def noisy_estimate(seq, seed=3):
    rng = np.random.default_rng(seed)
    return [P @ synth.pose_matrix(*rng.normal(0, 0.03, 3), np.deg2rad(rng.normal(0, 0.5))) for P in seq.poses]


def _stub_missing_modules():
    for name in ("matplotlib", "matplotlib.pyplot", "pyproj", "open3d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


@pytest.fixture()
def dropin_env():
    """Drop-in first, reference second on sys.path; everything imported inside is forgotten afterwards."""
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_engine import OracleEngine
    fake = OracleEngine()
    runtime.set_engine(fake)
    _stub_missing_modules()
    for m in [k for k in sys.modules if k.split(".")[0] in ("config", "keyframemanager", "graphslam")]:
        del sys.modules[m]
    if os.path.isdir(REF):
        sys.path.insert(0, REF)
    sys.path.insert(0, DROPIN)
    try:
        yield fake
    finally:
        runtime.set_engine(None)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]


def _manager(tmp_path, seq):
    import keyframemanager.keyframemanager as kfm
    d = str(tmp_path / "euroc")
    times = euroc_synth.write_euroc_tree(d, seq)
    km = kfm.KeyFrameManager(directory=d, scan_times=times, voxel_size=None, method="icppointplane")
    km.add_keyframes(keyframe_sampling=1)
    return km


def test_package_overlay_resolves_only_loopclosing(dropin_env):
    import graphslam.loopclosing as lc
    assert lc.__file__.startswith(DROPIN)
    if os.path.isdir(REF):
        spec = importlib.util.find_spec("graphslam.graphSLAM")          # not replaced: still the reference's file
        assert spec is not None and spec.origin.startswith(REF)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_batched_loop_closing_adds_the_reference_edges(dropin_env, tmp_path):
    fake = dropin_env
    seq = small_loop_sequence()
    km = _manager(tmp_path, seq)
    from artelib.homogeneousmatrix import HomogeneousMatrix as RefH
    T0_gps = RefH(synth.pose_matrix(0.36, 0.0, 0.0, 0.0))
    est = noisy_estimate(seq)
    spec = importlib.util.spec_from_file_location("ref_loopclosing", os.path.join(REF, "graphslam", "loopclosing.py"))
    ref_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_mod)
    import graphslam.loopclosing as ours_mod

    last = len(seq.poses) - 1
    # ---- triangle procedure
    g_ref = FakeGraphSLAM(est, T0_gps)
    np.random.seed(11)
    added_ref = ref_mod.LoopClosing(g_ref, distance_backwards=7, radius_threshold=5.0).loop_closing_triangle(
        current_index=last, number_of_triplets_loop_closing=3, keyframe_manager=km)
    n_seq_calls = len([c for c in fake.calls if c[0] == "icp_batch"])
    assert n_seq_calls == 6 and len(added_ref) > 0                          # 2 registrations per triplet, one at a time

    fake.calls.clear()
    g_new = FakeGraphSLAM(est, T0_gps)
    np.random.seed(11)
    added_new = ours_mod.LoopClosing(g_new, distance_backwards=7, radius_threshold=5.0).loop_closing_triangle(
        current_index=last, number_of_triplets_loop_closing=3, keyframe_manager=km)
    assert [c for c in fake.calls if c[0] == "icp_batch"] == [("icp_batch", 6)]      # ONE batch
    assert added_new == added_ref
    assert len(g_new.edges) == len(g_ref.edges)
    for a, b in zip(g_ref.edges, g_new.edges):
        assert a[:2] == b[:2] and a[3] == b[3] == 'SM'
        np.testing.assert_allclose(b[2], a[2], rtol=0, atol=1e-12)
        gt = T0_gps.inv().array @ seq.relative_gt(a[0], a[1]) @ T0_gps.array
        assert np.linalg.norm(b[2][:3, 3] - gt[:3, 3]) < 0.1                 # and they are good loop closures

    # ---- simple procedure
    fake.calls.clear()
    g_ref, g_new = FakeGraphSLAM(est, T0_gps), FakeGraphSLAM(est, T0_gps)
    np.random.seed(5)
    ref_mod.LoopClosing(g_ref).loop_closing_simple(current_index=last, number_of_candidates_DA=4, keyframe_manager=km)
    np.random.seed(5)
    ours_mod.LoopClosing(g_new).loop_closing_simple(current_index=last, number_of_candidates_DA=4, keyframe_manager=km)
    assert [c for c in fake.calls if c[0] == "icp_batch"] == [("icp_batch", 1)] * 4 + [("icp_batch", 4)]
    assert len(g_ref.edges) == len(g_new.edges) == 4
    for a, b in zip(g_ref.edges, g_new.edges):
        assert a[:2] == b[:2]
        np.testing.assert_allclose(b[2], a[2], rtol=0, atol=1e-12)

    # ---- the host-side candidate search is the reference's
    lr, ln = ref_mod.LoopClosing(g_ref), ours_mod.LoopClosing(g_new)
    assert [list(map(int, t)) for t in lr.find_feasible_triplets(last)] == [list(map(int, t)) for t in ln.find_feasible_triplets(last)]
    assert lr.find_index_backwards() == ln.find_index_backwards()
    np.testing.assert_array_equal(lr.find_candidates(), ln.find_candidates())


def test_loop_closing_without_candidates_and_sequential_fallback(dropin_env, tmp_path):
    """No revisit -> nothing to do (reference: returns None); a manager without the batched entry point is served
    pair by pair with the reference's call sequence."""
    import graphslam.loopclosing as ours_mod
    seq = synth.Sequence(6, synth.TINY_16, start=30.0)
    km = _manager(tmp_path, seq)
    from lidar_slam_arvc_b200.homogeneousmatrix import result_type
    T0 = result_type()(np.eye(4))              # the reference's class when its tree is importable, ours otherwise
    g = FakeGraphSLAM(seq.poses, T0)
    lc = ours_mod.LoopClosing(g, distance_backwards=1.0, radius_threshold=5.0)
    assert lc.loop_closing_triangle(current_index=5, number_of_triplets_loop_closing=3, keyframe_manager=km) is None
    assert g.edges == []

    class Sequential:                      # the reference manager's surface only
        def __init__(self, km):
            self.km, self.log = km, []

        def load_pointcloud(self, i):
            self.log.append(("load", i)); self.km.load_pointcloud(i)

        def pre_process(self, i):
            self.log.append(("pre", i)); self.km.pre_process(i)

        def compute_transformation(self, i, j, Tij):
            self.log.append(("icp", i, j)); return self.km.compute_transformation(i, j, Tij)

    sq = Sequential(km)
    np.random.seed(1)
    lc.loop_closing_simple(current_index=5, number_of_candidates_DA=2, keyframe_manager=sq)
    assert len(g.edges) == 2
    assert sq.log[:5] == [("load", 5), ("pre", 5), ("load", g.edges[0][1]), ("pre", g.edges[0][1]), ("icp", 5, g.edges[0][1])]
    g2 = FakeGraphSLAM(seq.poses, T0)
    np.random.seed(1)
    ours_mod.LoopClosing(g2, distance_backwards=1.0).loop_closing_simple(current_index=5, number_of_candidates_DA=2, keyframe_manager=km)
    for a, b in zip(g.edges, g2.edges):
        assert a[:2] == b[:2]
        np.testing.assert_allclose(b[2], a[2], rtol=0, atol=1e-12)
