"""Host logic of the drop-in boundary, on CPU: PCD I/O, the HomogeneousMatrix duck-type against the reference's own
artelib (golden vectors), the config singleton, and — when /root/reference is present — the reference's UNMODIFIED
run_scanmatcher.scanmatcher() running on top of the drop-in keyframemanager (arithmetic supplied by the oracle test
double, since there is no GPU in this container)."""
import os
import sys
import types

import numpy as np
import pytest

from lidar_slam_arvc_b200 import euroc_synth, pcd, runtime, synth
from lidar_slam_arvc_b200.homogeneousmatrix import HomogeneousMatrix, rot2quaternion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "lidar_slam_arvc_b200", "dropin")
REF = "/root/reference"


def test_pcd_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    xyz = rng.normal(size=(1000, 3)).astype(np.float32)
    xyz[3] = np.nan
    for binary in (True, False):
        fn = str(tmp_path / ("a_%d.pcd" % binary))
        pcd.write_pcd_xyz(fn, xyz, binary=binary)
        back = pcd.read_pcd_xyz(fn)
        assert back.dtype == np.float32 and back.shape == xyz.shape
        np.testing.assert_array_equal(back, xyz)
    # extra fields and float64 coordinates
    fn = str(tmp_path / "b.pcd")
    rec = np.zeros(5, dtype=[("x", "<f8"), ("y", "<f8"), ("z", "<f8"), ("intensity", "<f4")])
    rec["x"], rec["y"], rec["z"] = np.arange(5), np.arange(5) * 2.0, 0.1
    with open(fn, "wb") as f:
        f.write(b"VERSION 0.7\nFIELDS x y z intensity\nSIZE 8 8 8 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 5\nHEIGHT 1\nPOINTS 5\nDATA binary\n")
        f.write(rec.tobytes())
    back = pcd.read_pcd_xyz(fn)
    assert back.dtype == np.float64
    np.testing.assert_array_equal(back[:, 1], np.arange(5) * 2.0)
    # binary_compressed (LZF): literal-only stream written by the test helper, and a hand-made stream with back references
    fn = str(tmp_path / "c.pcd")
    pcd.write_pcd_xyz(fn, xyz, compressed=True)
    np.testing.assert_array_equal(pcd.read_pcd_xyz(fn), xyz)
    from lidar_slam_arvc_b200.engine import lzf_decompress
    stream = bytes([2]) + b"abc" + bytes([(4 << 5) | 0, 2]) + bytes([0]) + b"Z" + bytes([(7 << 5) | 0, 3, 0])
    assert lzf_decompress(stream, 3 + 6 + 1 + 12) == b"abcabcabcZ" + b"Z" * 12
    with pytest.raises(ValueError):
        lzf_decompress(bytes([(1 << 5) | 0, 9]), 3)            # back reference before the start of the output
    # PCL-style padding: repeated fields named "_" with COUNT > 1 (o3d.io.read_point_cloud reads such files)
    fn = str(tmp_path / "d.pcd")
    rec = np.zeros(7, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("p0", "u1", (4,)), ("intensity", "<f4"), ("p1", "u1", (12,))])
    rec["x"], rec["y"], rec["z"], rec["intensity"] = np.arange(7), -np.arange(7), 0.5, 9.0
    with open(fn, "wb") as f:
        f.write(b"VERSION 0.7\nFIELDS x y z _ intensity _\nSIZE 4 4 4 1 4 1\nTYPE F F F U F U\nCOUNT 1 1 1 4 1 12\nWIDTH 7\nHEIGHT 1\nPOINTS 7\nDATA binary\n")
        f.write(rec.tobytes())
    back = pcd.read_pcd_xyz(fn)
    assert back.dtype == np.float32
    np.testing.assert_array_equal(back, np.stack([rec["x"], rec["y"], rec["z"]], axis=1))
    # an unsupported field type is reported, not a bare KeyError
    with open(fn, "wb") as f:
        f.write(b"VERSION 0.7\nFIELDS x y z w\nSIZE 4 4 4 3\nTYPE F F F U\nCOUNT 1 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n" + b"\0" * 15)
    with pytest.raises(ValueError, match="unsupported PCD field type"):
        pcd.read_pcd_xyz(fn)
    # the allocator hook (page-locked staging on the load path): same values, caller-supplied memory
    made = []

    def alloc(n_rows, dtype):
        made.append(np.empty((n_rows, 3), dtype=dtype))
        return made[-1]
    for binary, compressed in ((True, False), (False, False), (True, True)):
        fn = str(tmp_path / ("h_%d_%d.pcd" % (binary, compressed)))
        pcd.write_pcd_xyz(fn, xyz, binary=binary, compressed=compressed)
        out = pcd.read_pcd_xyz(fn, alloc=alloc)
        assert out is made[-1]
        np.testing.assert_array_equal(out, xyz)
    empty = str(tmp_path / "e.pcd")
    pcd.write_pcd_xyz(empty, np.zeros((0, 3)))
    assert pcd.read_pcd_xyz(empty).shape == (0, 3)


def test_homogeneous_matrix_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "se3_helpers.npz"))
    for T, Ti, P, q in zip(g["mats"], g["invs"], g["prods"], g["quats"]):
        H = HomogeneousMatrix(T)
        np.testing.assert_allclose(H.inv().array, Ti, atol=1e-12)
        np.testing.assert_allclose((H * H.inv() * H).array, P, atol=1e-12)
        np.testing.assert_allclose(rot2quaternion(T), q, atol=1e-12)
        np.testing.assert_array_equal(H.pos(), T[:3, 3])


def test_config_singleton_has_reference_fields():
    sys.path.insert(0, DROPIN)
    try:
        for m in [k for k in sys.modules if k == "config" or k.startswith("config.")]:
            del sys.modules[m]
        from config import ICP_PARAMETERS as P
        assert (P.max_radius, P.min_radius, P.min_height, P.max_height) == (35, 0.5, -1.0, 50.0)
        assert P.voxel_size is None and P.max_nn == 300 and P.distance_threshold == 10.0
        assert (P.relative_fitness, P.relative_rmse, P.max_iteration) == (1e-6, 1e-6, 30)
        if os.path.isdir(REF):
            import yaml
            ref = yaml.safe_load(open(os.path.join(REF, "config", "icp_parameters.yaml")))
            ours = yaml.safe_load(open(os.path.join(DROPIN, "config", "icp_parameters.yaml")))
            ours.pop("icp_criteria")
            assert ours == ref                           # same schema, same defaults
    finally:
        sys.path.remove(DROPIN)


def _stub_missing_modules():
    for name in ("matplotlib", "matplotlib.pyplot", "pyproj", "open3d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(sys.modules["pyproj"], "Proj"):
        sys.modules["pyproj"].Proj = lambda *a, **k: None


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_reference_scanmatcher_runs_unmodified_on_dropin(tmp_path):
    """run_scanmatcher.scanmatcher(directory) — the reference's own driver, byte for byte — on the drop-in classes."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_engine import OracleEngine
    from oracle import oracle as orc
    seq = synth.Sequence(5, synth.TINY_16, start=30.0)
    d = str(tmp_path / "euroc")
    times = euroc_synth.write_euroc_tree(d, seq)
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    fake = OracleEngine()
    runtime.set_engine(fake)
    try:
        _stub_missing_modules()
        for m in [k for k in sys.modules if k.split(".")[0] in ("config", "keyframemanager")]:
            del sys.modules[m]
        sys.path.insert(0, REF)          # reference second ...
        sys.path.insert(0, DROPIN)       # ... drop-in first
        import run_scanmatcher
        import keyframemanager.keyframemanager as kfm
        assert kfm.__file__.startswith(DROPIN)
        run_scanmatcher.scanmatcher(directory=d)
        import pandas as pd
        rel = pd.read_csv(os.path.join(d, "robot0", "scanmatcher", "scanmatcher_relative.csv"))
        glob = pd.read_csv(os.path.join(d, "robot0", "scanmatcher", "scanmatcher_global.csv"))
        assert len(rel) == 4 and len(glob) == 5
        assert list(rel.columns)[1:] == ["#timestamp [ns]", "x", "y", "z", "qx", "qy", "qz", "qw"]
        np.testing.assert_array_equal(rel["#timestamp [ns]"].to_numpy(), times[:4])
        # the relative transforms written by the reference driver are the oracle's ICP results for (i, i+1), init = odometry
        pre = [orc.preprocess(s) for s in seq.scans]
        for i in range(4):
            ref = orc.icp(pre[i + 1][0], pre[i][0], pre[i][1], seq.relative_odo(i, i + 1), orc.P2PLANE)
            np.testing.assert_allclose(rel.loc[i, ["x", "y", "z"]].to_numpy(dtype=float), ref.transformation[:3, 3], atol=1e-6)
            gt = seq.relative_gt(i, i + 1)
            assert np.linalg.norm(ref.transformation[:3, 3] - gt[:3, 3]) < 0.05
        # call pattern of run_scanmatcher.py:191-213: one preprocess per keyframe, one ICP per consecutive pair
        assert [c for c in fake.calls if c[0] == "icp_batch"] == [("icp_batch", 1)] * 4
        assert len([c for c in fake.calls if c[0] == "preprocess"]) == 5
    finally:
        runtime.set_engine(None)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]


def test_two_planes_pieces_match_reference_golden(golden_dir):
    """'icp2planes' (SURVEY.md §8 f-2): the reference's own numpy split (keyframe.py:438-461) and component merge
    (keyframe.py:282-292), generated by importing the reference (tests/golden/make_golden.py)."""
    from oracle import oracle as orc
    g = np.load(os.path.join(golden_dir, "two_planes.npz"))
    near, far = orc.segment_plane(g["cloud"], g["plane"], 0.4)
    np.testing.assert_array_equal(g["cloud"][near], g["near"])
    np.testing.assert_array_equal(g["cloud"][far], g["far"])
    assert 0 < len(near) < len(g["cloud"])
    for T, t in zip(g["mats"], g["t2v3"]):
        np.testing.assert_allclose(HomogeneousMatrix(T).t2v(n=3), t, rtol=0, atol=1e-12)
    sys.path.insert(0, DROPIN)
    saved = dict(sys.modules)
    try:
        for m in [k for k in sys.modules if k.split(".")[0] in ("config", "keyframemanager")]:
            del sys.modules[m]
        from keyframemanager.keyframe import merge_two_planes
        for k in range(16):
            M = merge_two_planes(g["mats"][k], g["mats"][k + 16])
            np.testing.assert_allclose(M.array, g["merged"][k], rtol=0, atol=1e-12)
    finally:
        sys.path.remove(DROPIN)
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]


def test_oracle_plane_fit_is_reproducible_and_sane():
    from oracle import oracle as orc
    seq = synth.Sequence(1, synth.SMALL_32, start=12.0)
    p, _ = orc.preprocess(seq.scans[0], method="icppointpoint")
    pl, n_in = orc.fit_plane(p, seed=0)
    pl2, n2 = orc.fit_plane(p, seed=0)
    np.testing.assert_array_equal(pl, pl2)
    assert n_in == n2 > 0.5 * (p[:, 2] < -0.5).sum()
    assert abs(np.linalg.norm(pl[:3]) - 1) < 1e-12 and abs(abs(pl[2]) - 1) < 1e-3 and abs(abs(pl[3]) - synth.SENSOR_HEIGHT) < 0.02
    # n_in counts the inliers of the winning 3-point hypothesis; the returned model is Open3D's least-squares refit to
    # them (GetPlaneFromPoints), so it explains at least about as many points and is close to the SVD plane of its inliers
    low = p[p[:, 2] < -0.5]
    m = np.abs(low @ pl[:3] + pl[3]) < 0.01
    assert m.sum() >= 0.95 * n_in
    q = low[m]
    nn = np.linalg.svd(q - q.mean(0))[2][2]
    nn = nn * np.sign(nn @ pl[:3])
    assert np.abs(nn - pl[:3]).max() < 2e-3 and abs(-nn @ q.mean(0) - pl[3]) < 2e-3
    pl3, _ = orc.fit_plane(p, seed=1)
    assert np.abs(pl3 - pl).max() < 0.02                 # another seed: another, equally good, hypothesis
    # too few points below the height: no model
    assert orc.fit_plane(p, max_z=-5.0)[1] == 0


def test_dropin_icp2planes_on_oracle_engine(tmp_path):
    """Host logic of the 'icp2planes' method: preprocess -> plane -> split -> two registrations in one batch -> merge."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_engine import OracleEngine
    from oracle import oracle as orc
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    fake = OracleEngine()
    runtime.set_engine(fake)
    try:
        for m in [k for k in sys.modules if k.split(".")[0] in ("config", "keyframemanager")]:
            del sys.modules[m]
        sys.path.insert(0, DROPIN)
        import keyframemanager.keyframemanager as kfm
        from keyframemanager.keyframe import merge_two_planes
        seq = synth.Sequence(3, synth.TINY_16, start=30.0)
        d = str(tmp_path / "euroc")
        times = euroc_synth.write_euroc_tree(d, seq, method="icp2planes")
        km = kfm.KeyFrameManager(directory=d, scan_times=times, voxel_size=None, method="icp2planes")
        km.add_keyframes(keyframe_sampling=1)
        km.load_pointclouds()
        for i in range(3):
            km.pre_process(i)
        T01 = km.compute_transformation(0, 1, HomogeneousMatrix(seq.relative_odo(0, 1)))
        assert [c for c in fake.calls if c[0] == "icp_batch"] == [("icp_batch", 2)]
        pre = [orc.preprocess_two_planes(s) for s in seq.scans]
        np.testing.assert_array_equal(km.keyframes[0].plane_model, pre[0][0])
        np.testing.assert_array_equal(km.keyframes[0].pointcloud_ground_plane.points, pre[0][1][0])
        np.testing.assert_array_equal(km.keyframes[0].pointcloud_non_ground_plane.points, pre[0][2][0])
        init = seq.relative_odo(0, 1)
        ra = orc.icp(pre[1][1][0], pre[0][1][0], pre[0][1][1], init, orc.P2PLANE)
        rb = orc.icp(pre[1][2][0], pre[0][2][0], pre[0][2][1], init, orc.P2PLANE)
        want = merge_two_planes(ra.transformation, rb.transformation).array
        np.testing.assert_allclose(T01.array, want, rtol=0, atol=1e-12)
        gt = seq.relative_gt(0, 1)
        assert np.linalg.norm(T01.array[:3, 3] - gt[:3, 3]) < 0.05
        # batched variant and a caller-supplied plane model
        Ts, _ = km.compute_transformations([(0, 1), (1, 2)], [HomogeneousMatrix(seq.relative_odo(0, 1)), HomogeneousMatrix(seq.relative_odo(1, 2))])
        np.testing.assert_array_equal(Ts[0].array, T01.array)
        kf = kfm.KeyFrame(d, times[0], None)
        kf.load_pointcloud()
        kf.fixed_plane_model = np.array([0.0, 0.0, 1.0, 0.69])
        kf.pre_process(method="icp2planes")
        assert ("fit_plane", kf._scan_id) not in fake.calls
        np.testing.assert_array_equal(kf.plane_model, [0.0, 0.0, 1.0, 0.69])
        pw = orc.preprocess_two_planes(seq.scans[0], plane_model=[0.0, 0.0, 1.0, 0.69])
        np.testing.assert_array_equal(kf.pointcloud_ground_plane.points, pw[1][0])
        # without a pinned model the plane is recomputed on every pre_process, as the reference does (keyframe.py:173):
        # new points must not be split by the plane of the old ones
        kf.fixed_plane_model = None
        kf.set_points(seq.scans[1])
        kf.pre_process(method="icp2planes")
        np.testing.assert_array_equal(kf.plane_model, pre[1][0])
        np.testing.assert_array_equal(kf.pointcloud_ground_plane.points, pre[1][1][0])
        km.unload_pointcloud(0)
        assert km.keyframes[0].pointcloud_ground_plane is None
    finally:
        runtime.set_engine(None)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_reference_scanmatcher_runs_icp2planes_on_dropin(tmp_path):
    """The reference's unmodified driver with `method: icp2planes` in scanmatcher_parameters.yaml (run_scanmatcher.py:159-166):
    every pair becomes one device batch of two registrations (ground / non-ground parts)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_engine import OracleEngine
    seq = synth.Sequence(4, synth.TINY_16, start=30.0)
    d = str(tmp_path / "euroc")
    euroc_synth.write_euroc_tree(d, seq, method="icp2planes")
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    fake = OracleEngine()
    runtime.set_engine(fake)
    try:
        _stub_missing_modules()
        for m in [k for k in sys.modules if k.split(".")[0] in ("config", "keyframemanager")]:
            del sys.modules[m]
        sys.path.insert(0, REF)
        sys.path.insert(0, DROPIN)
        import run_scanmatcher
        run_scanmatcher.scanmatcher(directory=d)
        import pandas as pd
        rel = pd.read_csv(os.path.join(d, "robot0", "scanmatcher", "scanmatcher_relative.csv"))
        assert len(rel) == 3
        assert [c for c in fake.calls if c[0] == "icp_batch"] == [("icp_batch", 2)] * 3
        assert len([c for c in fake.calls if c[0] == "fit_plane"]) == 4 and len([c for c in fake.calls if c[0] == "split_plane"]) == 4
        for i in range(3):
            gt = seq.relative_gt(i, i + 1)
            assert np.linalg.norm(rel.loc[i, ["x", "y", "z"]].to_numpy(dtype=float) - gt[:3, 3]) < 0.06
    finally:
        runtime.set_engine(None)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k not in saved_mods:
                del sys.modules[k]
